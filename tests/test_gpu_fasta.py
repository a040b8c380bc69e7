"""FASTA ingest on the device (tnt_engine_add_fasta) against the oracle's restatement of the
reference reader (oracle/tnt_oracle_fasta.c, pinned to the compiled reference by
tests/test_oracle_golden.py) and against the committed golden vectors: record table, deflines,
fragment queue and the base codes of every registered fragment bit-exact; a search on the
ingested database equal to a search on the fragments the oracle reader produces."""
import base64
import json
import os

import numpy as np
import pytest

import gen
import harness as H
from test_gpu_parity import hit_floats, hit_key, to_opts

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng(engine_lib):
    from thermonucleotideblast_b200 import Engine
    e = Engine()
    yield e
    e.close()


def check_against(eng, text, want, threshold, overlap):
    """`want`: harness.fasta_records() view (offset, approx_len, defline, codes, pieces)."""
    eng.clear_targets()
    recs, frags = eng.add_fasta(text, fragment_threshold=threshold, overlap=overlap)
    assert len(recs) == len(want)
    for r, (off, alen, defline, codes, pieces) in zip(recs, want):
        if off is not None:
            assert r.text_offset == off
        assert r.text_bytes == alen
        assert text[r.defline_offset:r.defline_offset + r.defline_len].decode("latin-1") == defline
        assert r.bases == len(codes)
        mine = frags[r.first_fragment:r.first_fragment + r.n_fragments]
        if threshold:
            assert [(f.start, f.stop) for f in mine] == [(p[0], p[1]) for p in pieces]
        else:
            assert [(f.start, f.stop) for f in mine] == [(0, alen - 1)]
            pieces = [(0, alen - 1, codes)]
        for f, (s0, s1, pc) in zip(mine, pieces):
            assert f.max_stop == alen - 1 and f.len == len(pc)
            if len(pc) == 0:
                assert f.target_id == 0xFFFFFFFF
            else:
                got = eng.target_codes(f.target_id, 0, f.len)
                assert got.tolist() == np.asarray(pc).tolist()


def test_fasta_golden_vectors(eng):
    fixtures = json.load(open(os.path.join(GOLD, "fasta.json")))
    for fx in fixtures:
        text = base64.b64decode(fx["text"])
        want = [(None, r["approx_len"], r["defline"], gen.str_to_codes(r["codes"]),
                 [(p[0], p[1], gen.str_to_codes(p[2])) for p in r["pieces"]]) for r in fx["records"]]
        check_against(eng, text, want, fx["threshold"], fx["overlap"])


def test_fasta_edge_cases_and_random_texts(eng, oracle):
    rng = np.random.default_rng(4242)
    texts = list(gen.FASTA_EDGE_CASES) + [b"", b"no record at all\nACGT\n"]
    for it in range(12):
        texts.append(gen.rand_fasta(rng, n_records=int(rng.integers(1, 9)), max_len=[3000, 40000, 300][it % 3],
                                    width=[60, 80, 0, 7][it % 4], crlf=bool(it % 3 == 1), iupac=0.02))
    for text in texts:
        for threshold, overlap in [(0, 0), (1000, 50), (64, 5)]:
            check_against(eng, text, oracle.fasta_records(text, threshold=threshold, overlap=overlap), threshold, overlap)


def test_fasta_block_and_slab_boundaries(eng, oracle):
    """Texts larger than a parser block (16 KB) and than a slab (64 MB): lines, deflines and
    records that straddle the boundaries, one very long unwrapped line."""
    rng = np.random.default_rng(99)
    # deflines placed right at multiples of 16 KB, with and without '\r'
    parts, size = [], 0
    for k in range(12):
        want_at = (k + 1) * 16384 - int(rng.integers(0, 40))
        fill = want_at - size - len(b">r%d x\n" % k)
        body = rng.choice(list(b"ACGTN"), size=max(fill - 1, 1)).astype(np.uint8).tobytes()
        rec = b">r%d x%s" % (k, b"\r\n" if k % 2 else b"\n") + body + b"\n"
        parts.append(rec)
        size += len(rec)
    text = b"".join(parts) + b">tail  with a long defline " + b"d" * 300 + b"\nACGT"
    check_against(eng, text, oracle.fasta_records(text, threshold=5000, overlap=100), 5000, 100)
    # > 64 MB: two records, the first unwrapped (one 70 MB line), the second wrapped at 80
    n1, n2 = 70_000_000, 3_000_000
    s1 = rng.integers(0, 4, size=n1, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    s2 = lut[rng.integers(0, 4, size=n2, dtype=np.uint8)]
    rows = s2.reshape(-1, 80)
    wrapped = np.concatenate([rows, np.full((rows.shape[0], 1), 10, dtype=np.uint8)], axis=1).reshape(-1)
    text = b">big one\n" + lut[s1].tobytes() + b"\n>second\n" + wrapped.tobytes()
    eng.clear_targets()
    recs, frags = eng.add_fasta(text, fragment_threshold=500000, overlap=2002)
    assert [r.bases for r in recs] == [n1, n2]
    assert [r.text_bytes for r in recs] == [len(b">big one\n") + n1 + 1, len(b">second\n") + wrapped.size]
    codes = [s1, np.searchsorted(lut, s2).astype(np.uint8)]
    for r, c in zip(recs, codes):
        pieces = oracle._pieces(int(r.text_bytes), 500000, 2002, lambda a, b: c[a:b + 1])
        mine = frags[r.first_fragment:r.first_fragment + r.n_fragments]
        assert [(f.start, f.stop, f.len) for f in mine] == [(p[0], p[1], len(p[2])) for p in pieces]
        for f, p in zip(mine, pieces):
            if f.len:
                step = max(f.len // 3, 1)
                for a in (0, step, max(f.len - 4096, 0)):
                    m = min(4096, f.len - a)
                    assert eng.target_codes(f.target_id, a, m).tolist() == p[2][a:a + m].tolist()


def test_fasta_refused_inputs(eng):
    from thermonucleotideblast_b200.engine import EngineError
    for text in (b">abc", b">a\nAC\n>", b">\nACGT\n>r2\nGG\n", b">a\nAC\n>   \r\nACGT\n"):
        eng.clear_targets()
        with pytest.raises(EngineError):
            eng.add_fasta(text)
    eng.clear_targets()


def test_search_on_ingested_fasta_equals_fragment_upload(engine_lib, oracle):
    """End to end: FASTA text -> device parse -> search, against the same search on the fragments
    the oracle reader cuts from the text (registered through tnt_engine_add_targets)."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(31337)
    db = [gen.random_codes(int(rng.integers(30000, 90000)), rng) for _ in range(6)]
    gen.sprinkle_degenerate(db[1], rng, frac=2e-3, n_runs_per_50kb=5)
    assays = gen.make_assays(rng, db, 5, "taqman", variants=2)
    letters = np.frombuffer(b"ACGTIMRSVWYHKDBN", dtype=np.uint8)
    text = b""
    for i, c in enumerate(db):
        s = letters[c]
        body = b"\n".join(s[k:k + 70].tobytes() for k in range(0, s.size, 70))
        text += b">seq%d synthetic\n" % i + body + b"\n"
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    threshold, overlap = 25000, 2002
    want_recs = oracle.fasta_records(text, threshold=threshold, overlap=overlap)
    pieces = [p[2] for r in want_recs for p in r[4] if len(p[2])]

    def keys(e, hits):
        return [(h.target_id, h.assay_index) + hit_key(e, h, assays[h.assay_index]) + hit_floats(h) for h in hits]

    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        recs, frags = e.add_fasta(text, fragment_threshold=threshold, overlap=overlap)
        assert [f.target_id for f in frags if f.len] == list(range(len(pieces)))
        got = keys(e, e.search(to_opts(o)))
    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        e.add_targets(pieces)
        want = keys(e, e.search(to_opts(o)))
    assert got == want and len(got) >= 5


def test_packed_snapshot_round_trip(engine_lib, oracle, tmp_path):
    """Export of the resident packed database, a trip through a file, import into a fresh engine:
    identical fragments (codes read back), identical hits; damaged snapshots are refused."""
    from thermonucleotideblast_b200 import Assay, Engine
    from thermonucleotideblast_b200.engine import EngineError
    rng = np.random.default_rng(808)
    db = [gen.random_codes(int(rng.integers(20000, 70000)), rng) for _ in range(7)]
    gen.sprinkle_degenerate(db[2], rng, frac=3e-3, n_runs_per_50kb=8)
    gen.sprinkle_degenerate(db[5], rng, frac=1e-3, n_runs_per_50kb=2)
    db.append(gen.random_codes(5, rng))      # shorter than a seed word
    assays = gen.make_assays(rng, db[:7], 5, "taqman", variants=2)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)

    def keys(e, hits):
        return [(h.target_id, h.assay_index) + hit_key(e, h, assays[h.assay_index]) + hit_floats(h) for h in hits]

    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        e.add_targets(db)
        want = keys(e, e.search(to_opts(o)))
        snap = e.export_packed()
    assert len(want) >= 5
    assert int(snap["info"][2]) == len(db) and int(snap["info"][6]) == sum(len(c) for c in db)
    # 0.375 B/base (+ alignment gaps of < 64 bases per fragment) + the sparse non-ACGT list
    assert snap["db2"].nbytes + snap["nmask"].nbytes <= 0.375 * (sum(len(c) for c in db) + 64 * len(db)) + 64
    path = tmp_path / "db.tntpacked.npz"
    np.savez(path, **snap)
    loaded = dict(np.load(path))
    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        e.import_packed(loaded)
        for t, c in enumerate(db):
            assert e.target_codes(t, 0, len(c)).tolist() == c.tolist()
        assert keys(e, e.search(to_opts(o))) == want
        # fragments registered after an import join the set like after any other upload
        extra = db[0].copy()
        assert e.add_target(extra) == len(db)
        more = keys(e, e.search(to_opts(o)))
        assert [k for k in more if k[0] < len(db)] == want
        assert [k[1:] for k in more if k[0] == len(db)] == [k[1:] for k in want if k[0] == 0]
        with pytest.raises(EngineError):
            e.import_packed(loaded)              # engine not empty
        e.clear_targets()
        bad = dict(loaded)
        bad["db2"] = loaded["db2"][:-4]
        bad["info"] = loaded["info"].copy()
        bad["info"][3] -= 4                      # fewer words than the fragment table needs
        with pytest.raises(EngineError):
            e.import_packed(bad)
        # a non-ACGT list that is not strictly ascending, or leaves the base space, is refused as well
        assert loaded["exc_pos"].size >= 2
        for damage in ("order", "range", "code"):
            bad = dict(loaded)
            bad["exc_pos"] = loaded["exc_pos"].copy()
            bad["exc_code"] = loaded["exc_code"].copy()
            if damage == "order":
                bad["exc_pos"][1] = bad["exc_pos"][0]
            elif damage == "range":
                bad["exc_pos"][-1] = int(loaded["info"][5]) + 5
            else:
                bad["exc_code"][0] = 2
            with pytest.raises(EngineError):
                e.import_packed(bad)
            e.clear_targets()


def test_add_fasta_twice_without_a_read_in_between(eng, oracle):
    """Two tnt_engine_add_fasta calls back to back (several FASTA files): the pieces of the first call
    are still queued as device-to-staging copies from the transient code buffer when the second call
    starts, and must be issued before that buffer is reused."""
    rng = np.random.default_rng(1331)
    texts = [gen.rand_fasta(rng, n_records=5, max_len=60000, width=80, iupac=0.005) for _ in range(3)]
    texts.append(gen.rand_fasta(rng, n_records=2, max_len=900000, width=60))   # grows the code buffer: reserve() path
    eng.clear_targets()
    tables = [eng.add_fasta(t, fragment_threshold=20000, overlap=500) for t in texts]   # no read-back in between
    for text, (recs, frags) in zip(texts, tables):
        want = oracle.fasta_records(text, threshold=20000, overlap=500)
        assert len(recs) == len(want)
        for r, (off, alen, defline, codes, pieces) in zip(recs, want):
            mine = frags[r.first_fragment:r.first_fragment + r.n_fragments]
            assert [(f.start, f.stop) for f in mine] == [(p[0], p[1]) for p in pieces]
            for f, (s0, s1, pc) in zip(mine, pieces):
                if len(pc):
                    assert eng.target_codes(f.target_id, 0, f.len).tolist() == np.asarray(pc).tolist()


def test_sharded_fasta_ingest_equals_whole(eng):
    """A FASTA text split with sharding.shard_fasta_text (what one process per GPU would ingest):
    the ranges, ingested one after the other, give the records, pieces and codes of the whole text."""
    from thermonucleotideblast_b200.sharding import shard_fasta_text
    rng = np.random.default_rng(2718)
    text = gen.rand_fasta(rng, n_records=23, max_len=30000, width=70, iupac=0.01)

    def ingest(ranges):
        eng.clear_targets()
        out = []
        for a, b in ranges:
            recs, frags = eng.add_fasta(text[a:b], fragment_threshold=8000, overlap=300)
            for r in recs:
                for f in frags[r.first_fragment:r.first_fragment + r.n_fragments]:
                    codes = eng.target_codes(f.target_id, 0, f.len).tolist() if f.len else []
                    out.append((r.text_offset + a, r.text_bytes, r.bases, f.start, f.stop, f.max_stop, codes))
        return out

    whole = ingest([(0, len(text))])
    assert len({w[0] for w in whole}) == 23
    for world in (2, 4, 8):
        assert ingest(shard_fasta_text(text, world)) == whole
