"""Host logic of the multi-GPU path: fragmenting like the reference driver, contiguous shards,
and a world_size-2 gloo run in which each rank searches its shard and rank 0 gathers the hits.
The per-rank search stands in through the oracle here (no GPU in this container); the sharding,
gathering and ordering code is the same one bench.py / the engine wrapper use."""
import os
import socket
import sys

import numpy as np
import pytest

import gen
import harness as H
from thermonucleotideblast_b200.sharding import fragment_record, shard_targets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fragment_record_follows_reference_rule():
    # seq_len_increment (sequence_data.cpp:739-754) + tntblast_local.cpp:282-289,448-468
    assert fragment_record(10, 500000) == [(0, 9)]
    assert fragment_record(500000, 500000) == [(0, 499999)]
    fr = fragment_record(5_000_000, 500000)
    assert len(fr) == 10 and fr[0] == (0, 500000) and fr[1][0] == 500001 and fr[-1][1] == 4_999_999
    fr = fragment_record(1_000_001, 500000)   # needs 3 pieces
    assert len(fr) == 3 and fr[-1][1] == 1_000_000
    for n in (1, 7, 499_999, 500_001, 1_234_567, 12_345_678):
        fr = fragment_record(n, 500000)
        assert fr[0][0] == 0 and fr[-1][1] == n - 1
        for (a, b), (c, d) in zip(fr, fr[1:]):
            assert c == b + 1 and b >= a
    assert fragment_record(0) == []


def test_shard_targets_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    for n, w in [(1, 1), (7, 2), (100, 8), (3, 8), (2000, 4)]:
        lengths = [int(x) for x in rng.integers(1000, 600000, size=n)]
        sh = shard_targets(lengths, w)
        assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == n
        for (a, b), (c, d) in zip(sh, sh[1:]):
            assert b == c and a <= b
        if n >= 4 * w:
            sizes = [sum(lengths[a:b]) for a, b in sh]
            assert max(sizes) <= 1.5 * (sum(lengths) / w) + max(lengths)
    with pytest.raises(ValueError):
        shard_targets([1, 2], 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    db = [gen.random_codes(int(rng.integers(8000, 20000)), rng) for _ in range(9)]
    assays = gen.make_assays(rng, db, 3, "pcr", variants=2)
    o = H.default_options(min_primer_tm=40.0)
    lo, hi = shard_targets([len(c) for c in db], world)[rank]
    mine = []
    for t in range(lo, hi):
        for ai, (F, R, P) in enumerate(assays):
            for h in H.oracle().search(db[t], F, R, P, o):
                mine.append((t, ai, h.amp_first, h.amp_last, h.forward_align.decode()))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = sorted(x for part in gathered for x in part)
        single = []
        for t in range(len(db)):
            for ai, (F, R, P) in enumerate(assays):
                for h in H.oracle().search(db[t], F, R, P, o):
                    single.append((t, ai, h.amp_first, h.amp_last, h.forward_align.decode()))
        q.put((merged == sorted(single), len(single)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_shards_cover_the_database():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n >= 3


def test_fasta_text_shards_hold_whole_records(oracle):
    """Sharding a FASTA text by byte ranges: every range starts where the reference's indexer is in
    its start state, so the records of the ranges, concatenated, are the records of the text."""
    import gen
    from thermonucleotideblast_b200.sharding import shard_fasta_text
    rng = np.random.default_rng(8)
    texts = list(gen.FASTA_EDGE_CASES) + [gen.rand_fasta(rng, n_records=int(rng.integers(1, 12)), max_len=2000,
                                                         width=[60, 0, 7][i % 3], crlf=bool(i % 2)) for i in range(9)]
    for text in texts:
        whole = oracle.fasta_records(text, threshold=300, overlap=20)
        for world in (1, 2, 3, 8):
            ranges = shard_fasta_text(text, world)
            assert len(ranges) == world and ranges[-1][1] == len(text)
            assert all(a <= b for a, b in ranges) and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            parts = []
            for a, b in ranges:
                for off, alen, defline, codes, pieces in oracle.fasta_records(text[a:b], threshold=300, overlap=20):
                    parts.append((off + a, alen, defline, codes.tolist(), [(p[0], p[1], p[2].tolist()) for p in pieces]))
            want = [(off, alen, d, c.tolist(), [(p[0], p[1], p[2].tolist()) for p in pcs]) for off, alen, d, c, pcs in whole]
            assert parts == want
            # every buffer-like input type gives the same cuts
            for other in (bytearray(text), memoryview(text), np.frombuffer(text, dtype=np.uint8), text.decode("latin-1")):
                assert shard_fasta_text(other, world) == ranges


def _fasta_worker(rank, world, port, q):
    """Each rank reads its byte range of one FASTA text (the reader stands in through the oracle
    here: no GPU in this container), searches its records and rank 0 compares the gathered hits
    with a single-process run over the whole text."""
    import torch.distributed as dist
    from thermonucleotideblast_b200.sharding import shard_fasta_text
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(23)
    db = [gen.random_codes(int(rng.integers(6000, 15000)), rng) for _ in range(7)]
    assays = gen.make_assays(rng, db, 3, "pcr", variants=2)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = b"".join(b">r%d\n" % i + b"\n".join(letters[c][k:k + 60].tobytes() for k in range(0, len(c), 60)) + b"\n"
                    for i, c in enumerate(db))
    o = H.default_options(min_primer_tm=40.0)

    def search(byte_range):
        a, b = byte_range
        out = []
        for off, alen, defline, codes, _ in H.oracle().fasta_records(text[a:b]):
            for ai, (F, R, P) in enumerate(assays):
                for h in H.oracle().search(codes, F, R, P, o):
                    out.append((defline, ai, h.amp_first, h.amp_last, h.forward_align.decode()))
        return out

    mine = search(shard_fasta_text(text, world)[rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = [x for part in gathered for x in part]   # ranges are ordered: no sort needed
        single = search((0, len(text)))
        q.put((merged == single, len(single)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_fasta_shards():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fasta_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n >= 3
