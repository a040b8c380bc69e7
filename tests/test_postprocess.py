"""tnt_finalize_hits (host side of the library, no GPU): the reference driver's post-processing of
hit lists -- truncation filter, record coordinates, select_best_match, uniquify_results, sort --
against the reference's own functions (oracle/_ref/libtntref.so: ref_finalize) and against committed
vectors made from them (tests/golden/postprocess.json, tests/golden/make_postprocess.py).

The hit lists come from the oracle searching every fragment of cut records (with the driver's right
overlap), i.e. they hold what uniquify_results exists for: matches reported twice in an overlap,
matches truncated by a cut, nested amplicons."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import gen
import harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postprocess.json")


def make_case(seed, kind, n_assays=6, threshold=6000, overlap=1502, max_len=1500):
    """Records cut like the driver cuts them, oracle hits per (fragment, assay) -> engine-style blocks."""
    from thermonucleotideblast_b200.sharding import fragment_record
    rng = np.random.default_rng(seed)
    records = [gen.random_codes(int(rng.integers(30000, 90000)), rng) for _ in range(3)]
    assays = gen.make_assays(rng, records, n_assays, kind, variants=4, amp=(80, 900))
    # a few sites right on the cuts so that overlaps and truncations occur
    for r, rec in enumerate(records):
        for (a, b) in fragment_record(len(rec), threshold)[:-1]:
            F, R, P = assays[int(rng.integers(0, n_assays))]
            if kind == "probe":
                text = gen.revcomp(P)
            else:
                mid = (gen.rand_oligo(10, rng) + P) if P else ""
                text = F + mid + gen.rand_oligo(int(rng.integers(60, 500)), rng) + gen.revcomp(R)
            gen.plant(rec, max(0, b - int(rng.integers(0, len(text) + 200))), text)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0, max_len=max_len,
                          assay_format=H.ASSAY_PROBE if kind == "probe" else H.ASSAY_PCR)
    orc = H.oracle()
    frags, hits = [], []   # fragment table; per fragment a list of (assay index, oracle Hit)
    for r, rec in enumerate(records):
        for (a, b) in fragment_record(len(rec), threshold):
            piece = rec[a:min(len(rec), b + 1 + overlap)]
            frags.append((r, a, b, len(rec) - 1, len(piece)))
            per = []
            for i, (F, R, P) in enumerate(assays):
                for h in orc.search(piece, F, R, P, o):
                    per.append((i, h))
            hits.append(per)
    return records, assays, frags, hits


def to_block(assays, frags, hits):
    """Oracle hits -> (tnt_hit bytes, n, arena bytes, fragment table) like one engine hands them out."""
    from thermonucleotideblast_b200.engine import BoundOligo, CHit
    arena = bytearray(b"\0")
    recs = []

    def intern(sv):
        if not sv:
            return 0
        off = len(arena)
        arena.extend(sv + b"\0")
        return off

    for t, per in enumerate(hits):
        for i, h in per:
            F, R, P = assays[i]
            c = CHit()
            c.assay_index, c.assay_id, c.target_id = i, 100 + i // 2, t     # two assays share one id (degenerate siblings)
            c.primer_strand, c.probe_strand = h.primer_strand, h.probe_strand
            c.amp_first, c.amp_last, c.probe_first, c.probe_last = h.amp_first, h.amp_last, h.probe_first, h.probe_last
            prim = bool(h.forward_oligo)
            def slot(which, tm, dH, dS, mm, gap, text, name):
                b = BoundOligo()
                b.oligo = -1 if not name else (0 if name == (F or "").encode() else (1 if name == (R or "").encode() else 2))
                b.tm, b.dH, b.dS, b.num_mm, b.num_gap = tm, dH, dS, mm, gap
                b.align_off = intern(text)
                return b
            c.forward = slot(0, h.forward_tm, h.forward_dH, h.forward_dS, h.forward_mm, h.forward_gap, h.forward_align, h.forward_oligo if prim else b"")
            c.reverse = slot(1, h.reverse_tm, h.reverse_dH, h.reverse_dS, h.reverse_mm, h.reverse_gap, h.reverse_align, h.reverse_oligo if prim else b"")
            has_probe = bool(h.probe_align)
            c.probe = slot(2, h.probe_tm, h.probe_dH, h.probe_dS, h.probe_mm, h.probe_gap, h.probe_align, (P or "").encode() if has_probe else b"")
            if has_probe:
                c.probe.oligo = 2
            c.forward_clamp, c.reverse_clamp = h.forward_clamp, h.reverse_clamp
            recs.append(c)
    raw = b"".join(bytes(r) for r in recs)
    return raw, len(recs), bytes(arena), frags


def reference_order(ref, assays, frags, hits, best_match, uniquify):
    """The same post-processing through the reference's own functions: truncation + offsets + list order
    as the driver produces them (tntblast_local.cpp:635-654, :701-706), then ref_finalize per assay id."""
    lists = {}
    index = 0
    for t, per in enumerate(hits):
        rec, start, stop, max_stop, flen = frags[t]
        by_assay = {}
        for i, h in per:
            by_assay.setdefault(i, []).append((index, i, h))
            index += 1
        for i in sorted(by_assay):
            local = []
            for idx, _, h in by_assay[i]:
                prim = bool(h.forward_oligo)
                first, last = (h.amp_first, h.amp_last) if prim else (h.probe_first, h.probe_last)
                if start != 0 and first <= 0:
                    continue
                if stop != max_stop and last >= flen - 1:
                    continue
                local.append((idx, i, h, rec, start))
            aid = 100 + i // 2
            lists[aid] = local + lists.get(aid, [])
    out = []
    for aid in sorted(lists):
        ph, src = [], []
        for idx, i, h, rec, start in lists[aid]:
            prim = bool(h.forward_oligo)
            has_probe = bool(h.probe_align)
            ph.append(H.PostHit(aid, i, rec, int(prim), int(has_probe),
                                h.amp_first + (start if prim else 0), h.amp_last + (start if prim else 0),
                                h.probe_first + (start if has_probe else 0), h.probe_last + (start if has_probe else 0),
                                h.forward_tm, h.reverse_tm, h.probe_tm, len(h.forward_oligo), len(h.reverse_oligo),
                                h.forward_align or None, h.reverse_align or None, h.probe_align or None))
            src.append(idx)
        out += [src[k] for k in ref.finalize(ph, best_match, uniquify)]
    return out


CASES = [(1, "taqman"), (2, "pcr"), (3, "probe"), (4, "pcr")]


def product_order(assays, frags, hits, best_match, uniquify):
    from thermonucleotideblast_b200 import Assay
    from thermonucleotideblast_b200.engine import finalize_hits
    raw, n, arena, fr = to_block(assays, frags, hits)
    alist = [Assay(100 + i // 2, *a) for i, a in enumerate(assays)]
    res = finalize_hits([(raw, n, arena, fr)], alist, best_match=best_match, uniquify=uniquify)
    return [idx for (_, idx, _) in res], res


def test_finalize_matches_the_reference_functions(engine_lib, ref):
    total = dropped = merged = 0
    for seed, kind in CASES:
        records, assays, frags, hits = make_case(seed, kind)
        n_in = sum(len(p) for p in hits)
        for best, uniq in ((False, True), (True, True), (False, False), (True, False)):
            got, res = product_order(assays, frags, hits, best, uniq)
            want = reference_order(ref, assays, frags, hits, best, uniq)
            assert got == want, (seed, kind, best, uniq)
            total += len(want)
            dropped += n_in - len(want)
            if uniq and not best:
                merged += len(reference_order(ref, assays, frags, hits, False, False)) - len(want)
        # record coordinates and record index
        got, res = product_order(assays, frags, hits, False, True)
        flat = [(t, i, h) for t, per in enumerate(hits) for i, h in per]
        for (_, idx, c) in res:
            t, i, h = flat[idx]
            rec, start = frags[t][0], frags[t][1]
            assert c.target_id == rec
            if h.forward_oligo:
                assert (c.amp_first, c.amp_last) == (h.amp_first + start, h.amp_last + start)
            if h.probe_align:
                assert (c.probe_first, c.probe_last) == (h.probe_first + start, h.probe_last + start)
    assert total > 100 and dropped > 20 and merged >= 15   # uniquify_results had real work


def test_finalize_golden_vectors(engine_lib):
    """The same comparison against committed answers (made from the compiled reference)."""
    gold = json.load(open(GOLD))
    assert len(gold) == len(CASES) * 4
    k = 0
    for seed, kind in CASES:
        records, assays, frags, hits = make_case(seed, kind)
        for best, uniq in ((False, True), (True, True), (False, False), (True, False)):
            got, _ = product_order(assays, frags, hits, best, uniq)
            g = gold[k]
            assert (g["seed"], g["kind"], g["best_match"], g["uniquify"]) == (seed, kind, best, uniq)
            assert got == g["order"], (seed, kind, best, uniq)
            k += 1


def test_finalize_several_blocks_equal_one(engine_lib):
    """Hits of several engines (shards of the fragment list) give what one engine holding everything gives."""
    from thermonucleotideblast_b200 import Assay
    from thermonucleotideblast_b200.engine import finalize_hits
    records, assays, frags, hits = make_case(2, "pcr")
    alist = [Assay(100 + i // 2, *a) for i, a in enumerate(assays)]
    one = finalize_hits([to_block(assays, frags, hits)], alist)
    key = lambda c: (c.assay_index, c.target_id, c.amp_first, c.amp_last, c.probe_first, c.probe_last, c.forward.tm, c.reverse.tm)
    for world in (2, 3):
        cuts = [len(frags) * r // world for r in range(world + 1)]
        blocks = [to_block(assays, frags[cuts[r]:cuts[r + 1]], hits[cuts[r]:cuts[r + 1]]) for r in range(world)]
        many = finalize_hits(blocks, alist)
        assert [key(c) for (_, _, c) in many] == [key(c) for (_, _, c) in one]
    assert len(one) > 10
