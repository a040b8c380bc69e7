"""Parity of the CUDA engine (through the C ABI) with the oracle, stage by stage and end to end.

Bar (BASELINE.json north_star): seeds, hit coordinates, strands, counts and alignment strings
bit-exact; Tm within 0.01 C and dG within 0.001 kcal/mol.  The engine evaluates dH/dS/Tm with the
reference's rounding, so the tests additionally demand bit-identical floats and only fall back to
the stated tolerance in the assertion message.
"""
import numpy as np
import pytest

import gen
import harness as H

pytestmark = pytest.mark.gpu

TM_TOL = 0.01
DG_TOL = 0.001


@pytest.fixture(scope="module")
def eng(engine_lib):
    from thermonucleotideblast_b200 import Engine
    e = Engine()
    yield e
    e.close()


def hit_key(engine, h, assay):
    """Same tuple layout as harness.Hit.exact_key()."""
    F, R, P = assay
    names = {0: F, 1: R, 2: P, -1: ""}
    seq = engine.hit_sequence(h)
    fnv = 1469598103934665603
    for ch in seq.encode():
        fnv = ((fnv ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h.primer_strand, h.probe_strand, h.amp_first, h.amp_last, h.probe_first, h.probe_last,
            h.forward.num_mm, h.forward.num_gap, h.reverse.num_mm, h.reverse.num_gap,
            h.probe.num_mm, h.probe.num_gap, h.forward_clamp, h.reverse_clamp,
            len(seq), fnv,
            (names[h.forward.oligo] or "").encode(), (names[h.reverse.oligo] or "").encode(),
            h.forward_align.encode(), h.reverse_align.encode(), h.probe_align.encode(),
            seq[:255].encode())


def hit_floats(h):
    return (h.forward.tm, h.forward.dH, h.forward.dS, h.reverse.tm, h.reverse.dH, h.reverse.dS,
            h.probe.tm, h.probe.dH, h.probe.dS)


def assert_hits_equal(engine, got, want, assay, T=310.15):
    gk = [hit_key(engine, h, assay) for h in got]
    wk = [h.exact_key() for h in want]
    assert len(gk) == len(wk), f"hit count {len(gk)} != {len(wk)}"
    # the reference emits hits of one (fragment, assay) in join order; compare as ordered lists
    assert gk == wk
    for g, w in zip(got, want):
        gf, wf = hit_floats(g), w.floats()
        for k in (0, 3, 6):
            assert abs(gf[k] - wf[k]) <= TM_TOL, "Tm outside 0.01 C"
            dg_g = gf[k + 1] - T * gf[k + 2]
            dg_w = wf[k + 1] - T * wf[k + 2]
            assert abs(dg_g - dg_w) <= DG_TOL, "dG outside 0.001 kcal/mol"
        assert gf == wf, "floats are expected to be bit-identical"


def to_opts(o):
    """harness.Options -> engine SearchOptions (same field names)."""
    from thermonucleotideblast_b200 import search_options
    s = search_options()
    for name, _ in s._fields_:
        setattr(s, name, getattr(o, name))
    return s


# ---------------------------------------------------------------------------------------------
def test_seeds_bit_exact(eng, oracle):
    rng = np.random.default_rng(101)
    eng.clear_targets()
    frags = []
    for k in range(4):
        n = int(rng.integers(20000, 70000))
        codes = gen.random_codes(n, rng)
        if k % 2:
            gen.sprinkle_degenerate(codes, rng, frac=5e-3, n_runs_per_50kb=20)
            codes[rng.integers(0, n, size=20)] = 16
            codes[rng.integers(0, n, size=20)] = 17
        frags.append(codes)
    oligos = [gen.rand_oligo(20, rng), gen.rand_oligo(30, rng), gen.rand_oligo(7, rng), gen.rand_oligo(56, rng)]
    o = list(gen.rand_oligo(24, rng)); o[9] = "I"; oligos.append("".join(o))
    o = list(gen.rand_oligo(26, rng)); o[3] = "N"; o[20] = "R"; oligos.append("".join(o))
    oligos.append("ACGTAC")          # shorter than the word size: no seeds
    oligos.append("AAAAAAAAAAAAAAAAAAAA")
    # plant exact + shifted copies so several words share a diagonal
    for ol in oligos[:2]:
        gen.plant(frags[0], 1000, ol)
        gen.plant(frags[0], 5000, gen.revcomp(ol))
    eng.clear_targets()
    ids = [eng.add_target(c) for c in frags]
    total = 0
    for tid, codes in zip(ids, frags):
        for ol in oligos:
            for plus in (False, True):
                want = oracle.seeds(codes, ol, 7, plus, unique=True)
                got = eng.seeds(tid, ol, plus)
                assert got == want, (tid, ol, plus)
                total += len(want)
    assert total > 500


def test_seeds_fragment_edges(eng, oracle):
    """Tiny fragments, fragments shorter than a word, a seed on the very last position."""
    rng = np.random.default_rng(7)
    eng.clear_targets()
    ol = gen.rand_oligo(20, rng)
    frags = [gen.str_to_codes(ol[:5]), gen.str_to_codes(ol[:7]), gen.str_to_codes("ACGT" + ol),
             gen.str_to_codes(gen.revcomp(ol)), gen.random_codes(8192, rng), gen.random_codes(8193, rng),
             gen.random_codes(8191 + 7, rng)]
    gen.plant(frags[4], 8192 - 20, ol)
    gen.plant(frags[5], 8193 - 7, ol[:7])
    ids = [eng.add_target(c) for c in frags]
    for tid, codes in zip(ids, frags):
        for plus in (False, True):
            assert eng.seeds(tid, ol, plus) == oracle.seeds(codes, ol, 7, plus, unique=True)


def _compare_align(eng, oracle, tid, codes, ol, plus, seeds, ct=9.0e-7, T=310.15, na=0.05, d5=0, d3=0):
    got = eng.align(tid, ol, plus, seeds, ct=ct)
    nvalid = 0
    for (q, t), g in zip(seeds, got):
        w = oracle.bind_window(codes, ol, plus, q, t, T=T, na=na, ct=ct, dangle5=d5, dangle3=d3)
        assert g.valid == w.valid, (ol, plus, q, t)
        assert (g.target_start, g.target_stop) == (w.target_start, w.target_stop)
        if not w.valid:
            continue
        nvalid += 1
        assert abs(g.tm - w.tm) <= TM_TOL and abs(g.dG - w.dG) <= DG_TOL, (ol, plus, q, t, g.tm, w.tm)
        assert (g.tm, g.dH, g.dS, g.dG) == (w.tm, w.dH, w.dS, w.dG), "floats expected bit-identical"
        assert (g.anchor5, g.anchor3, g.num_mismatch, g.num_gap, g.max_poly_degen) == \
            (w.anchor5, w.anchor3, w.num_mismatch, w.num_gap, w.max_poly_degen), (ol, plus, q, t)
        assert (g.q_first, g.q_last, g.t_first, g.t_last) == (w.q_first, w.q_last, w.t_first, w.t_last)
        assert (g.loc_5, g.loc_3) == (w.loc_5, w.loc_3)
        assert g.alignment == w.alignment, (ol, plus, q, t, g.alignment, w.alignment)
    return nvalid


def test_align_random_and_planted_windows(eng, oracle):
    rng = np.random.default_rng(202)
    n = 120000
    codes = gen.random_codes(n, rng)
    oligos = [gen.rand_oligo(L, rng) for L in (12, 18, 20, 22, 25, 30, 35, 56)]
    # planted sites with 0..4 edits incl. indels, on both strands, some at the fragment edges
    for i, ol in enumerate(oligos):
        for k in range(12):
            text = gen.mutate(gen.revcomp(ol) if k % 2 else ol, k % 5, rng)
            gen.plant(codes, 2000 + (i * 12 + k) * 400, text)
    gen.plant(codes, 0, gen.revcomp(oligos[2])[3:])
    gen.plant(codes, n - 15, oligos[3][:15])
    eng.clear_targets()
    tid = eng.add_target(codes)
    nvalid = 0
    for ol in oligos:
        for plus in (False, True):
            seeds = oracle.seeds(codes, ol, 7, plus, unique=True)
            nvalid += _compare_align(eng, oracle, tid, codes, ol, plus, seeds)
    assert nvalid > 1000


def test_align_degenerate_target_and_oligo(eng, oracle):
    rng = np.random.default_rng(303)
    n = 60000
    codes = gen.random_codes(n, rng)
    oligos = []
    for L in (20, 24, 30):
        o = list(gen.rand_oligo(L, rng))
        o[int(rng.integers(7, L - 7))] = "I"
        o[int(rng.integers(0, L))] = "RYMKSWN"[int(rng.integers(0, 7))]
        oligos.append("".join(o))
    for i, ol in enumerate(oligos):
        plain = ol.replace("I", "A").replace("R", "A").replace("Y", "C").replace("M", "A").replace("K", "G") \
            .replace("S", "C").replace("W", "A").replace("N", "T")
        for k in range(10):
            gen.plant(codes, 1000 + (i * 10 + k) * 500, gen.mutate(gen.revcomp(plain) if k % 2 else plain, k % 4, rng))
    gen.sprinkle_degenerate(codes, rng, frac=2e-2, n_runs_per_50kb=40)
    codes[rng.integers(0, n, size=200)] = 16   # DB_GAP: dropped from windows
    codes[rng.integers(0, n, size=200)] = 17   # DB_UNKNOWN
    codes[rng.integers(0, n, size=200)] = 4    # inosine in the target
    eng.clear_targets()
    tid = eng.add_target(codes)
    nvalid = 0
    for ol in oligos:
        for plus in (False, True):
            seeds = oracle.seeds(codes, ol, 7, plus, unique=True)
            nvalid += _compare_align(eng, oracle, tid, codes, ol, plus, seeds, ct=2.5e-7)
    assert nvalid > 200


def test_readme_known_answer(eng):
    """README.md:138,162-203 of the reference: gibb-marburg TaqMan assay."""
    amp = ("TTCCCCTTTGGAGGCATCCAAGCGATGGGCTTTCAGGACAGGTGTACCTCCCAAGAATGTTGAGTATACAGAAGGGGAGGAAGCCAAAACATGCTACAATATAAG"
           "TGTAACGGATCCCTCTGGAAAATCCTTGCTGTTGGATCCTCC")
    rng = np.random.default_rng(1)
    codes = gen.random_codes(6121 + len(amp) + 3000, rng)
    gen.plant(codes, 6121, amp)
    from thermonucleotideblast_b200 import Assay, search_options
    eng.clear_targets()
    eng.add_target(codes)
    eng.set_assays([Assay(0, "TTCCCCTTTGGAGGCATC", "GGAGGATCCAACAGCAAGG", "CGATGGGCTTTCAGGACAGGTGT")])
    hits = eng.search(search_options(min_primer_tm=40.0, min_probe_tm=45.0))
    assert len(hits) == 1
    h = hits[0]
    assert (h.amp_first, h.amp_last, h.probe_first, h.probe_last) == (6121, 6267, 6143, 6165)
    T = 310.15
    assert abs((h.forward.dH - T * h.forward.dS) - (-16.8574)) < 1e-3 and abs(h.forward.dH + 135.5) < 1e-3
    assert abs((h.reverse.dH - T * h.reverse.dS) - (-17.8955)) < 1e-3 and abs(h.reverse.dH + 146.5) < 1e-3
    assert abs((h.probe.dH - T * h.probe.dS) - (-22.9778)) < 1e-3 and abs(h.probe.dH + 180.2) < 1e-3
    assert (h.forward_clamp, h.reverse_clamp) == (18, 19)
    assert h.forward_align == "5' TTCCCCTTTGGAGGCATC 3'\n   ||||||||||||||||||\n3' AAGGGGAAACCTCCGTAG 5'"
    assert h.probe_align == "5' CGATGGGCTTTCAGGACAGGTGT 3'\n   |||||||||||||||||||||||\n3' GCTACCCGAAAGTCCTGTCCACA 5'"
    # the probe text occurs verbatim in the plus strand, i.e. the oligo binds the minus strand
    assert h.probe_strand == 1 and h.primer_strand == 0
    assert eng.hit_sequence(h) == amp


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_search_pcr_and_taqman(eng, oracle, seed):
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(1000 + seed)
    total = 0
    for it in range(8):
        probe = bool(it % 2)
        codes, F, R, P = gen.make_pcr_case(rng, int(rng.integers(20000, 90000)), n_sites=int(rng.integers(1, 6)), probe=probe)
        if it % 4 == 3:
            gen.sprinkle_degenerate(codes, rng, frac=2e-3, n_runs_per_50kb=5)
        o = H.default_options(min_primer_tm=float(rng.choice([35.0, 40.0, 45.0])), min_probe_tm=40.0,
                              max_len=int(rng.choice([500, 2000])), single_primer_pcr=int(rng.integers(0, 2)),
                              primer_clamp=int(rng.integers(0, 3)),
                              min_max_primer_clamp=int(rng.choice([-1, -1, 2])))
        want = oracle.search(codes, F, R, P, o)
        eng.clear_targets()
        eng.add_target(codes)
        eng.set_assays([Assay(7, F, R, P)])
        got = eng.search(to_opts(o))
        assert_hits_equal(eng, got, want, (F, R, P))
        total += len(want)
    assert total > 5


def test_search_probe_and_padlock(eng, oracle):
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(77)
    total = 0
    for it in range(10):
        n = int(rng.integers(20000, 60000))
        db = [gen.random_codes(n, rng)]
        kind = "probe" if it % 2 else "padlock"
        (F, R, P), = gen.make_assays(rng, db, 1, kind, variants=4)
        codes = db[0]
        if kind == "probe":
            o = H.default_options(assay_format=H.ASSAY_PROBE, min_probe_tm=float(rng.choice([30.0, 45.0])),
                                  target_strand=int(rng.choice([1, 2, 3])), probe_clamp_5=int(rng.integers(0, 2)))
        else:
            o = H.default_options(assay_format=int(rng.choice([H.ASSAY_PADLOCK, H.ASSAY_MIPS])), min_probe_tm=30.0,
                                  max_len=int(rng.choice([0, 3, 50])), probe_clamp_3=int(rng.integers(0, 3)))
        want = oracle.search(codes, F, R, P, o)
        eng.clear_targets()
        eng.add_target(codes)
        eng.set_assays([Assay(3, F, R, P)])
        got = eng.search(to_opts(o))
        assert_hits_equal(eng, got, want, (F, R, P))
        assert eng.hit_sequences() == [eng.hit_sequence(h) for h in got]
        total += len(want)
    assert total > 5


def test_search_multi_fragment_multi_assay(eng, oracle):
    """Batch semantics: many fragments x many assays in one call == the reference's nested loops."""
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(4242)
    db = [gen.random_codes(int(rng.integers(15000, 40000)), rng) for _ in range(6)]
    assays = gen.make_assays(rng, db, 8, "taqman", variants=3)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    eng.clear_targets()
    for c in db:
        eng.add_target(c)
    eng.set_assays([Assay(100 + i, *a) for i, a in enumerate(assays)])
    got = eng.search(to_opts(o))
    st = eng.stats()
    assert st.alignments > 0 and st.dp_cells > 0 and st.kernel_launches >= 2
    # the text of all hits in one call == the per-hit calls (which assert_hits_equal checks against the oracle)
    assert eng.hit_sequences() == [eng.hit_sequence(h) for h in got] and len(got) >= 8
    k = 0
    total = 0
    for t, codes in enumerate(db):
        for i, a in enumerate(assays):
            want = oracle.search(codes, a[0], a[1], a[2], o)
            mine = [h for h in got if h.target_id == t and h.assay_index == i]
            assert all(h.assay_id == 100 + i for h in mine)
            assert_hits_equal(eng, mine, want, a)
            total += len(want)
            k += len(mine)
    assert k == len(got) and total >= 8


def test_engine_matches_compiled_reference(eng, ref):
    """Same comparison against the unmodified reference build (only where oracle/_ref travelled)."""
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(99)
    for it in range(4):
        codes, F, R, P = gen.make_pcr_case(rng, 50000, n_sites=4, probe=bool(it % 2))
        o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
        want = ref.search(codes, F, R, P, o)
        eng.clear_targets()
        eng.add_target(codes)
        eng.set_assays([Assay(0, F, R, P)])
        assert_hits_equal(eng, eng.search(to_opts(o)), want, (F, R, P))


def test_refuses_unsupported(engine_lib):
    from thermonucleotideblast_b200 import Assay, Engine, EngineError, search_options
    with pytest.raises(EngineError):
        Engine(word_size=12)
    e = Engine()
    try:
        e.add_target(gen.random_codes(1000, np.random.default_rng(0)))
        e.set_assays([Assay(0, "A" * 60, "ACGTACGTACGTACGTAC", None)])
        with pytest.raises(EngineError):
            e.search(search_options(min_primer_tm=40.0))
        with pytest.raises(EngineError):
            e.set_assays([Assay(0, "ACGTACGTACGTACGTAC", None, None)])
    finally:
        e.close()


def test_bounds_that_accept_everything(eng, oracle):
    """The bounds of tntblast.h:31-38 (Tm in [0, 9999], dG in [-9999, 0]) -- what a user gets who only
    gives e.g. `-x 70`: every seed window with an alignment is a bound site.  Windows without any
    alignment (Tm = 0, dG = 0 in the reference, reported there with the coordinates of whatever the
    thread aligned before) are dropped and counted; here every window holds its exact seed word, so
    none occurs and the hit list equals the oracle's."""
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(515)
    total = 0
    for probe in (False, True):
        codes, F, R, P = gen.make_pcr_case(rng, 30000, n_sites=3, probe=probe)
        o = H.default_options(max_len=300)
        assert o.min_primer_tm == 0.0 and o.max_primer_dg == 0.0
        want = oracle.search(codes, F, R, P, o)
        eng.clear_targets()
        eng.add_target(codes)
        eng.set_assays([Assay(0, F, R, P)])
        got = eng.search(to_opts(o))
        assert eng.stats().nonbinding_dropped == 0
        assert_hits_equal(eng, got, want, (F, R, P))
        total += len(want)
    assert total >= 6


def test_cull_anomaly_is_reproduced(engine_lib, oracle, monkeypatch):
    """SURVEY 8a row C1.  cull_oligo_match (amplicon_search.cpp:679-765) sorts bound sites by loc_5 and
    unbound seeds by seed position in one list and ends its partner scan on an unsigned seed
    distance; when a primer binds both strands of a palindromic site the two bound sites overlap
    and the reference loses the amplicon that the later one opens.  Two such neighbourhoods from the
    config-5 data (tests/golden/cull_cases.json, answers recorded from the compiled reference): the
    engine reproduces the reference by default, and reports the lost amplicon as well when asked to
    keep culled sites."""
    import json
    import os
    from thermonucleotideblast_b200 import Assay, Engine
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cull_cases.json")))
    for c in cases:
        codes = gen.str_to_codes(c["codes"])
        F, R = c["forward"], c["reverse"]
        o = H.default_options(min_primer_tm=c["min_primer_tm"])
        want = oracle.search(codes, F, R, None, o)
        assert [(h.amp_first, h.amp_last) for h in want] == [(h["amp_first"], h["amp_last"]) for h in c["reference_hits"]]
        lost = tuple(c["lost_amplicon"])
        with Engine() as e:
            e.add_target(codes)
            e.set_assays([Assay(0, F, R, None)])
            got = e.search(to_opts(o))
            assert e.stats().replayed_groups == 1
            assert_hits_equal(e, got, want, (F, R, None))
            assert lost not in [(h.amp_first, h.amp_last) for h in got]
        with Engine(keep_culled_sites=True) as e:
            e.add_target(codes)
            e.set_assays([Assay(0, F, R, None)])
            more = e.search(to_opts(o))
            assert e.stats().replayed_groups == 0
            coords = [(h.amp_first, h.amp_last) for h in more]
            assert lost in coords and set((h.amp_first, h.amp_last) for h in got) < set(coords)


def test_replay_of_every_group_changes_nothing(engine_lib, oracle, monkeypatch):
    """The step-by-step replay (every seed of the group aligned, the reference's list operations
    followed literally) applied to *every* group with a hit must give what the selective default
    gives -- and what the oracle gives."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(60221)
    db = [gen.random_codes(int(rng.integers(30000, 60000)), rng) for _ in range(5)]
    gen.sprinkle_degenerate(db[1], rng, frac=1e-3, n_runs_per_50kb=3)
    # nothing here makes the orders of bound sites disagree: no group is replayed unless forced
    assays = gen.make_assays(rng, db, 4, "taqman", variants=3) + gen.make_assays(rng, db, 4, "pcr", variants=3) + \
        gen.make_assays(rng, db, 4, "taqman", variants=0) + gen.make_assays(rng, db, 4, "pcr", variants=0)
    o = H.default_options(min_primer_tm=38.0, min_probe_tm=38.0, max_len=1000)

    def run():
        with Engine() as e:
            e.add_targets(db)
            e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
            hits = e.search(to_opts(o))
            return [(h.target_id, h.assay_index) + hit_key(e, h, assays[h.assay_index]) + hit_floats(h) for h in hits], e.stats().replayed_groups

    base, n_default = run()
    monkeypatch.setenv("TNT_REPLAY_ALL", "1")
    forced, n_forced = run()
    assert forced == base and len(base) >= 12
    assert n_forced >= 12 and n_default <= 2
    k = 0
    for t, codes in enumerate(db):
        for i, a in enumerate(assays):
            want = oracle.search(codes, a[0], a[1], a[2], o)
            mine = [x for x in forced if x[0] == t and x[1] == i]
            assert [x[2:2 + len(w.exact_key())] for x, w in zip(mine, want)] == [w.exact_key() for w in want] and len(mine) == len(want)
            k += len(want)
    assert k == len(forced)


def test_repeats_and_forced_multi_pass(engine_lib, oracle, monkeypatch):
    """Low-complexity / tandem-repeat fragments overflow the seed buckets (estimated for random
    sequence): the engine has to shrink the pass and retry, and the result must not change.  A tiny
    candidate budget additionally forces many passes over the tiles."""
    from thermonucleotideblast_b200 import Assay, Engine
    monkeypatch.setenv("TNT_CAND_BUDGET_MB", "1")
    rng = np.random.default_rng(31337)
    unit = gen.rand_oligo(37, rng)
    F = unit[3:23]
    R = gen.revcomp(unit[10:30])
    frags = []
    for k in range(3):
        codes = gen.random_codes(60000, rng)
        gen.plant(codes, 5000, unit * 400)            # 14.8 kb tandem repeat holding both primer sites
        gen.plant(codes, 30000, "A" * 3000)
        gen.plant(codes, 40000, gen.mutate(unit * 30, 25, rng))
        frags.append(codes)
    o = H.default_options(min_primer_tm=45.0, max_len=300)
    e = Engine()
    try:
        for c in frags:
            e.add_target(c)
        assays = [(F, R, None), ("A" * 20, "T" * 20, None), (gen.rand_oligo(20, rng), gen.rand_oligo(22, rng), None)]
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        got = e.search(to_opts(o))
        total = 0
        for t, codes in enumerate(frags):
            for i, a in enumerate(assays):
                want = oracle.search(codes, a[0], a[1], a[2], o)
                mine = [h for h in got if h.target_id == t and h.assay_index == i]
                assert_hits_equal(e, mine, want, a)
                total += len(want)
        assert total > 100
    finally:
        e.close()


def test_config3_like_degenerate_probes(eng, oracle):
    """BASELINE config 3 in miniature: 30-mer probes with one inosine and one two-fold IUPAC code,
    enumerated to two oligos with degeneracy 2 (expand_degenerate_signatures keeps the inosine,
    degenerate_na.cpp:94-96), against a target with IUPAC codes and N runs, both strands."""
    from thermonucleotideblast_b200 import Assay
    rng = np.random.default_rng(333)
    two_fold = {"R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "CG", "W": "AT"}
    total = 0
    for it in range(4):
        codes = gen.random_codes(80000, rng)
        base = gen.rand_oligo(30, rng)
        code = "RYMKSW"[int(rng.integers(0, 6))]
        pi, pd = sorted(rng.choice(np.arange(8, 22), size=2, replace=False).tolist())
        variants = []
        for alt in two_fold[code]:
            o = list(base)
            o[pi] = "I"
            o[pd] = alt
            variants.append("".join(o))
        for k in range(6):
            site = list(base)
            site[pd] = two_fold[code][k % 2]
            gen.plant(codes, 3000 + 9000 * k, gen.mutate(gen.revcomp("".join(site)) if k % 2 else "".join(site), k % 3, rng))
        gen.sprinkle_degenerate(codes, rng, frac=1e-3, n_runs_per_50kb=1)
        o = H.default_options(assay_format=H.ASSAY_PROBE, min_probe_tm=50.0)
        eng.clear_targets()
        eng.add_target(codes)
        eng.set_assays([Assay(i, None, None, v, probe_degen=2) for i, v in enumerate(variants)])
        got = eng.search(to_opts(o))
        for i, v in enumerate(variants):
            want = oracle.search(codes, None, None, v, o, degen=(1, 1, 2))
            mine = [h for h in got if h.assay_index == i]
            assert_hits_equal(eng, mine, want, (None, None, v))
            total += len(want)
    assert total >= 8


def test_full_size_properties(eng, oracle):
    """Size-independent properties at a larger scale (50 Mbp x 20 TaqMan assays): every exactly
    planted amplicon is reported with zero mismatches, searching twice gives identical hits, and
    the hit set of a fragment does not depend on which other fragments are resident."""
    from thermonucleotideblast_b200 import Assay, search_options
    rng = np.random.default_rng(5150)
    frags = [gen.random_codes(500000, rng) for _ in range(100)]
    assays = []
    planted = []
    for a in range(20):
        F, R, P = gen.rand_oligo(20, rng), gen.rand_oligo(21, rng), gen.rand_oligo(25, rng)
        t = int(rng.integers(0, len(frags)))
        pos = int(rng.integers(1000, 400000))
        text = F + gen.rand_oligo(15, rng) + P + gen.rand_oligo(150, rng) + gen.revcomp(R)
        gen.plant(frags[t], pos, text)
        assays.append((F, R, P))
        planted.append((a, t, pos, pos + len(text) - 1))
    opts = search_options(min_primer_tm=45.0, min_probe_tm=50.0)
    eng.clear_targets()
    for c in frags:
        eng.add_target(c)
    eng.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
    hits1 = eng.search(opts)
    key = lambda h: (h.target_id, h.assay_index, h.amp_first, h.amp_last, h.probe_first, h.probe_last, h.forward_align, h.reverse_align, h.probe_align, h.forward.tm, h.reverse.tm, h.probe.tm)
    o = H.default_options(min_primer_tm=45.0, min_probe_tm=50.0)
    exact = 0
    for (a, t, lo, hi) in planted:
        # the oracle arbitrates (a random 20-mer may melt below the 45 C bound and then has no hit)
        want = oracle.search(frags[t], assays[a][0], assays[a][1], assays[a][2], o)
        mine = [h for h in hits1 if h.assay_index == a and h.target_id == t]
        assert_hits_equal(eng, mine, want, assays[a])
        exact += any(h.amp_first == lo and h.amp_last == hi and h.forward.num_mm == 0 and h.reverse.num_mm == 0
                     and h.probe.num_mm == 0 for h in mine)
    assert exact >= len(planted) // 2
    hits2 = eng.search(opts)
    assert [key(h) for h in hits1] == [key(h) for h in hits2]
    # a subset of the fragments, registered alone, yields the same hits for those fragments
    sub = sorted({t for (_, t, _, _) in planted})[:5]
    eng.clear_targets()
    for t in sub:
        eng.add_target(frags[t])
    hits3 = eng.search(opts)
    for new_id, t in enumerate(sub):
        a = [key(h)[1:] for h in hits1 if h.target_id == t]
        b = [key(h)[1:] for h in hits3 if h.target_id == new_id]
        assert a == b


@pytest.mark.parametrize("mode", ["dense", "dense_global", "sparse"])
def test_seeds_both_scan_kernels(eng, oracle, monkeypatch, mode):
    """k_seed_scan_smem (dense tables; tile, table and packed oligos in shared memory), k_seed_scan (dense
    tables too large for shared memory: table in L2) and k_seed_scan_sparse (grouped pre-filter) give the
    same seeds."""
    monkeypatch.setenv("TNT_SCAN_MODE", mode.split("_")[0])
    if mode == "dense_global":
        monkeypatch.setenv("TNT_SCAN_GLOBAL_TABLE", "1")
    rng = np.random.default_rng(808)
    eng.clear_targets()
    frags = [gen.random_codes(n, rng) for n in (300000, 131072, 131073, 131072 + 70, 64, 7, 200001)]
    gen.sprinkle_degenerate(frags[0], rng, frac=1e-3, n_runs_per_50kb=5)
    ol = gen.rand_oligo(22, rng)
    for f in frags:
        if len(f) > 100:
            gen.plant(f, len(f) - 22, ol)                     # site ending on the last base
            gen.plant(f, 0, gen.revcomp(ol))                  # site on the first base
    gen.plant(frags[0], 131072 - 10, ol)                      # straddles a sparse tile boundary
    gen.plant(frags[0], 8192 - 5, gen.revcomp(ol))            # straddles a dense tile boundary
    ids = [eng.add_target(c) for c in frags]
    total = 0
    for tid, codes in zip(ids, frags):
        for o in (ol, "ACGTACGTAC", gen.rand_oligo(30, rng)):
            for plus in (False, True):
                want = oracle.seeds(codes, o, 7, plus, unique=True)
                assert eng.seeds(tid, o, plus) == want, (mode, tid, o, plus)
                total += len(want)
    assert total > 300


@pytest.mark.parametrize("mode", ["dense", "dense_global", "sparse"])
def test_search_both_scan_kernels(engine_lib, oracle, monkeypatch, mode):
    from thermonucleotideblast_b200 import Assay, Engine
    monkeypatch.setenv("TNT_SCAN_MODE", mode.split("_")[0])
    if mode == "dense_global":
        monkeypatch.setenv("TNT_SCAN_GLOBAL_TABLE", "1")
    rng = np.random.default_rng(909)
    db = [gen.random_codes(int(rng.integers(100000, 300000)), rng) for _ in range(3)]
    assays = gen.make_assays(rng, db, 4, "pcr", variants=3)
    o = H.default_options(min_primer_tm=42.0)
    e = Engine()
    try:
        for c in db:
            e.add_target(c)
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        got = e.search(to_opts(o))
        for t, codes in enumerate(db):
            for i, a in enumerate(assays):
                want = oracle.search(codes, a[0], a[1], a[2], o)
                assert_hits_equal(e, [h for h in got if h.target_id == t and h.assay_index == i], want, a)
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------
# Alignment tiers.  The lean tier (2-bit trace, gapless evaluator) takes most windows, the
# full-trace tier the rest; TNT_NO_LEAN routes every window through the full-trace tier.  All of
# them must reproduce the oracle for other temperatures / salt and with dangling ends enabled
# (virtual bases in the alignment: the lean tier then falls back to the general evaluator).
@pytest.mark.parametrize("T,na,d5,d3,no_lean", [
    (310.15, 0.05, 1, 1, False),
    (310.15, 0.05, 1, 0, False),
    (310.15, 0.05, 0, 1, False),
    (330.15, 0.1, 0, 0, False),
    (295.15, 1.0, 0, 0, False),
    (310.15, 0.05, 0, 0, True),
    (310.15, 0.05, 1, 1, True),
])
def test_align_tiers_and_parameters(engine_lib, oracle, monkeypatch, T, na, d5, d3, no_lean):
    from thermonucleotideblast_b200 import Engine
    if no_lean:
        monkeypatch.setenv("TNT_NO_LEAN", "1")
    rng = np.random.default_rng(int(T * 10) + d5 * 2 + d3 + (7 if no_lean else 0))
    n = 60000
    codes = gen.random_codes(n, rng)
    oligos = [gen.rand_oligo(L, rng) for L in (17, 19, 21, 22, 23, 26, 27, 33, 41, 55)]
    for i, ol in enumerate(oligos):
        for k in range(10):
            text = gen.mutate(gen.revcomp(ol) if k % 2 else ol, k % 5, rng)
            gen.plant(codes, 1500 + (i * 10 + k) * 450, text)
    gen.plant(codes, 0, gen.revcomp(oligos[3])[2:])
    gen.plant(codes, n - 12, oligos[4][:12])
    with Engine(target_T=T, salt=na, dangle5=bool(d5), dangle3=bool(d3)) as e:
        tid = e.add_target(codes)
        nvalid = 0
        for ol in oligos:
            for plus in (False, True):
                seeds = oracle.seeds(codes, ol, 7, plus, unique=True)
                nvalid += _compare_align(e, oracle, tid, codes, ol, plus, seeds, T=T, na=na, d5=d5, d3=d3)
        assert nvalid > 500


def test_search_lean_and_full_tiers_agree(engine_lib, oracle, monkeypatch):
    """The same TaqMan batch search (a) as shipped, (b) without the lean tier, (c) with the lean
    tier but without its "too few columns to reach min Tm" shortcut: all equal the oracle."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(97531)
    db = [gen.random_codes(int(rng.integers(15000, 30000)), rng) for _ in range(4)]
    assays = gen.make_assays(rng, db, 6, "taqman", variants=3)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    want = {(t, i): oracle.search(codes, a[0], a[1], a[2], o) for t, codes in enumerate(db) for i, a in enumerate(assays)}
    assert sum(len(v) for v in want.values()) >= 6
    for env in ({}, {"TNT_NO_LEAN": "1"}, {"TNT_NO_LEAN_SKIP": "1"}):
        for k in ("TNT_NO_LEAN", "TNT_NO_LEAN_SKIP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Engine() as e:
            for c in db:
                e.add_target(c)
            e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
            got = e.search(to_opts(o))
            n = 0
            for (t, i), w in want.items():
                mine = [h for h in got if h.target_id == t and h.assay_index == i]
                assert_hits_equal(e, mine, w, assays[i])
                n += len(mine)
            assert n == len(got)


def test_upload_paths_agree(engine_lib, monkeypatch):
    """Fragments reach the device through a ring of staging slots on an upload stream while stage 1
    already runs on the first batches.  Pageable vs page-locked sources, one call per fragment vs
    tnt_engine_add_targets, a full-size ring vs a two-slot ring that has to recycle slots (and with
    them the lazy emission of the non-ACGT lists): the hit lists must be identical."""
    import torch
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(20261)
    nfrag, flen = 5, 24_000_000              # 120 MB: four upload batches
    F, R, P = gen.rand_oligo(20, rng), gen.rand_oligo(21, rng), gen.rand_oligo(25, rng)
    amp = F + gen.rand_oligo(30, rng) + P + gen.rand_oligo(40, rng) + gen.revcomp(R)
    pinned = torch.empty(nfrag * flen, dtype=torch.uint8, pin_memory=True).numpy()
    pinned[:] = rng.integers(0, 4, size=nfrag * flen, dtype=np.uint8)
    frags_pinned = [pinned[i * flen:(i + 1) * flen] for i in range(nfrag)]
    for i, f in enumerate(frags_pinned):
        gen.plant(f, 1000 + i * 4_000_000, amp if i % 2 == 0 else gen.revcomp(amp))
        gen.plant(f, flen - 5000, gen.mutate(amp, 2, rng))
        f[rng.integers(0, flen, size=2000)] = 15       # N: exceptions in every batch
        f[10_000_000:10_000_400] = 15
    frags_pageable = [f.copy() for f in frags_pinned]
    opts = H.default_options(min_primer_tm=45.0, min_probe_tm=45.0)
    assays = [Assay(0, F, R, P)]

    def run(frags, batch_call, slots):
        if slots:
            monkeypatch.setenv("TNT_UPLOAD_SLOTS", str(slots))
        else:
            monkeypatch.delenv("TNT_UPLOAD_SLOTS", raising=False)
        with Engine() as e:
            e.set_assays(assays)
            if batch_call:
                e.add_targets(frags)
            else:
                for f in frags:
                    e.add_target(f)
            hits = e.search(to_opts(opts))
            out = [(h.target_id,) + hit_key(e, h, (F, R, P)) + hit_floats(h) for h in hits]
            # a second search on the now resident fragments must not change anything
            again = e.search(to_opts(opts))
            assert [(h.target_id,) + hit_key(e, h, (F, R, P)) + hit_floats(h) for h in again] == out
            return out

    base = run(frags_pageable, False, 0)
    assert len(base) >= 2 * nfrag
    assert run(frags_pinned, True, 0) == base
    assert run(frags_pageable, True, 2) == base
    assert run(frags_pinned, False, 2) == base


def test_incremental_targets(engine_lib, oracle):
    """Fragments registered after a search join the resident set: the next search covers old and
    new fragments and equals a fresh engine that got all of them at once."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(1717)
    db = [gen.random_codes(int(rng.integers(20000, 40000)), rng) for _ in range(5)]
    assays = gen.make_assays(rng, db, 4, "taqman", variants=2)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)

    def keys(e, hits):
        return [(h.target_id, h.assay_index) + hit_key(e, h, assays[h.assay_index]) + hit_floats(h) for h in hits]

    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        for c in db[:2]:
            e.add_target(c)
        first = keys(e, e.search(to_opts(o)))
        for c in db[2:]:
            e.add_target(c)
        both = keys(e, e.search(to_opts(o)))
    with Engine() as e:
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        e.add_targets(db)
        fresh = keys(e, e.search(to_opts(o)))
    assert both == fresh
    assert [k for k in both if k[0] < 2] == first
    assert len(fresh) >= 4


def test_hits_on_a_threshold_are_listed_separately(engine_lib, oracle):
    """north_star: hits whose Tm lies within the comparison tolerance (0.01 C / 0.001 kcal/mol) of
    a filter bound are listed separately.  With the bound put exactly on the Tm of a known hit the
    hit is still reported (the reference rejects `tm < min`, bind_oligo.cpp:598-607), equals the
    oracle's, and tnt_engine_hits_near_threshold names it; with the default bounds nothing is listed."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(606)
    codes, F, R, P = gen.make_pcr_case(rng, 40000, n_sites=3, probe=True)
    with Engine() as e:
        e.add_target(codes)
        e.set_assays([Assay(0, F, R, P)])
        o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
        hits = e.search(to_opts(o))
        assert len(hits) >= 1 and e.hits_near_threshold() == []
        # the weaker primer of the best hit: that hit stays and sits exactly on the bound
        edge = float(np.float32(max(min(h.forward.tm, h.reverse.tm) for h in hits)))
        o2 = H.default_options(min_primer_tm=edge, min_probe_tm=40.0)
        got = e.search(to_opts(o2))
        want = oracle.search(codes, F, R, P, o2)
        assert_hits_equal(e, got, want, (F, R, P))
        near = e.hits_near_threshold()
        assert near and all(min(abs(got[i].forward.tm - edge), abs(got[i].reverse.tm - edge)) <= TM_TOL for i in near)
        others = [i for i in range(len(got)) if i not in near]
        assert all(min(abs(got[i].forward.tm - edge), abs(got[i].reverse.tm - edge)) > TM_TOL for i in others)


def test_oligo_dimers_on_the_device(eng, oracle):
    """tnt_engine_oligo_dimer (generic NucCruc kernel with the second oligo as an explicit target,
    symmetry entropy for homodimers) against the oracle: bit-identical Tm / dH / dS, identical text."""
    import json
    import os
    rng = np.random.default_rng(515)
    cases = [(fx["q"], fx["t"], fx["ca"], fx["cb"]) for fx in
             json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dimers.json")))]
    for it in range(150):
        L = int(rng.integers(8, 45))
        q = gen.rand_oligo(L, rng)
        if it % 4 == 0:
            h = gen.rand_oligo(L // 2, rng)
            q = h + gen.revcomp(h)
        t = gen.rand_oligo(int(rng.integers(8, 60)), rng) if it % 2 else None
        if it % 6 == 1:
            t = gen.mutate(gen.revcomp(q), int(rng.integers(0, 4)), rng)
        cases.append((q, t, *[(9e-7, 9e-7), (2e-6, 5e-7)][it % 2]))
    for q, t, ca, cb in cases:
        want = oracle.dimer(q, t, conc_a=ca, conc_b=cb)
        got = eng.oligo_dimer(q, t, ca, cb)
        assert got.valid == want.valid, (q, t)
        assert abs(got.tm - want.tm) <= TM_TOL
        assert (got.tm, got.dH, got.dS) == (want.tm, want.dH, want.dS), (q, t)
        if want.valid:
            assert got.alignment == want.alignment, (q, t)
    # a resident database and a search are not disturbed by the explicit-target launches
    codes, F, R, P = gen.make_pcr_case(rng, 30000, n_sites=2, probe=True)
    eng.clear_targets()
    eng.add_target(codes)
    from thermonucleotideblast_b200 import Assay
    eng.set_assays([Assay(0, F, R, P)])
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    before = [hit_key(eng, h, (F, R, P)) for h in eng.search(to_opts(o))]
    eng.oligo_dimer(F, R)
    eng.oligo_dimer(P)
    assert [hit_key(eng, h, (F, R, P)) for h in eng.search(to_opts(o))] == before and before
    eng.clear_targets()


def test_hairpins_on_the_device(eng, oracle):
    """tnt_engine_oligo_hairpin (k_oligo_jobs: the triangular fill of align_hairpin, the three evaluations
    per traceback of enumerate_hairpin_alignments, loop entropy / special loops / terminal mismatch)
    against the oracle and the committed vectors of the compiled reference: floats bit for bit."""
    import json
    import os
    rng = np.random.default_rng(616)
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hairpins.json")))
    f32 = lambda x: float(np.float32(x)).hex()
    nvalid = 0
    for r in gold:
        if (r["T"], r["na"]) != (310.15, 0.05):
            continue
        g = eng.oligo_hairpin(r["q"])
        w = r["out"]
        assert g.valid == w["valid"], r["q"]
        assert (f32(g.tm), f32(g.dH), f32(g.dS)) == (w["tm"], w["dH"], w["dS"]), r["q"]
        if g.valid:
            assert [g.q_first, g.t_first] == w["loop"] and [g.q_last, g.t_last] == w["open_end"] and g.num_gap == w["columns"], r["q"]
            nvalid += 1
    for it in range(300):
        L = int(rng.integers(5, 57))
        q = gen.rand_oligo(L, rng)
        if it % 2 and L >= 16:
            stem = gen.rand_oligo(int(rng.integers(3, 10)), rng)
            q = (gen.rand_oligo(int(rng.integers(0, 3)), rng) + stem + gen.rand_oligo(int(rng.integers(3, 9)), rng) +
                 gen.mutate(gen.revcomp(stem), int(rng.integers(0, 2)), rng))[:56]
        w = oracle.hairpin(q)
        g = eng.oligo_hairpin(q)
        assert g.valid == w.valid, q
        assert abs(g.tm - w.tm) <= TM_TOL
        assert (g.tm, g.dH, g.dS) == (w.tm, w.dH, w.dS), q
        if w.valid:
            assert (g.q_first, g.t_first, g.q_last, g.t_last, g.num_gap) == (w.q_first, w.t_first, w.q_last, w.t_last, w.num_gap), q
            nvalid += 1
    assert nvalid > 200


def test_assay_structures_in_one_call(eng, oracle):
    """tnt_engine_assay_structures: the seven temperatures tntblast_local.cpp:657-686 attaches to every hit
    (plus the two single-primer heterodimers), for all assays with one launch, against the oracle."""
    from thermonucleotideblast_b200 import Assay, search_options
    rng = np.random.default_rng(717)
    assays = []
    for i in range(40):
        F, R, P = gen.rand_oligo(int(rng.integers(16, 31)), rng), gen.rand_oligo(int(rng.integers(16, 31)), rng), gen.rand_oligo(int(rng.integers(18, 36)), rng)
        if i % 5 == 0:
            h = gen.rand_oligo(9, rng)
            F = h + "GAAA" + gen.revcomp(h)       # a hairpin with a special tetra-loop, and a strong homodimer
        if i % 7 == 0:
            R = gen.revcomp(F)[:len(F) - 2]        # primer dimer
        assays.append(Assay(i, F, R, P) if i % 3 else (Assay(i, F, R, None) if i % 2 else Assay(i, None, None, P)))
    eng.set_assays(assays)
    o = search_options(min_primer_tm=45.0, min_probe_tm=50.0, forward_primer_strand=1.8e-6)   # asymmetric PCR: c_f != c_r
    got = eng.assay_structures(o, len(assays))
    fps, rps, ps = o.forward_primer_strand, o.reverse_primer_strand, o.probe_strand
    ne = 0
    for a, g in zip(assays, got):
        want_h = [oracle.hairpin(x).tm if x else -1.0 for x in (a.forward, a.reverse, a.probe)]
        want_d = [oracle.dimer(x, None, conc_a=c, conc_b=c).tm if x else -1.0 for x, c in ((a.forward, fps), (a.reverse, rps), (a.probe, ps))]
        if a.forward:
            want_x = [oracle.dimer(a.forward, a.reverse, conc_a=fps, conc_b=rps).tm, oracle.dimer(a.forward, a.forward, conc_a=fps, conc_b=rps).tm,
                      oracle.dimer(a.reverse, a.reverse, conc_a=fps, conc_b=rps).tm]
        else:
            want_x = [-1.0, -1.0, -1.0]
        assert list(g.hairpin_tm) == [np.float32(x) for x in want_h], a
        assert list(g.homodimer_tm) == [np.float32(x) for x in want_d], a
        assert list(g.heterodimer_tm) == [np.float32(x) for x in want_x], a
        ne += sum(1 for x in want_h + want_d + want_x if x > 0)
    assert ne > 60


def test_device_pairing_equals_host_join(engine_lib, oracle, monkeypatch):
    """The F x R (x P) join on the device (k_pair_keys / radix sorts / k_pair_unique / k_pair_join) against
    the same loops on the host (TNT_HOST_JOIN=1) and against the oracle: ordered hit lists, also with
    single-primer amplicons forbidden, a min-max clamp, duplicate sites (indel copies), nested
    amplicons and probe-only assays mixed into the PCR run."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(4711)
    db = [gen.random_codes(int(rng.integers(30000, 80000)), rng) for _ in range(6)]
    assays = gen.make_assays(rng, db, 6, "taqman", variants=4) + gen.make_assays(rng, db, 6, "pcr", variants=4) + \
        gen.make_assays(rng, db, 3, "probe", variants=3)
    # nested amplicons: a second reverse site behind the first one
    F, R, _ = assays[6]
    gen.plant(db[0], 5000, F + gen.rand_oligo(120, rng) + gen.revcomp(R) + gen.rand_oligo(60, rng) + gen.revcomp(R))
    total = 0
    for single, mmc, max_len in ((1, -1, 2000), (0, 3, 600)):
        o = H.default_options(min_primer_tm=38.0, min_probe_tm=38.0, max_len=max_len, single_primer_pcr=single, min_max_primer_clamp=mmc)

        def run():
            with Engine() as e:
                e.add_targets(db)
                e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
                return [(h.target_id, h.assay_index) + hit_key(e, h, assays[h.assay_index]) + hit_floats(h) for h in e.search(to_opts(o))]

        dev = run()
        monkeypatch.setenv("TNT_HOST_JOIN", "1")
        host = run()
        monkeypatch.delenv("TNT_HOST_JOIN")
        assert dev == host and len(dev) >= 20
        k = 0
        for t, codes in enumerate(db):
            for i, a in enumerate(assays):
                want = oracle.search(codes, a[0], a[1], a[2], o)
                mine = [x for x in dev if x[0] == t and x[1] == i]
                assert len(mine) == len(want)
                assert [x[2:2 + len(w.exact_key())] for x, w in zip(mine, want)] == [w.exact_key() for w in want]
                k += len(want)
        assert k == len(dev)
        total += k
    assert total >= 50


@pytest.mark.parametrize("mode", ["auto", "sparse", "dense"])
@pytest.mark.parametrize("W", [4, 5, 6, 8])
def test_other_word_sizes(engine_lib, oracle, monkeypatch, W, mode):
    """DNAHash word sizes other than the default 7 (`-W`, tntblast.h:68; DNAHash accepts 3..8): seeds bit-exact
    and searches equal to the oracle through every scan kernel (table sizes 4^W = 256 ... 65536 keys: the
    rank-compressed shared-memory table, the sparse group bitmap and the region scan all depend on W).
    Short words in the sparse kernel overflow its group queue, i.e. take its position-by-position path, on
    fragments whose length is no multiple of 32 (found by tests/fuzz_parity.py: that loop must keep the warp
    together)."""
    from thermonucleotideblast_b200 import Assay, Engine
    if mode != "auto":
        monkeypatch.setenv("TNT_SCAN_MODE", mode)
    rng = np.random.default_rng(1000 + W)
    n = 12000 if W <= 5 else 60000
    db = [gen.random_codes(n, rng), gen.random_codes(n // 2 + 37, rng), gen.random_codes(6145, rng)]
    gen.sprinkle_degenerate(db[0], rng, frac=1e-3, n_runs_per_50kb=4)
    assays = gen.make_assays(rng, db, 3, "taqman", variants=3)
    o = H.default_options(min_primer_tm=42.0, min_probe_tm=42.0, word_size=W)
    with Engine(word_size=W) as e:
        ids = [e.add_target(c) for c in db]
        ol = assays[0][0]
        total = 0
        for tid, codes in zip(ids, db):
            for oligo in (ol, gen.rand_oligo(25, rng), ol[:9] + "N" + ol[10:]):
                for plus in (False, True):
                    want = oracle.seeds(codes, oligo, W, plus, unique=True)
                    assert e.seeds(tid, oligo, plus) == want, (W, tid, oligo, plus)
                    total += len(want)
        assert total > 50
        e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
        got = e.search(to_opts(o))
        nh = 0
        for t, codes in enumerate(db):
            for i, a in enumerate(assays):
                want = oracle.search(codes, a[0], a[1], a[2], o)
                assert_hits_equal(e, [h for h in got if h.target_id == t and h.assay_index == i], want, a)
                nh += len(want)
        assert nh >= 3


FILTER_CASES = [
    dict(max_gap=0),
    dict(max_mismatch=1),
    dict(max_gap=1, max_mismatch=2, max_poly_degen=0),
    dict(max_primer_tm=58.0, max_probe_tm=62.0),
    dict(min_primer_tm=0.0, min_probe_tm=0.0, max_primer_dg=-9.0, max_probe_dg=-10.0),
    dict(min_primer_dg=-16.0, min_probe_dg=-20.0),
    dict(min_primer_tm=38.0, max_primer_dg=-7.5, min_primer_dg=-30.0, max_primer_tm=75.0),
    dict(primer_clamp=4, probe_clamp_5=3, probe_clamp_3=3),
    dict(min_max_primer_clamp=6),
    dict(target_strand=1), dict(target_strand=2),
    dict(single_primer_pcr=0, max_len=300),
]


@pytest.mark.parametrize("case", range(len(FILTER_CASES)))
def test_search_filter_cascade(engine_lib, oracle, case):
    """Every per-oligo bound of the reference's filter cascade (bind_oligo.cpp:593-705: Tm window, dG window,
    clamps, gaps, mismatches, degenerate runs; amplicon_search.cpp:359-441: length, min-max clamp, strands,
    single-primer products) alone and in combinations, TaqMan and probe-only assays, lists equal to the
    oracle's; where the compiled reference travels with the repository it is asked as well."""
    from thermonucleotideblast_b200 import Assay, Engine
    kw = dict(min_primer_tm=36.0, min_probe_tm=36.0)
    kw.update(FILTER_CASES[case])
    rng = np.random.default_rng(5000 + case)
    db = [gen.random_codes(int(rng.integers(20000, 40000)), rng) for _ in range(3)]
    gen.sprinkle_degenerate(db[1], rng, frac=2e-3, n_runs_per_50kb=6)
    taq = gen.make_assays(rng, db, 4, "taqman", variants=5)
    prb = gen.make_assays(rng, db, 2, "probe", variants=5)
    ref = H.ref() if H.have_ref() else None
    nhits = 0
    for assays, fmt in ((taq, H.ASSAY_PCR), (prb, H.ASSAY_PROBE)):
        o = H.default_options(assay_format=fmt, **kw)
        with Engine() as e:
            for c in db:
                e.add_target(c)
            e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
            got = e.search(to_opts(o))
            n = 0
            for t, codes in enumerate(db):
                for i, a in enumerate(assays):
                    want = oracle.search(codes, a[0], a[1], a[2], o)
                    mine = [h for h in got if h.target_id == t and h.assay_index == i]
                    assert_hits_equal(e, mine, want, a)
                    if ref is not None and t == 0:
                        r = ref.search(codes, a[0], a[1], a[2], o)
                        assert [(h.exact_key(), h.floats()) for h in r] == [(h.exact_key(), h.floats()) for h in want], (case, fmt, i)
                    n += len(mine)
            assert n == len(got)
            nhits += n
    # a cascade that rejects everything would make the comparison vacuous
    assert nhits >= 1 or case in (2,), (case, nhits)


def test_probe_sites_overhanging_the_fragment_ends(engine_lib, oracle):
    """A probe site cut by the end (or the start) of a fragment keeps coordinates beyond the fragment, and the
    reference prints its text from the clamped end for the full length (probe_search.cpp:129-142, :205-219):
    bases outside [probe_first, probe_last].  Found by tests/fuzz_parity.py (the text fetch threw)."""
    from thermonucleotideblast_b200 import Assay, Engine
    rng = np.random.default_rng(31337)
    P = gen.rand_oligo(38, rng)
    site = gen.revcomp(P)
    n_over = 0
    for strand_text in (site, P):
        n = 9000
        codes = gen.random_codes(n, rng)
        gen.plant(codes, n - 26, strand_text[:26])      # the last 12 bases of the site lie beyond the end
        gen.plant(codes, 0, strand_text[10:])           # the first 10 in front of the start
        gen.plant(codes, 4000, strand_text)
        o = H.default_options(assay_format=H.ASSAY_PROBE, min_probe_tm=30.0)
        want = oracle.search(codes, None, None, P, o)
        n_over += sum(1 for h in want if h.probe_first < 0 or h.probe_last >= n)
        with Engine() as e:
            e.add_target(codes)
            e.set_assays([Assay(0, None, None, P)])
            got = e.search(to_opts(o))
            assert_hits_equal(e, got, want, (None, None, P))
            assert e.hit_sequences() == [e.hit_sequence(h) for h in got]
    assert n_over >= 2
