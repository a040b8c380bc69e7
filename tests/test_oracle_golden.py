"""The plain-C oracle against the committed golden vectors.

tests/golden/*.json were produced by tests/golden/make_golden.py from the UNMODIFIED reference
(oracle/_ref/libtntref.so).  This is what pins the oracle where /root/reference is absent.
"""
import json
import os

import numpy as np
import pytest

import gen
import harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NB = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'I': 4, 'M': 7, 'R': 8, 'S': 9, 'V': 10, 'W': 11, 'Y': 12, 'H': 13,
      'K': 14, 'D': 15, 'B': 16, 'N': 17}


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def f32(x):
    return float(np.float32(x)).hex()


def align_rec(a):
    return {"tm": f32(a.tm), "dH": f32(a.dH), "dS": f32(a.dS), "dG": f32(a.dG), "valid": a.valid,
            "ints": [a.anchor5, a.anchor3, a.num_mismatch, a.num_gap, a.max_poly_degen,
                     a.q_first, a.q_last, a.t_first, a.t_last, a.target_start, a.target_stop, a.loc_5, a.loc_3],
            "alignment": a.alignment.decode()}


def test_delta_g_tables(oracle):
    for rec in load("tables.json"):
        t = oracle.dump_tables(rec["T"], rec["na"])
        assert list(t.delta_g) == rec["delta_g"]


def test_alignments(oracle):
    recs = load("alignments.json")
    nvalid = 0
    for r in recs:
        tb = np.array([NB[c] for c in r["t"]], dtype=np.uint8)
        a = oracle.align(r["q"], tb, T=r["T"], na=r["na"], ct=r["ct"], dangle5=r["d5"], dangle3=r["d3"])
        got = align_rec(a)
        want = r["out"]
        if not want["valid"]:
            assert not got["valid"]
            continue
        nvalid += 1
        assert got == want, (r["q"], r["t"])
    assert nvalid > 300


def test_seeds_and_windows(oracle):
    for rec in load("seeds.json"):
        codes = gen.str_to_codes(rec["codes"])
        for plus in (0, 1):
            assert [list(x) for x in oracle.seeds(codes, rec["oligo"], rec["W"], bool(plus), unique=False)] == rec["raw%d" % plus]
            uniq = oracle.seeds(codes, rec["oligo"], rec["W"], bool(plus), unique=True)
            assert [list(x) for x in uniq] == rec["uniq%d" % plus]
            for (q, t), want in zip(uniq, rec.get("bind%d" % plus, [])):
                got = align_rec(oracle.bind_window(codes, rec["oligo"], bool(plus), q, t))
                if want["valid"]:
                    assert got == want
                else:
                    assert not got["valid"]


def test_searches(oracle):
    total = 0
    for rec in load("searches.json"):
        codes = gen.str_to_codes(rec["codes"])
        o = H.default_options(**rec["opts"])
        hits = oracle.search(codes, rec["F"], rec["R"], rec["P"], o)
        got = [{"key": [x.decode() if isinstance(x, bytes) else x for x in h.exact_key()],
                "floats": [f32(x) for x in h.floats()]} for h in hits]
        assert got == rec["hits"], rec["kind"]
        total += len(got)
    assert total >= 50


def test_readme_known_answer(oracle):
    """Reference README.md:138,162-203 (gibb-marburg): dG/dH/dS, clamps, coordinates, strings."""
    amp = ("TTCCCCTTTGGAGGCATCCAAGCGATGGGCTTTCAGGACAGGTGTACCTCCCAAGAATGTTGAGTATACAGAAGGGGAGGAAGCCAAAACATGCTACAATATAAG"
           "TGTAACGGATCCCTCTGGAAAATCCTTGCTGTTGGATCCTCC")
    rng = np.random.default_rng(1)
    codes = gen.random_codes(6121 + len(amp) + 3000, rng)
    gen.plant(codes, 6121, amp)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=45.0)
    hits = oracle.search(codes, "TTCCCCTTTGGAGGCATC", "GGAGGATCCAACAGCAAGG", "CGATGGGCTTTCAGGACAGGTGT", o)
    assert len(hits) == 1
    h = hits[0]
    T = 310.15
    assert (h.amp_first, h.amp_last, h.probe_first, h.probe_last) == (6121, 6267, 6143, 6165)
    assert abs(h.forward_dH - T * h.forward_dS + 16.8574) < 1e-3 and abs(h.forward_dH + 135.5) < 1e-3
    assert abs(h.forward_dS + 0.382533) < 1e-5
    assert abs(h.reverse_dH - T * h.reverse_dS + 17.8955) < 1e-3 and abs(h.reverse_dH + 146.5) < 1e-3
    assert abs(h.probe_dH - T * h.probe_dS + 22.9778) < 1e-3 and abs(h.probe_dH + 180.2) < 1e-3
    assert min(h.forward_clamp, h.reverse_clamp) == 18
    assert h.amplicon_len == 147
    assert h.forward_align == b"5' TTCCCCTTTGGAGGCATC 3'\n   ||||||||||||||||||\n3' AAGGGGAAACCTCCGTAG 5'"
    assert h.reverse_align == b"5' GGAGGATCCAACAGCAAGG 3'\n   |||||||||||||||||||\n3' CCTCCTAGGTTGTCGTTCC 5'"
    assert h.probe_align == b"5' CGATGGGCTTTCAGGACAGGTGT 3'\n   |||||||||||||||||||||||\n3' GCTACCCGAAAGTCCTGTCCACA 5'"
    assert h.amplicon_head.decode() == amp


def test_oracle_vs_compiled_reference(oracle, ref):
    """Differential run against the reference itself (skipped where oracle/_ref is absent)."""
    rng = np.random.default_rng(5)
    for it in range(300):
        L = int(rng.integers(14, 34))
        q = gen.rand_oligo(L, rng)
        t = gen.rand_oligo(4, rng) + gen.mutate(gen.revcomp(q), int(rng.integers(0, 5)), rng) + gen.rand_oligo(4, rng)
        tb = np.array([NB[c] for c in t], dtype=np.uint8)
        a, b = ref.align(q, tb), oracle.align(q, tb)
        assert a.key() == b.key() and (a.tm, a.dH, a.dS) == (b.tm, b.dH, b.dS)
    for it in range(12):
        codes, F, R, P = gen.make_pcr_case(rng, 30000, n_sites=3, probe=bool(it % 2))
        o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
        a, b = ref.search(codes, F, R, P, o), oracle.search(codes, F, R, P, o)
        assert [(h.exact_key(), h.floats()) for h in a] == [(h.exact_key(), h.floats()) for h in b]


def _fasta_view(recs):
    return [{"approx_len": a, "defline": d, "codes": gen.codes_to_str(w),
             "pieces": [[s0, s1, gen.codes_to_str(c)] for s0, s1, c in pcs]} for _, a, d, w, pcs in recs]


def test_fasta_reader_golden(oracle):
    """FASTA index, defline, base codes and the driver's fragment queue against vectors produced by
    the compiled reference's own sequence_data class (tests/golden/make_golden.py fasta)."""
    import base64
    fixtures = load("fasta.json")
    assert len(fixtures) >= 20
    for fx in fixtures:
        text = base64.b64decode(fx["text"])
        got = _fasta_view(oracle.fasta_records(text, threshold=fx["threshold"], overlap=fx["overlap"]))
        assert got == fx["records"]


def test_fasta_reader_vs_compiled_reference(oracle, ref, tmp_path):
    rng = np.random.default_rng(11)
    for it in range(20):
        text = gen.rand_fasta(rng, n_records=int(rng.integers(1, 8)), max_len=3000, width=[60, 80, 0, 7][it % 4],
                              crlf=bool(it % 3 == 1), iupac=0.02)
        path = tmp_path / ("t%d.fna" % it)
        path.write_bytes(text)
        for threshold, overlap in [(0, 0), (500, 30), (64, 5)]:
            a = _fasta_view(ref.fasta_records(text, str(path), threshold=threshold, overlap=overlap))
            b = _fasta_view(oracle.fasta_records(text, threshold=threshold, overlap=overlap))
            assert a == b
    for length, max_len in [(1, 10), (10, 10), (11, 10), (1000, 7), (500000, 500000), (500001, 500000), (5062500, 500000)]:
        assert ref.seq_len_increment(length, max_len) == oracle.seq_len_increment(length, max_len)


def test_oligo_dimers_golden(oracle):
    """Homodimer / heterodimer Tm of oligos (tntblast_local.cpp:657-686) against vectors from the
    compiled reference: floats bit-identical; the alignment text is compared for heterodimers only
    (for a homodimer the reference's printer takes the unaligned flanks from the target buffer that
    set_duplex filled with the reverse complement -- a display quirk, the numbers are what is stored)."""
    fixtures = load("dimers.json")
    assert len(fixtures) >= 100
    for fx in fixtures:
        a = oracle.dimer(fx["q"], fx["t"], conc_a=fx["ca"], conc_b=fx["cb"])
        assert (f32(a.tm), f32(a.dH), f32(a.dS), f32(a.dG), a.valid) == (fx["tm"], fx["dH"], fx["dS"], fx["dG"], fx["valid"])
        if fx["t"]:
            assert a.alignment.decode() == fx["alignment"]


def test_oligo_dimers_vs_compiled_reference(oracle, ref):
    rng = np.random.default_rng(99)
    for it in range(600):
        L = int(rng.integers(8, 45))
        q = gen.rand_oligo(L, rng)
        if it % 4 == 0:
            h = gen.rand_oligo(L // 2, rng)
            q = h + gen.revcomp(h)
        t = gen.rand_oligo(int(rng.integers(8, 45)), rng) if it % 2 else None
        if it % 6 == 1:
            t = gen.mutate(gen.revcomp(q), int(rng.integers(0, 4)), rng)
        ca, cb = [(9e-7, 9e-7), (2e-6, 5e-7), (1e-7, 3e-6)][it % 3]
        a, b = ref.dimer(q, t, conc_a=ca, conc_b=cb), oracle.dimer(q, t, conc_a=ca, conc_b=cb)
        assert (a.tm, a.dH, a.dS, a.dG, a.valid) == (b.tm, b.dH, b.dS, b.dG, b.valid)
        if t:
            assert a.alignment == b.alignment


def test_cull_anomaly_cases(oracle):
    """Two neighbourhoods of the config-5 data in which the reference's cull loses a real amplicon
    (tests/golden/make_cull_cases.py): the oracle has to report what the compiled reference reported,
    i.e. without the lost amplicon."""
    cases = load("cull_cases.json")
    assert len(cases) == 2
    for c in cases:
        codes = gen.str_to_codes(c["codes"])
        o = H.default_options(min_primer_tm=c["min_primer_tm"])
        hits = oracle.search(codes, c["forward"], c["reverse"], None, o)
        got = [{"amp_first": h.amp_first, "amp_last": h.amp_last, "primer_strand": h.primer_strand,
                "forward_tm": float(np.float32(h.forward_tm)), "reverse_tm": float(np.float32(h.reverse_tm)),
                "forward_align": h.forward_align.decode(), "reverse_align": h.reverse_align.decode()} for h in hits]
        assert got == c["reference_hits"]
        assert tuple(c["lost_amplicon"]) not in [(h.amp_first, h.amp_last) for h in hits]


def test_hairpins(oracle):
    """approximate_tm_hairpin (tntblast_local.cpp:661,667,682): the oracle against vectors made from the
    compiled reference (tests/golden/make_hairpins.py), floats bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_hairpins", os.path.join(GOLD, "make_hairpins.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    recs = load("hairpins.json")
    assert len(recs) > 200 and sum(r["out"]["valid"] for r in recs) > 100
    for r in recs:
        assert mk.rec(oracle.hairpin(r["q"], r["T"], r["na"])) == r["out"], r["q"]


def test_hairpins_against_the_compiled_reference(oracle, ref):
    rng = np.random.default_rng(31)
    n = 0
    for it in range(1500):
        L = int(rng.integers(5, 57))
        q = gen.rand_oligo(L, rng)
        if it % 2 and L >= 16:
            stem = gen.rand_oligo(int(rng.integers(3, 9)), rng)
            q = (stem + gen.rand_oligo(int(rng.integers(3, 9)), rng) + gen.revcomp(stem) + gen.rand_oligo(3, rng))[:56]
        T, na = [(310.15, 0.05), (285.0, 0.5), (340.0, 0.02)][it % 3]
        a, b = ref.hairpin(q, T, na), oracle.hairpin(q, T, na)
        assert (a.valid, a.tm, a.dH, a.dS, a.dp_dg) == (b.valid, b.tm, b.dH, b.dS, b.dp_dg), q
        if a.valid:
            assert (a.q_first, a.t_first, a.q_last, a.t_last, a.num_gap) == (b.q_first, b.t_first, b.q_last, b.t_last, b.num_gap), q
            n += 1
    assert n > 800


@pytest.fixture
def dinkelbach_oracle(oracle):
    oracle.set_dinkelbach(True)
    yield oracle
    oracle.set_dinkelbach(False)


def _load_make_dinkelbach():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_dinkelbach", os.path.join(GOLD, "make_dinkelbach.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def test_dinkelbach_golden(dinkelbach_oracle):
    """NucCruc::dinkelbach(true) (nuc_cruc.cpp:2399-2440, :2459-2500, :2548-2588): the iterative Tm of single
    windows, oligo dimers, hairpins and whole searches against vectors made from the compiled reference
    (tests/golden/make_dinkelbach.py), floats bit for bit."""
    oracle = dinkelbach_oracle
    mk = _load_make_dinkelbach()
    gold = load("dinkelbach.json")
    assert len(gold["windows"]) >= 400 and sum(r["out"]["valid"] for r in gold["windows"]) > 300
    for r in gold["windows"]:
        tb = np.array([NB[c] for c in r["t"]], dtype=np.uint8)
        assert mk.align_rec(oracle.align(r["q"], tb, T=r["T"], na=r["na"], ct=r["ct"])) == r["out"], (r["q"], r["t"])
    for r in gold["dimers"]:
        assert mk.struct_rec(oracle.dimer(r["q"], r["t"])) == r["out"], (r["q"], r["t"])
    for r in gold["hairpins"]:
        assert mk.struct_rec(oracle.hairpin(r["q"])) == r["out"], r["q"]
    cases = {(kind, t, i): (db[t], assays[i]) for kind, db, assays in mk.search_cases() for t in range(len(db)) for i in range(len(assays))}
    nhits = 0
    for s in gold["searches"]:
        codes, a = cases[(s["kind"], s["target"], s["assay"])]
        hits = oracle.search(codes, a[0], a[1], a[2], mk.search_options_for(s["kind"]))
        assert [mk.hit_rec(h) for h in hits] == s["hits"], (s["kind"], s["target"], s["assay"])
        nhits += len(hits)
    assert nhits > 20


def test_dinkelbach_changes_the_answer(oracle):
    """The mode is not a no-op: a fifth of partially matching windows change their record."""
    rng = np.random.default_rng(7)
    changed = 0
    for it in range(300):
        q = gen.rand_oligo(int(rng.integers(16, 30)), rng)
        t = H.encode(gen.rand_oligo(4, rng) + gen.mutate(gen.revcomp(q), int(rng.integers(0, 5)), rng, indel=False) + gen.rand_oligo(4, rng))
        oracle.set_dinkelbach(False)
        a = oracle.align(q, t).key()
        oracle.set_dinkelbach(True)
        b = oracle.align(q, t).key()
        oracle.set_dinkelbach(False)
        changed += a != b
    assert 20 < changed < 300


def test_dinkelbach_vs_compiled_reference(dinkelbach_oracle, ref):
    oracle = dinkelbach_oracle
    ref.set_dinkelbach(True)
    try:
        rng = np.random.default_rng(4242)
        for it in range(4000):
            L = int(rng.integers(10, 45))
            q = gen.rand_oligo(L, rng)
            t = gen.rand_oligo(4, rng) + gen.mutate(gen.revcomp(q), int(rng.integers(0, 6)), rng) + gen.rand_oligo(4, rng)
            if it % 4 == 0:
                t = gen.rand_oligo(L + 8, rng)
            T, na = [(310.15, 0.05), (285.0, 0.5), (340.0, 0.02)][it % 3]
            tb = H.encode(t)
            a, b = ref.align(q, tb, T=T, na=na), oracle.align(q, tb, T=T, na=na)
            assert a.key() == b.key() and (a.valid, a.tm, a.dH, a.dS, a.dG) == (b.valid, b.tm, b.dH, b.dS, b.dG), (q, t)
            if it % 8 == 0:
                for x, y in ((ref.hairpin(q, T, na), oracle.hairpin(q, T, na)), (ref.dimer(q, None, T, na), oracle.dimer(q, None, T, na))):
                    assert (x.valid, x.tm, x.dH, x.dS) == (y.valid, y.tm, y.dH, y.dS), q
        for kind in ("pcr", "taqman", "probe", "padlock"):
            db = [gen.random_codes(15000, rng) for _ in range(2)]
            assays = gen.make_assays(rng, db, 3, kind, variants=3)
            o = _load_make_dinkelbach().search_options_for(kind)
            n = 0
            for codes in db:
                for a in assays:
                    x, y = ref.search(codes, a[0], a[1], a[2], o), oracle.search(codes, a[0], a[1], a[2], o)
                    assert [(h.exact_key(), h.floats()) for h in x] == [(h.exact_key(), h.floats()) for h in y], kind
                    n += len(x)
            assert n >= 3, kind
    finally:
        ref.set_dinkelbach(False)


def test_filter_cascade_vs_compiled_reference(oracle, ref):
    """Every bound of the reference's filter cascade, alone and in combinations (the cases of
    tests/test_gpu_parity.py::test_search_filter_cascade): the oracle's hit lists equal the compiled
    reference's, floats bit for bit -- the GPU test compares the engine with the oracle on the same cases."""
    from test_gpu_parity import FILTER_CASES
    total = 0
    for case, extra in enumerate(FILTER_CASES):
        kw = dict(min_primer_tm=36.0, min_probe_tm=36.0)
        kw.update(extra)
        rng = np.random.default_rng(5000 + case)
        db = [gen.random_codes(int(rng.integers(20000, 40000)), rng) for _ in range(3)]
        gen.sprinkle_degenerate(db[1], rng, frac=2e-3, n_runs_per_50kb=6)
        taq = gen.make_assays(rng, db, 4, "taqman", variants=5)
        prb = gen.make_assays(rng, db, 2, "probe", variants=5)
        for assays, fmt in ((taq, H.ASSAY_PCR), (prb, H.ASSAY_PROBE)):
            o = H.default_options(assay_format=fmt, **kw)
            for codes in db:
                for a in assays:
                    x, y = ref.search(codes, a[0], a[1], a[2], o), oracle.search(codes, a[0], a[1], a[2], o)
                    assert [(h.exact_key(), h.floats()) for h in x] == [(h.exact_key(), h.floats()) for h in y], (case, fmt)
                    total += len(x)
    assert total > 200


def test_overhanging_probe_sites_vs_compiled_reference(oracle, ref):
    """Probe sites cut by a fragment end keep coordinates beyond the fragment and their text is read from the
    clamped end for the full length (probe_search.cpp:129-142, :205-219): the oracle's restatement of that quirk
    against the compiled reference (the GPU test of the same name compares the engine with the oracle)."""
    rng = np.random.default_rng(31337)
    P = gen.rand_oligo(38, rng)
    site = gen.revcomp(P)
    n_over = 0
    for strand_text in (site, P):
        n = 9000
        codes = gen.random_codes(n, rng)
        gen.plant(codes, n - 26, strand_text[:26])
        gen.plant(codes, 0, strand_text[10:])
        gen.plant(codes, 4000, strand_text)
        o = H.default_options(assay_format=H.ASSAY_PROBE, min_probe_tm=30.0)
        x, y = ref.search(codes, None, None, P, o), oracle.search(codes, None, None, P, o)
        assert [(h.exact_key(), h.floats()) for h in x] == [(h.exact_key(), h.floats()) for h in y]
        n_over += sum(1 for h in x if h.probe_first < 0 or h.probe_last >= n)
    assert n_over >= 2


@pytest.mark.parametrize("W", [4, 5, 6, 8])
def test_word_sizes_vs_compiled_reference(oracle, ref, W):
    """Hash word sizes other than 7: seeds and searches of the oracle against the compiled reference
    (W = 3 is equal as well; with 64 keys nearly every position is a seed and the case runs for minutes)."""
    rng = np.random.default_rng(1000 + W)
    n = 12000 if W <= 5 else 60000
    db = [gen.random_codes(n, rng), gen.random_codes(n // 2 + 37, rng), gen.random_codes(6145, rng)]
    gen.sprinkle_degenerate(db[0], rng, frac=1e-3, n_runs_per_50kb=4)
    assays = gen.make_assays(rng, db, 3, "taqman", variants=3)
    o = H.default_options(min_primer_tm=42.0, min_probe_tm=42.0, word_size=W)
    nseeds = nhits = 0
    for codes in db:
        for oligo in (assays[0][0], assays[0][0][:9] + "N" + assays[0][0][10:]):
            for plus in (False, True):
                a, b = ref.seeds(codes, oligo, W, plus, unique=True), oracle.seeds(codes, oligo, W, plus, unique=True)
                assert a == b, (W, oligo, plus)
                nseeds += len(a)
        for a in assays:
            x, y = ref.search(codes, a[0], a[1], a[2], o), oracle.search(codes, a[0], a[1], a[2], o)
            assert [(h.exact_key(), h.floats()) for h in x] == [(h.exact_key(), h.floats()) for h in y], W
            nhits += len(x)
    assert nseeds > 20 and nhits >= 3
