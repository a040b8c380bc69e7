"""The Dinkelbach mode of the engine (`tnt_engine_params.dinkelbach`, the reference's `--dinkelbach T`:
NucCruc::approximate_tm_heterodimer / _homodimer / _hairpin with use_dinkelbach, nuc_cruc.cpp:2399-2440,
:2459-2500, :2548-2588) against the oracle in the same mode: every window is aligned at 0 degC and then
again at the Tm of the previous pass (penalty table re-derived per pass, per window) until dH - T*dS stops
rising.  Same bar as the default mode: records and hit lists identical, floats bit for bit."""
import json
import os

import numpy as np
import pytest

import gen
import harness as H
from test_gpu_parity import TM_TOL, _compare_align, assert_hits_equal, to_opts

pytestmark = pytest.mark.gpu


@pytest.fixture
def dink(engine_lib, oracle):
    from thermonucleotideblast_b200 import Engine
    oracle.set_dinkelbach(True)
    e = Engine(dinkelbach=True)
    yield e, oracle
    e.close()
    oracle.set_dinkelbach(False)


def test_windows(dink):
    eng, oracle = dink
    rng = np.random.default_rng(2024)
    n = 60000
    codes = gen.random_codes(n, rng)
    oligos = [gen.rand_oligo(L, rng) for L in (14, 18, 20, 23, 27, 34, 48, 56)]
    for i, ol in enumerate(oligos):
        for k in range(10):
            gen.plant(codes, 1500 + (i * 10 + k) * 450, gen.mutate(gen.revcomp(ol) if k % 2 else ol, k % 5, rng))
    gen.plant(codes, 0, gen.revcomp(oligos[2])[3:])
    gen.plant(codes, n - 14, oligos[3][:14])
    gen.sprinkle_degenerate(codes, rng, frac=2e-3, n_runs_per_50kb=4)
    tid = eng.add_target(codes)
    nvalid = 0
    for ol in oligos + [oligos[1][:8] + "I" + oligos[1][9:], "R" + oligos[3][1:]]:
        for plus in (False, True):
            seeds = oracle.seeds(codes, ol, 7, plus, unique=True)
            nvalid += _compare_align(eng, oracle, tid, codes, ol, plus, seeds)
    assert nvalid > 1000


def test_golden_windows_and_structures(dink):
    """The committed vectors of the compiled reference (tests/golden/make_dinkelbach.py)."""
    eng, oracle = dink
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dinkelbach.json")))
    f32 = lambda x: float(np.float32(x)).hex()
    n = 0
    for r in gold["dimers"]:
        g = eng.oligo_dimer(r["q"], r["t"])
        assert (f32(g.tm), f32(g.dH), f32(g.dS), int(g.valid)) == (r["out"]["tm"], r["out"]["dH"], r["out"]["dS"], r["out"]["valid"]), (r["q"], r["t"])
        n += int(g.valid)
    for r in gold["hairpins"]:
        g = eng.oligo_hairpin(r["q"])
        assert (f32(g.tm), f32(g.dH), f32(g.dS), int(g.valid)) == (r["out"]["tm"], r["out"]["dH"], r["out"]["dS"], r["out"]["valid"]), r["q"]
        n += int(g.valid)
    assert n > 150


@pytest.mark.parametrize("kind", ["pcr", "taqman", "probe", "padlock"])
def test_searches(dink, kind):
    from thermonucleotideblast_b200 import Assay
    eng, oracle = dink
    rng = np.random.default_rng({"pcr": 1, "taqman": 2, "probe": 3, "padlock": 4}[kind])
    db = [gen.random_codes(int(rng.integers(15000, 30000)), rng) for _ in range(3)]
    assays = gen.make_assays(rng, db, 5, kind, variants=3)
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    if kind == "probe":
        o.assay_format = 1
    elif kind == "padlock":
        o.assay_format = 2
    want = {(t, i): oracle.search(codes, a[0], a[1], a[2], o) for t, codes in enumerate(db) for i, a in enumerate(assays)}
    assert sum(len(v) for v in want.values()) >= 5
    for c in db:
        eng.add_target(c)
    eng.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
    got = eng.search(to_opts(o))
    n = 0
    for (t, i), w in want.items():
        mine = [h for h in got if h.target_id == t and h.assay_index == i]
        assert_hits_equal(eng, mine, w, assays[i])
        n += len(mine)
    assert n == len(got)


def test_assay_structures(dink):
    from thermonucleotideblast_b200 import Assay, search_options
    eng, oracle = dink
    rng = np.random.default_rng(818)
    assays = []
    for i in range(20):
        F, R, P = gen.rand_oligo(int(rng.integers(16, 31)), rng), gen.rand_oligo(int(rng.integers(16, 31)), rng), gen.rand_oligo(int(rng.integers(18, 36)), rng)
        if i % 4 == 0:
            h = gen.rand_oligo(9, rng)
            F = h + "GAAA" + gen.revcomp(h)
        if i % 5 == 0:
            R = gen.revcomp(F)[:len(F) - 2]
        assays.append(Assay(i, F, R, P))
    eng.set_assays(assays)
    o = search_options(min_primer_tm=45.0, min_probe_tm=50.0)
    got = eng.assay_structures(o, len(assays))
    fps, rps, ps = o.forward_primer_strand, o.reverse_primer_strand, o.probe_strand
    ne = 0
    for a, g in zip(assays, got):
        want_h = [oracle.hairpin(x).tm for x in (a.forward, a.reverse, a.probe)]
        want_d = [oracle.dimer(x, None, conc_a=c, conc_b=c).tm for x, c in ((a.forward, fps), (a.reverse, rps), (a.probe, ps))]
        want_x = [oracle.dimer(a.forward, a.reverse, conc_a=fps, conc_b=rps).tm, oracle.dimer(a.forward, a.forward, conc_a=fps, conc_b=rps).tm,
                  oracle.dimer(a.reverse, a.reverse, conc_a=fps, conc_b=rps).tm]
        assert list(g.hairpin_tm) == [np.float32(x) for x in want_h], a
        assert list(g.homodimer_tm) == [np.float32(x) for x in want_d], a
        assert list(g.heterodimer_tm) == [np.float32(x) for x in want_x], a
        assert all(abs(float(x) - float(y)) <= TM_TOL for x, y in zip(list(g.hairpin_tm), want_h))
        ne += sum(1 for x in want_h + want_d + want_x if x > 0)
    assert ne > 30
