"""Randomised differential of the engine against the oracle (tests/harness.py): many small searches with
random assay formats, word sizes, temperatures, salt, bounds, clamps and strand settings.  Not a test of the
suite (those use fixed seeds); a wider net run by hand on a GPU box:
    python tests/fuzz_parity.py [seconds] [seed]
Prints one line per mismatch and a summary; exit code 1 when anything differed."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

import gen  # noqa: E402
import harness as H  # noqa: E402
from test_gpu_parity import assert_hits_equal, to_opts  # noqa: E402
from thermonucleotideblast_b200 import Assay, Engine  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
oracle = H.oracle()
t_end = time.time() + budget
ncase = nhits = nbad = 0
it = int(os.environ.get("FUZZ_FROM", "1")) - 1
while time.time() < t_end:
    it += 1
    rng = np.random.default_rng(seed0 * 100003 + it)
    kind = ["pcr", "taqman", "probe", "padlock"][int(rng.integers(0, 4))]
    W = int(rng.choice([4, 5, 6, 7, 7, 7, 8]))
    # kernel choices the engine normally makes itself (test hooks read at set-up time)
    for k in ("TNT_SCAN_MODE", "TNT_SCAN_GLOBAL_TABLE", "TNT_NO_LEAN", "TNT_NO_LEAN_SKIP"):
        os.environ.pop(k, None)
    pick = int(rng.integers(0, 8))
    if pick == 0:
        os.environ["TNT_SCAN_MODE"] = "sparse"
    elif pick == 1:
        os.environ["TNT_SCAN_MODE"] = "dense"
    elif pick == 2:
        os.environ["TNT_SCAN_MODE"] = "dense"
        os.environ["TNT_SCAN_GLOBAL_TABLE"] = "1"
    elif pick == 3:
        os.environ["TNT_NO_LEAN"] = "1"
    elif pick == 4:
        os.environ["TNT_NO_LEAN_SKIP"] = "1"
    T = float(rng.choice([310.15, 310.15, 300.15, 325.15]))
    na = float(rng.choice([0.05, 0.05, 0.2, 1.0]))
    dink = bool(rng.integers(0, 8) == 0)
    d5, d3 = (int(rng.integers(0, 2)), int(rng.integers(0, 2))) if rng.integers(0, 5) == 0 else (0, 0)
    lens = (int(rng.integers(14, 31)), int(rng.integers(14, 31)), int(rng.integers(16, 41)))
    if rng.integers(0, 6) == 0:
        lens = (int(rng.integers(14, 55)), int(rng.integers(14, 55)), int(rng.integers(14, 55)))   # make_assays varies each by +-2
    top = 150000 if rng.integers(0, 10) == 0 else 40000
    db = [gen.random_codes(int(rng.integers(4000, top)), rng) for _ in range(int(rng.integers(1, 4)))]
    if rng.integers(0, 3) == 0:
        gen.sprinkle_degenerate(db[0], rng, frac=float(rng.choice([5e-4, 5e-3])), n_runs_per_50kb=int(rng.integers(0, 8)))
    assays = gen.make_assays(rng, db, int(rng.integers(1, 5)), kind, lens=lens, variants=int(rng.integers(1, 5)))
    if rng.integers(0, 5) == 0:
        # degenerate letters in the oligos (the planted sites keep the concrete base)
        def degen(ol):
            if ol is None:
                return None
            ol = list(ol)
            for _ in range(int(rng.integers(1, 3))):
                i = int(rng.integers(0, len(ol)))
                ol[i] = "I" if rng.integers(0, 2) else {"A": "R", "G": "R", "C": "Y", "T": "Y"}[ol[i]] if ol[i] in "ACGT" else ol[i]
            return "".join(ol)
        assays = [tuple(degen(x) for x in a) for a in assays]
    kw = dict(word_size=W, target_T=T, salt=na, dangle5=d5, dangle3=d3,
              min_primer_tm=float(rng.choice([0.0, 30.0, 38.0, 45.0])), min_probe_tm=float(rng.choice([0.0, 30.0, 40.0])),
              max_len=int(rng.choice([300, 2000])), single_primer_pcr=int(rng.integers(0, 2)),
              primer_clamp=int(rng.integers(0, 4)), min_max_primer_clamp=int(rng.choice([-1, -1, 3])),
              probe_clamp_5=int(rng.integers(0, 3)), probe_clamp_3=int(rng.integers(0, 3)),
              max_gap=int(rng.choice([999, 999, 0, 1])), max_mismatch=int(rng.choice([999, 999, 2])),
              target_strand=int(rng.choice([3, 3, 1, 2])))
    if kw["min_primer_tm"] == 0.0:
        kw["max_primer_dg"] = float(rng.choice([-6.0, -9.0]))
    if kw["min_probe_tm"] == 0.0:
        kw["max_probe_dg"] = float(rng.choice([-6.0, -9.0]))
    if kind == "probe":
        kw["assay_format"] = H.ASSAY_PROBE
    elif kind == "padlock":
        kw["assay_format"] = int(rng.choice([H.ASSAY_PADLOCK, H.ASSAY_MIPS]))
        kw["max_len"] = int(rng.choice([0, 3, 50]))
    o = H.default_options(**kw)
    if os.environ.get("FUZZ_VERBOSE"):
        print("it=%d kind=%s W=%d T=%g na=%g dink=%d d5=%d d3=%d db=%s lens=%s nassay=%d env=%s opts=%s" % (it, kind, W, T, na, dink, d5, d3, [len(c) for c in db], lens, len(assays), {k: os.environ[k] for k in os.environ if k.startswith("TNT_")}, kw), flush=True)
    oracle.set_dinkelbach(dink)
    try:
        with Engine(target_T=T, salt=na, dangle5=bool(d5), dangle3=bool(d3), word_size=W, dinkelbach=dink) as e:
            for c in db:
                e.add_target(c)
            e.set_assays([Assay(i, *a) for i, a in enumerate(assays)])
            got = e.search(to_opts(o))
            for t, codes in enumerate(db):
                for i, a in enumerate(assays):
                    want = oracle.search(codes, a[0], a[1], a[2], o)
                    mine = [h for h in got if h.target_id == t and h.assay_index == i]
                    ncase += 1
                    nhits += len(want)
                    try:
                        assert_hits_equal(e, mine, want, a, T=T)
                    except AssertionError as ex:
                        nbad += 1
                        print("MISMATCH it=%d kind=%s W=%d T=%g na=%g dink=%d d5=%d d3=%d target=%d assay=%s opts=%s: %s"
                              % (it, kind, W, T, na, dink, d5, d3, t, a, kw, str(ex)[:300]), flush=True)
    finally:
        oracle.set_dinkelbach(False)
print("fuzz: %d iterations, %d (fragment, assay) searches, %d hits compared, %d mismatches" % (it, ncase, nhits, nbad))
sys.exit(1 if nbad else 0)
