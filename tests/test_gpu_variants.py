"""Paths of the kernels that ordinary inputs never reach, exercised with a variant build.

More than MAX_MAXCELLS (64) DP cells tied for the maximum need low-complexity sequence of a length
the 56-base oligo limit hardly allows; a library built with -DTNT_MAX_MAXCELLS=2 takes the same code
path (hand-over of the fast tiers to the generic kernel, chunked enumeration there) for every window
with three tied cells.  The variant is compiled on the spot (nvcc is part of the image) and used by
a child process through TNT_LIB."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import json, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
    import numpy as np, gen, harness as H
    from thermonucleotideblast_b200 import Engine
    orc = H.oracle()
    rng = np.random.default_rng(7)
    out = {"windows": 0, "tied": 0, "bad": []}
    oligos = ["ACACACACACACACACACAC", "ATATATATATATATATATATAT", "GGGGGGGGGGGGGGGGGGGG", "ACGACGACGACGACGACGACGACG",
              "AAAAAAAAAACCCCCCCCCC", gen.rand_oligo(22, rng)]
    with Engine() as e:
        for ol in oligos:
            unit = gen.revcomp(ol)
            codes = gen.random_codes(6000, rng)
            # tandem copies of the site and of its shifted / truncated forms: many co-optimal cells
            text = (unit * 6) + "ACGT" + unit[3:] + unit[:9] + "TTGCA" + (ol * 4)
            gen.plant(codes, 1000, text)
            tid = e.add_target(codes)
            for plus in (False, True):
                seeds = orc.seeds(codes, ol, 7, plus, unique=True)
                got = e.align(tid, ol, plus, seeds)
                for (q, t), g in zip(seeds, got):
                    w = orc.bind_window(codes, ol, plus, q, t)
                    out["windows"] += 1
                    if (g.valid, g.alignment.decode() if g.valid else "") != (w.valid, w.alignment.decode() if w.valid else "") or \\
                            (w.valid and (g.tm, g.dH, g.dS, g.loc_5, g.loc_3) != (w.tm, w.dH, w.dS, w.loc_5, w.loc_3)):
                        out["bad"].append((ol, plus, q, t))
            e.clear_targets()
    print(json.dumps(out))
""")


def test_more_tied_cells_than_a_chunk_holds(tmp_path):
    from thermonucleotideblast_b200 import build as b
    lib = str(tmp_path / "libtntb200_ties.so")
    b.build(out=lib, extra=["-DTNT_MAX_MAXCELLS=2"])
    env = dict(os.environ, TNT_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests")}],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["windows"] > 200 and out["bad"] == [], out["bad"][:5]
