"""Golden hairpin vectors from the compiled reference (oracle/_ref/libtntref.so: ref_hairpin =
NucCruc::set_duplex + approximate_tm_hairpin).  python tests/golden/make_hairpins.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gen  # noqa: E402
import harness as H  # noqa: E402


def cases():
    rng = np.random.default_rng(2718)
    out = ["GGGGCGAAAGCCCC", "CCGGAGTTTCCGGGTCTAATT", "AAAAAAAAAAAAAAAAAAAA", "ACGT", "ACGTA", "GCGCGAAGCGC", "ATATATATATATATAT",
           "GGGGGGGGGGCCCCCCCCCC", "TTCCCCTTTGGAGGCATC", "GGAGGATCCAACAGCAAGG", "CGATGGGCTTTCAGGACAGGTGT"]
    for it in range(240):
        L = int(rng.integers(5, 57))
        q = gen.rand_oligo(L, rng)
        if it % 3 == 0 and L >= 14:
            stem = gen.rand_oligo(int(rng.integers(3, 9)), rng)
            loop = gen.rand_oligo(int(rng.integers(3, 8)), rng)
            q = (gen.rand_oligo(int(rng.integers(0, 4)), rng) + stem + loop + gen.mutate(gen.revcomp(stem), int(rng.integers(0, 2)), rng)
                 + gen.rand_oligo(int(rng.integers(0, 4)), rng))[:56]
        if it % 9 == 0:
            q = list(q)
            q[int(rng.integers(0, len(q)))] = "I"
            q = "".join(q)
        out.append(q)
    return out


def rec(o):
    f = lambda x: float(np.float32(x)).hex()
    return {"valid": int(o.valid), "tm": f(o.tm), "dH": f(o.dH), "dS": f(o.dS), "dp_dg": f(o.dp_dg),
            "loop": [o.q_first, o.t_first], "open_end": [o.q_last, o.t_last], "columns": o.num_gap}


if __name__ == "__main__":
    ref = H.ref()
    out = []
    for i, q in enumerate(cases()):
        T, na = [(310.15, 0.05), (298.15, 0.2), (333.15, 0.01)][i % 3]
        out.append({"q": q, "T": T, "na": na, "out": rec(ref.hairpin(q, T, na))})
    json.dump(out, open(os.path.join(HERE, "hairpins.json"), "w"), indent=0)
    print(len(out), "cases,", sum(r["out"]["valid"] for r in out), "with a hairpin")
