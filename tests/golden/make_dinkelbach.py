"""Golden vectors of the Dinkelbach mode (NucCruc::dinkelbach(true), nuc_cruc.cpp:2399-2440, :2459-2500,
:2548-2588) from the compiled reference (oracle/_ref/libtntref.so with ref_set_dinkelbach(1)): single
windows, oligo dimers, hairpins and whole searches.  python tests/golden/make_dinkelbach.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gen  # noqa: E402
import harness as H  # noqa: E402


def f32(x):
    return float(np.float32(x)).hex()


def align_rec(a):
    return {"tm": f32(a.tm), "dH": f32(a.dH), "dS": f32(a.dS), "dG": f32(a.dG), "valid": int(a.valid),
            "ints": [a.anchor5, a.anchor3, a.num_mismatch, a.num_gap, a.max_poly_degen,
                     a.q_first, a.q_last, a.t_first, a.t_last] if a.valid else [],
            "alignment": a.alignment.decode() if a.valid else ""}


def struct_rec(a):
    return {"tm": f32(a.tm), "dH": f32(a.dH), "dS": f32(a.dS), "valid": int(a.valid)}


def window_cases():
    rng = np.random.default_rng(1618)
    out = []
    for it in range(400):
        L = int(rng.integers(12, 57))
        q = gen.rand_oligo(L, rng)
        if it % 11 == 0:
            q = list(q)
            q[int(rng.integers(0, L))] = "IRYN"[it % 4]
            q = "".join(q)
        plain = q.replace("I", "A").replace("R", "A").replace("Y", "C").replace("N", "T")
        t = gen.mutate(gen.revcomp(plain), int(rng.integers(0, 6)), rng)
        t = gen.rand_oligo(4, rng) + t + gen.rand_oligo(4, rng)
        if it % 5 == 0:
            t = gen.rand_oligo(L + 8, rng)
        if it % 13 == 0:
            t = list(t)
            t[int(rng.integers(0, len(t)))] = "RYKMN"[it % 5]
            t = "".join(t)
        T, na = [(310.15, 0.05), (298.15, 0.2), (333.15, 0.01)][it % 3]
        out.append((q, t[:64], T, na, [9.0e-7, 2.5e-7][it % 2]))
    return out


def search_cases():
    rng = np.random.default_rng(577)
    out = []
    for kind in ("pcr", "taqman", "probe", "padlock"):
        db = [gen.random_codes(20000, rng) for _ in range(2)]
        assays = gen.make_assays(rng, db, 3, kind, variants=3)
        out.append((kind, db, assays))
    return out


def search_options_for(kind):
    o = H.default_options(min_primer_tm=40.0, min_probe_tm=40.0)
    if kind == "probe":
        o.assay_format = 1
    elif kind == "padlock":
        o.assay_format = 2
    return o


def hit_rec(h):
    return {"key": [int(x) if not isinstance(x, (bytes, str)) else (x.decode() if isinstance(x, bytes) else x) for x in h.exact_key()],
            "floats": [f32(x) for x in h.floats()]}


if __name__ == "__main__":
    ref = H.ref()
    ref.set_dinkelbach(True)
    NB = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'I': 4, 'M': 7, 'R': 8, 'S': 9, 'V': 10, 'W': 11, 'Y': 12, 'H': 13,
          'K': 14, 'D': 15, 'B': 16, 'N': 17}
    out = {"windows": [], "dimers": [], "hairpins": [], "searches": []}
    for q, t, T, na, ct in window_cases():
        tb = np.array([NB[c] for c in t], dtype=np.uint8)
        out["windows"].append({"q": q, "t": t, "T": T, "na": na, "ct": ct, "out": align_rec(ref.align(q, tb, T=T, na=na, ct=ct))})
    rng = np.random.default_rng(99)
    for it in range(150):
        q = gen.rand_oligo(int(rng.integers(10, 40)), rng)
        if it % 3 == 0:
            h = gen.rand_oligo(int(rng.integers(4, 10)), rng)
            q = (h + gen.rand_oligo(int(rng.integers(3, 7)), rng) + gen.revcomp(h) + gen.rand_oligo(2, rng))[:56]
        t = gen.rand_oligo(int(rng.integers(10, 40)), rng) if it % 2 else None
        if it % 6 == 1:
            t = gen.mutate(gen.revcomp(q), int(rng.integers(0, 3)), rng)[:64]
        out["dimers"].append({"q": q, "t": t, "out": struct_rec(ref.dimer(q, t))})
        out["hairpins"].append({"q": q, "out": struct_rec(ref.hairpin(q))})
    for kind, db, assays in search_cases():
        o = search_options_for(kind)
        for t, codes in enumerate(db):
            for i, a in enumerate(assays):
                hits = ref.search(codes, a[0], a[1], a[2], o)
                out["searches"].append({"kind": kind, "target": t, "assay": i, "hits": [hit_rec(h) for h in hits]})
    ref.set_dinkelbach(False)
    json.dump(out, open(os.path.join(HERE, "dinkelbach.json"), "w"), indent=0)
    print({k: len(v) for k, v in out.items()}, "valid windows", sum(r["out"]["valid"] for r in out["windows"]),
          "hits", sum(len(s["hits"]) for s in out["searches"]))
