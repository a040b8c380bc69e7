#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libtntref.so, built from
/root/reference by oracle/Makefile).  Needs this container (the reference sources); the JSON files
are committed and are what pins oracle/tnt_oracle.c on machines without /root/reference.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import gen  # noqa: E402
import harness as H  # noqa: E402

NB = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'I': 4, 'M': 7, 'R': 8, 'S': 9, 'V': 10, 'W': 11, 'Y': 12, 'H': 13,
      'K': 14, 'D': 15, 'B': 16, 'N': 17}


def f32(x):
    return float(np.float32(x)).hex()


def align_rec(a):
    return {"tm": f32(a.tm), "dH": f32(a.dH), "dS": f32(a.dS), "dG": f32(a.dG), "valid": a.valid,
            "ints": [a.anchor5, a.anchor3, a.num_mismatch, a.num_gap, a.max_poly_degen,
                     a.q_first, a.q_last, a.t_first, a.t_last, a.target_start, a.target_stop, a.loc_5, a.loc_3],
            "alignment": a.alignment.decode()}


def hit_rec(h):
    return {"key": [x.decode() if isinstance(x, bytes) else x for x in h.exact_key()],
            "floats": [f32(x) for x in h.floats()]}


def main():
    r = H.ref()
    rng = np.random.default_rng(20261017)

    # 1. parameter tables
    tables = []
    for T, na in [(310.15, 0.05), (298.15, 0.2), (333.15, 0.01)]:
        t = r.dump_tables(T, na)
        tables.append({"T": T, "na": na, "delta_g": list(t.delta_g)})
    json.dump(tables, open(os.path.join(HERE, "tables.json"), "w"))

    # 2. single alignments: random, planted + edits, degenerate, dangling ends
    aligns = []
    for it in range(400):
        L = int(rng.integers(12, 40))
        q = gen.rand_oligo(L, rng)
        mode = it % 4
        if mode == 0:
            t = gen.rand_oligo(L + 8, rng)
        else:
            t = gen.rand_oligo(4, rng) + gen.mutate(gen.revcomp(q), int(rng.integers(0, 5)), rng) + gen.rand_oligo(4, rng)
        if mode == 2:
            ql = list(q)
            ql[int(rng.integers(0, L))] = "IMRSVWYHKDBN"[int(rng.integers(0, 12))]
            q = "".join(ql)
            tl = list(t)
            for _ in range(int(rng.integers(0, 4))):
                tl[int(rng.integers(0, len(tl)))] = "MRSVWYHKDBNI"[int(rng.integers(0, 12))]
            t = "".join(tl)
        d5, d3 = (int(rng.integers(0, 2)), int(rng.integers(0, 2))) if mode == 3 else (0, 0)
        T, na, ct = [(310.15, 0.05, 9.0e-7), (310.15, 0.05, 2.5e-7), (320.0, 0.1, 1.0e-6)][it % 3]
        tb = np.array([NB[c] for c in t], dtype=np.uint8)
        a = r.align(q, tb, T=T, na=na, ct=ct, dangle5=d5, dangle3=d3)
        aligns.append({"q": q, "t": t, "T": T, "na": na, "ct": ct, "d5": d5, "d3": d3, "out": align_rec(a)})
    json.dump(aligns, open(os.path.join(HERE, "alignments.json"), "w"))

    # 3. seeds + bind windows on small fragments
    seeds = []
    for it in range(12):
        n = int(rng.integers(2000, 9000))
        codes = gen.random_codes(n, rng)
        if it % 3 == 0:
            gen.sprinkle_degenerate(codes, rng, frac=1e-2, n_runs_per_50kb=50)
        ol = gen.rand_oligo(int(rng.integers(8, 30)), rng)
        if it % 4 == 1:
            o = list(ol)
            o[len(o) // 2] = "I"
            ol = "".join(o)
        gen.plant(codes, 100, ol)
        gen.plant(codes, 900, gen.revcomp(ol))
        W = [7, 7, 6, 8][it % 4]
        rec = {"codes": gen.codes_to_str(codes), "oligo": ol, "W": W}
        for plus in (0, 1):
            rec["raw%d" % plus] = r.seeds(codes, ol, W, bool(plus), unique=False)
            rec["uniq%d" % plus] = r.seeds(codes, ol, W, bool(plus), unique=True)
            if W == 7:
                rec["bind%d" % plus] = [align_rec(r.bind_window(codes, ol, bool(plus), q, t)) for (q, t) in rec["uniq%d" % plus][:40]]
        seeds.append(rec)
    json.dump(seeds, open(os.path.join(HERE, "seeds.json"), "w"))

    # 4. end-to-end searches
    searches = []
    for it in range(24):
        kind = ["pcr", "taqman", "probe", "padlock"][it % 4]
        n = int(rng.integers(8000, 20000))
        if kind in ("pcr", "taqman"):
            codes, F, R, P = gen.make_pcr_case(rng, n, n_sites=int(rng.integers(1, 5)), probe=(kind == "taqman"))
            o = H.default_options(min_primer_tm=float(rng.choice([35.0, 45.0])), min_probe_tm=40.0,
                                  single_primer_pcr=int(rng.integers(0, 2)), max_len=int(rng.choice([600, 2000])))
        else:
            db = [gen.random_codes(n, rng)]
            (F, R, P), = gen.make_assays(rng, db, 1, kind, variants=3)
            codes = db[0]
            if kind == "probe":
                o = H.default_options(assay_format=H.ASSAY_PROBE, min_probe_tm=35.0)
            else:
                o = H.default_options(assay_format=[H.ASSAY_PADLOCK, H.ASSAY_MIPS][it // 4 % 2], min_probe_tm=30.0, max_len=20)
        if it % 6 == 5:
            gen.sprinkle_degenerate(codes, rng, frac=3e-3, n_runs_per_50kb=10)
        hits = r.search(codes, F, R, P, o)
        searches.append({"kind": kind, "codes": gen.codes_to_str(codes), "F": F, "R": R, "P": P,
                         "opts": {k: getattr(o, k) for k, _ in o._fields_},
                         "hits": [hit_rec(h) for h in hits]})
    json.dump(searches, open(os.path.join(HERE, "searches.json"), "w"))
    print("hits in fixtures:", sum(len(s["hits"]) for s in searches))

    # 5. FASTA reader + fragment queue (sequence_data on a file; written separately so that the
    #    fixtures above need not be regenerated: `make_golden.py fasta`)
    make_fasta(r)
    make_dimers(r)


def dimer_cases():
    rng = np.random.default_rng(4711)
    cases = []
    for it in range(120):
        L = int(rng.integers(10, 40))
        q = gen.rand_oligo(L, rng)
        if it % 5 == 0:                      # self-complementary
            h = gen.rand_oligo(L // 2, rng)
            q = h + gen.revcomp(h)
        t = gen.rand_oligo(int(rng.integers(10, 40)), rng) if it % 2 else None
        if it % 7 == 0 and t:
            t = gen.mutate(gen.revcomp(q), 2, rng)   # a real primer dimer
        if it % 11 == 0:
            q = q[:5] + "I" + q[6:]
        ca, cb = (9e-7, 9e-7) if it % 3 else (2e-6, 5e-7)
        cases.append((q, t, ca, cb))
    return cases


def make_dimers(r):
    out = []
    for q, t, ca, cb in dimer_cases():
        a = r.dimer(q, t, conc_a=ca, conc_b=cb)
        out.append({"q": q, "t": t, "ca": ca, "cb": cb, "tm": f32(a.tm), "dH": f32(a.dH), "dS": f32(a.dS), "dG": f32(a.dG),
                    "valid": a.valid, "alignment": a.alignment.decode() if t else None})
    json.dump(out, open(os.path.join(HERE, "dimers.json"), "w"))
    print("dimer fixtures:", len(out))


def fasta_texts():
    rng = np.random.default_rng(777)
    texts = list(gen.FASTA_EDGE_CASES)
    for it in range(6):
        texts.append(gen.rand_fasta(rng, n_records=int(rng.integers(1, 6)), max_len=900, width=[60, 70, 0][it % 3],
                                    crlf=bool(it % 2)))
    return texts


def make_fasta(r):
    import base64
    import tempfile
    out = []
    for text in fasta_texts():
        with tempfile.NamedTemporaryFile(suffix=".fna", delete=False) as f:
            f.write(text)
        try:
            for threshold, overlap in [(0, 0), (100, 12)]:
                try:
                    recs = r.fasta_records(text, f.name, threshold=threshold, overlap=overlap)
                    out.append({"text": base64.b64encode(text).decode(), "threshold": threshold, "overlap": overlap,
                                "records": [{"approx_len": a, "defline": d, "codes": gen.codes_to_str(w),
                                             "pieces": [[s0, s1, gen.codes_to_str(c)] for s0, s1, c in pcs]}
                                            for _, a, d, w, pcs in recs]})
                except RuntimeError as ex:
                    out.append({"text": base64.b64encode(text).decode(), "threshold": threshold, "overlap": overlap,
                                "error": str(ex)})
        finally:
            os.unlink(f.name)
    json.dump(out, open(os.path.join(HERE, "fasta.json"), "w"))
    print("fasta fixtures:", len(out), "errors:", sum("error" in o for o in out))


if __name__ == "__main__":
    if sys.argv[1:] == ["fasta"]:
        make_fasta(H.ref())
    elif sys.argv[1:] == ["dimers"]:
        make_dimers(H.ref())
    else:
        main()
