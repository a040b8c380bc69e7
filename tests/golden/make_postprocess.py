"""Golden answers for tests/test_postprocess.py from the compiled reference (oracle/_ref/libtntref.so):
the order in which select_best_match / uniquify_results / sort leave the hit lists of four synthetic
cases.  Run where /root/reference exists:  python tests/golden/make_postprocess.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness as H  # noqa: E402
import test_postprocess as T  # noqa: E402

out = []
ref = H.ref()
for seed, kind in T.CASES:
    records, assays, frags, hits = T.make_case(seed, kind)
    for best, uniq in ((False, True), (True, True), (False, False), (True, False)):
        order = T.reference_order(ref, assays, frags, hits, best, uniq)
        out.append({"seed": seed, "kind": kind, "best_match": best, "uniquify": uniq, "order": order})
        print(seed, kind, best, uniq, len(order), "of", sum(len(p) for p in hits))
json.dump(out, open(os.path.join(HERE, "postprocess.json"), "w"))
