"""Golden cases for the reference's cull anomaly (amplicon_search.cpp:679-765).

Found by diffing the text output of the engine-backed program against the reference binary on the
config-5 slice (1000 PCR assays x 50 Mbp, seeds 5 / 55, tests/test_gpu_shim.py): in two (fragment,
assay) groups a primer binds both strands of a palindromic site; the two bound sites overlap, their
order by loc_5 differs from their order by seed position, and the reference's cull loses the site
that closes a second amplicon.  This script cuts the neighbourhoods out of the same synthetic data
and records what the compiled reference (oracle/_ref/libtntref.so) reports for them.

    python tests/golden/make_cull_cases.py        # needs oracle/_ref (built where /root/reference exists)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import gen  # noqa: E402
import harness as H  # noqa: E402


def main():
    n = 50_000_000
    rng = np.random.default_rng(5)
    records = [gen.random_codes(min(5_000_000, n - i), rng) for i in range(0, n, 5_000_000)]
    assays = gen.make_assays(np.random.default_rng(55), records, 1000, "pcr", lens=(20, 21, 25), amp=(80, 400), variants=1)
    ref = H.ref()
    o = H.default_options(min_primer_tm=45.0)
    cases = []
    # (record, assay, position of the amplicon the reference loses, its end)
    for rec, a, lost_first, lost_last in ((6, 49, 4982607, 4983775), (8, 498, 519402, 519735)):
        lo = max(0, lost_first - 8000)
        hi = min(len(records[rec]), lost_last + 8000)
        codes = records[rec][lo:hi].copy()
        F, R, P = assays[a]
        hits = ref.search(codes, F, R, P, o)
        got = [(h.amp_first, h.amp_last) for h in hits]
        assert (lost_first - lo, lost_last - lo) not in got, "the anomaly does not reproduce in this window"
        cases.append({
            "record": rec, "assay": a, "window_start": lo, "codes": gen.codes_to_str(codes),
            "forward": F, "reverse": R, "min_primer_tm": 45.0,
            "lost_amplicon": [lost_first - lo, lost_last - lo],
            "reference_hits": [{"amp_first": h.amp_first, "amp_last": h.amp_last, "primer_strand": h.primer_strand,
                                "forward_tm": float(np.float32(h.forward_tm)), "reverse_tm": float(np.float32(h.reverse_tm)),
                                "forward_align": h.forward_align.decode(), "reverse_align": h.reverse_align.decode()} for h in hits],
        })
        print("record %d assay %d: reference reports %s, loses %s" % (rec, a, got, cases[-1]["lost_amplicon"]))
    with open(os.path.join(HERE, "cull_cases.json"), "w") as f:
        json.dump(cases, f, indent=1)


if __name__ == "__main__":
    main()
