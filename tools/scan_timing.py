"""Seed-scan timing of the bench workload (tnt_engine_scan_only: all fragments, stage-1 oligo strands, candidates
counted but not stored).  `python tools/scan_timing.py [Mbp] [assays]`; TNT_LIB selects a variant library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nassay = int(sys.argv[2]) if len(sys.argv) > 2 else 100
records, fragments, assays, db_bases = bench.build_workload(0, mbp, nassay, pinned=True)
opts = search_options(min_primer_tm=bench.MIN_PRIMER_TM, min_probe_tm=bench.MIN_PROBE_TM, max_len=bench.MAX_LEN)
with Engine() as e:
    e.set_assays([Assay(i, a[0], a[1], a[2]) for i, a in enumerate(assays)])
    e.add_targets(FragmentList(fragments))
    best = None
    for _ in range(5):
        n, ms = e.scan_only(opts)
        best = ms if best is None else min(best, ms)
    print("lib", os.environ.get("TNT_LIB", "default"), "candidates", n, "scan ms (best of 5)", round(best, 3))
