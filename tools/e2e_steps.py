"""Per-step wall times of the end-to-end leg of bench.py (host codes -> hits incl. text), 12 steps, to see
whether the mean hides outliers.  `python tools/e2e_steps.py [Mbp]`"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
records, fragments, assays, db_bases = bench.build_workload(0, mbp, 100, pinned=True)
opts = search_options(min_primer_tm=bench.MIN_PRIMER_TM, min_probe_tm=bench.MIN_PROBE_TM, max_len=bench.MAX_LEN)
fl = FragmentList(fragments)
with Engine() as e:
    e.set_assays([Assay(i, a[0], a[1], a[2]) for i, a in enumerate(assays)])
    e.add_targets(fl)
    for _ in range(3):
        e.search_raw(opts)
    out = []
    for step in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e.clear_targets()
        e.add_targets(fl)
        t1 = time.perf_counter()
        e.search_raw(opts)
        t2 = time.perf_counter()
        e.hit_records()
        e.hit_sequences_bytes()
        t3 = time.perf_counter()
        st = e.stats()
        out.append((t1 - t0, t2 - t1, t3 - t2, st.align_ms, st.scan_ms))
    for i, (a, b, c, am, sm) in enumerate(out):
        print("step %2d: upload call %6.2f ms  search %7.2f ms  results %5.2f ms  total %7.2f ms   (kernels: align %.1f scan %.1f)"
              % (i, a * 1e3, b * 1e3, c * 1e3, (a + b + c) * 1e3, am, sm))
