"""Host-side section timings (TNT_PROFILE=1) of one end-to-end step for the three ways a database
reaches the engine: fragments as codes, packed snapshot, FASTA text.  `python tools/e2e_profile.py [Mbp]`"""
import os
import sys
import time

os.environ["TNT_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
records, fragments, assays, db_bases = bench.build_workload(0, mbp, 100, pinned=True)
opts = search_options(min_primer_tm=45.0, min_probe_tm=50.0, max_len=2000)
fl = FragmentList(fragments)
with Engine() as e:
    e.set_assays([Assay(i, a[0], a[1], a[2]) for i, a in enumerate(assays)])
    e.add_targets(fl)
    e.search_raw(opts)
    snap = e.export_packed()
    pinned = {}
    for k, v in snap.items():
        if k == "info" or v.size == 0:
            pinned[k] = v
            continue
        t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(v.dtype)
        t[:] = v
        pinned[k] = t
    for name in ("codes", "snapshot", "snapshot_sync", "codes", "snapshot", "snapshot_sync"):
        torch.cuda.synchronize()
        print("==== %s" % name, file=sys.stderr)
        t0 = time.perf_counter()
        e.clear_targets()
        t1 = time.perf_counter()
        if name == "codes":
            e.add_targets(fl)
        else:
            e.import_packed(pinned)
            if name == "snapshot_sync":
                ts = time.perf_counter()
                torch.cuda.synchronize()
                print("import transfer alone: %.2f ms" % ((time.perf_counter() - ts) * 1e3), file=sys.stderr)
        t2 = time.perf_counter()
        n = e.search_raw(opts)
        t3 = time.perf_counter()
        print("clear %.2f ms, register %.2f ms, search %.2f ms, total %.2f ms, hits %d"
              % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, n), file=sys.stderr)
