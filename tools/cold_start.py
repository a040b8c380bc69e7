import os, sys, time
os.environ["TNT_PROFILE"] = "1"
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import bench
from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options
records, fragments, assays, db = bench.build_workload(0, 1000, 100, pinned=True)
opts = search_options(min_primer_tm=45.0, min_probe_tm=50.0, max_len=2000)
fl = FragmentList(fragments)
torch.cuda.synchronize()
t0 = time.perf_counter()
e = Engine()
t1 = time.perf_counter()
e.set_assays([Assay(i, a[0], a[1], a[2]) for i, a in enumerate(assays)])
t2 = time.perf_counter()
e.add_targets(fl)
t3 = time.perf_counter()
n = e.search_raw(opts)
t4 = time.perf_counter()
print("create %.1f ms, set_assays %.1f, add_targets %.1f, first search %.1f, hits %d" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, n), file=sys.stderr)
e.close()
