"""One warm full-size search of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` captures (tools/capture_profiles.sh): the capture then holds the launches
of exactly one `tnt_engine_search`.  `python tools/one_search.py [Mbp] [assays] [kind]`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nassay = int(sys.argv[2]) if len(sys.argv) > 2 else 100
kind = sys.argv[3] if len(sys.argv) > 3 else "taqman"
records, fragments, assays, db_bases = bench.build_workload(0, mbp, nassay, pinned=True, kind=kind)
opts = search_options(min_primer_tm=bench.MIN_PRIMER_TM, min_probe_tm=bench.MIN_PROBE_TM, max_len=bench.MAX_LEN)
if kind == "probe":
    opts.assay_format = 1
elif kind == "padlock":
    opts.assay_format = 2
    opts.min_probe_tm = 40.0
with Engine() as e:
    e.set_assays([Assay(i, a[0], a[1], a[2], probe_degen=(a[3] if len(a) > 3 else 1)) for i, a in enumerate(assays)])
    e.add_targets(FragmentList(fragments))
    e.search_raw(opts)
    e.search_raw(opts)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e.search_raw(opts)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    st = e.stats()
    print("alignments", st.alignments, "dp_cells", st.dp_cells)
