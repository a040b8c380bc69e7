"""FASTA ingest alone: `python tools/ingest_timing.py [Mbp]` builds a synthetic 80-column FASTA text in
page-locked memory, runs tnt_engine_add_fasta a few times and prints the ingest counters
(used under ncu for the launch list of the parser kernels)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import gen  # noqa: E402
from thermonucleotideblast_b200 import Engine  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(1)
records = [gen.random_codes(5_000_000, rng) for _ in range(mbp // 5)]
text, nbytes = bench.fasta_text_pinned(records)
with Engine() as e:
    for it in range(3):
        e.clear_targets()
        t0 = time.perf_counter()
        nf = e.add_fasta_raw(int(text.ctypes.data), nbytes, 500000, 2002)
        dt = time.perf_counter() - t0
        st = e.ingest_stats()
        if it == 0:
            # spot check at the far end of the database (offsets beyond 4 GB for large runs)
            recs, frags = None, None
            last = e.target_codes(nf - 1, 0, 1000)
            # the last registered fragment is the tail piece of the last record
            from thermonucleotideblast_b200.sharding import fragment_record
            a, b = fragment_record(len(records[-1]) + len(b">rec%d synthetic\n" % (len(records) - 1)) + (len(records[-1]) + 79) // 80, 500000)[-1]
            want = records[-1][a:a + 1000]
            assert last.tolist() == want.tolist(), "tail fragment differs"
            print("tail fragment verified (start %d of record %d)" % (a, len(records) - 1), file=sys.stderr)
        print(json.dumps({"fragments": nf, "call_ms": dt * 1e3, "parse_ms": st.parse_ms, "slabs": st.slabs,
                          "text_bytes": st.text_bytes, "bases": st.bases}), file=sys.stderr)
