"""Randomised differential of the device FASTA reader (tnt_engine_add_fasta) against the oracle's restatement of
the reference reader: random texts (line widths, CR LF, blank lines, lower case, IUPAC, stray blanks, '*', '-',
'>' inside lines, tabs, no final newline, records of length 0) x random fragment settings.  Run by hand on a GPU
box: python tests/fuzz_fasta.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

import gen  # noqa: E402
import harness as H  # noqa: E402
from test_gpu_fasta import check_against  # noqa: E402
from thermonucleotideblast_b200 import Engine  # noqa: E402
from thermonucleotideblast_b200.engine import EngineError  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
oracle = H.oracle()
t_end = time.time() + budget
it = nbad = nrefused = ndoc = 0
with Engine() as eng:
    while time.time() < t_end:
        it += 1
        rng = np.random.default_rng(seed0 * 7919 + it)
        text = gen.rand_fasta(rng, n_records=int(rng.integers(1, 12)), max_len=int(rng.choice([40, 700, 6000, 60000])),
                              width=int(rng.choice([0, 1, 7, 60, 61, 80, 100])), crlf=bool(rng.integers(0, 3) == 0),
                              iupac=float(rng.choice([0.0, 0.01, 0.2])), lower=float(rng.choice([0.0, 0.1, 1.0])))
        b = bytearray(text)
        # sprinkle characters the reader treats specially
        for _ in range(int(rng.integers(0, 12))):
            if not b:
                break
            p = int(rng.integers(0, len(b)))
            ch = bytes([int(rng.choice(list(b" \t*-\n\rUu>")))])
            if rng.integers(0, 2):
                b[p:p] = ch
            elif b[p] != ord(">"):
                b[p:p + 1] = ch
        if rng.integers(0, 4) == 0 and b.endswith(b"\n"):
            del b[-1]
        text = bytes(b)
        threshold = int(rng.choice([0, 64, 1000, 50000]))
        overlap = int(rng.choice([0, 5, 50, 2002]))
        try:
            want = oracle.fasta_records(text, threshold=threshold, overlap=overlap)
        except Exception as ex:           # texts the reference reader refuses (e.g. an empty defline)
            nrefused += 1
            try:
                eng.clear_targets()
                eng.add_fasta(text, fragment_threshold=threshold, overlap=overlap)
                print("ACCEPTED-BUT-ORACLE-REFUSED it=%d: %s" % (it, str(ex)[:100]), flush=True)
                nbad += 1
            except EngineError:
                pass
            continue
        try:
            check_against(eng, text, want, threshold, overlap)
        except EngineError as ex:
            if "empty defline" in str(ex):
                # documented refusal: a '>' directly followed by the end of the line;
                # the reference takes the next line for the defline or throws
                ndoc += 1
                continue
            nbad += 1
            print("ERROR it=%d threshold=%d overlap=%d len=%d: %s" % (it, threshold, overlap, len(text), str(ex)[:300]), flush=True)
        except AssertionError as ex:
            nbad += 1
            print("MISMATCH it=%d threshold=%d overlap=%d len=%d: %s" % (it, threshold, overlap, len(text), str(ex)[:300]), flush=True)
            open(os.path.join(ROOT, "gpurun_out", "fuzz_fasta_case_%d.fa" % it), "wb").write(text)
print("fuzz_fasta: %d texts, %d refused by both, %d refused by the engine as documented (empty defline), %d mismatches" % (it, nrefused, ndoc, nbad))
sys.exit(1 if nbad else 0)
