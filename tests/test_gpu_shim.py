"""The reference program with its search path on the engine (tests/shim/_build/tntblast_gpu: the
unmodified reference objects, its amplicon()/padlock()/hybrid() resolved to the C-ABI shim) against
the unmodified reference binary (oracle/_ref/tntblast) on the same command line: the text output
has to be byte-identical.  One case per BASELINE.json configuration, config 1 in full, the others on
leading slices of the same synthetic databases with the real flags.

Both binaries are built where /root/reference exists (oracle/Makefile, tests/shim/Makefile) and
travel to the GPU box; nothing here reads /root/reference at run time.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import gen
import harness as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tntblast")
GPU_BIN = os.path.join(ROOT, "tests", "shim", "_build", "tntblast_gpu")


def scale(default_mbp: float) -> int:
    """Slice size in bases; TNT_SHIM_SCALE shrinks or grows every case (default 1.0)."""
    return int(default_mbp * 1e6 * float(os.environ.get("TNT_SHIM_SCALE", "1.0")))


def run_pair(tmp_path, records, assays, flags, limit_bp=None, threads=None, env_gpu=None):
    if not (os.path.exists(REF_BIN) and os.path.exists(GPU_BIN)):
        pytest.skip("oracle/_ref/tntblast or tests/shim/_build/tntblast_gpu not built (needs /root/reference)")
    fa = str(tmp_path / "db.fa")
    q = str(tmp_path / "assays.txt")
    used = gen.write_fasta(fa, records, limit_bp=limit_bp)
    gen.write_assays(q, assays)
    threads = threads or min(len(os.sched_getaffinity(0)), 32)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    out_ref, out_gpu = str(tmp_path / "ref.out"), str(tmp_path / "gpu.out")
    r = subprocess.run([REF_BIN, "-i", q, "-d", fa, "-o", out_ref] + flags, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env2 = dict(env)
    env2.update(env_gpu or {})
    g = subprocess.run([GPU_BIN, "-i", q, "-d", fa, "-o", out_gpu] + flags, env=env2, capture_output=True, text=True)
    assert g.returncode == 0, g.stderr[-2000:] + g.stdout[-2000:]
    m = re.search(r"\[tntb200\] prefetched=(\d+) fragments=(\d+) bases=(\d+) batches=(\d+) alignments=(\d+) hits=(\d+) "
                  r"device_ms=([0-9.]+) calls_from_table=(\d+) calls_direct=(\d+)", g.stderr)
    assert m, g.stderr[-2000:]
    info = dict(zip(("prefetched", "fragments", "bases", "batches", "alignments", "hits", "device_ms", "table", "direct"),
                    [float(x) for x in m.groups()]))
    a, b = open(out_ref, "rb").read(), open(out_gpu, "rb").read()
    keep = os.environ.get("TNT_SHIM_KEEP")   # debugging aid: keep differing outputs (e.g. under gpurun_out/)
    if keep and a != b:
        os.makedirs(keep, exist_ok=True)
        import shutil
        shutil.copy(out_ref, os.path.join(keep, tmp_path.name + ".ref.out"))
        shutil.copy(out_gpu, os.path.join(keep, tmp_path.name + ".gpu.out"))
    # the stdout summary (counts, Tm / dG / length ranges of all matches) has to agree as well, apart
    # from the elapsed-time line and the progress meter
    def summary(s):
        s = s[s.index("Found"):] if "Found" in s else s
        keep = [ln for ln in s.splitlines() if not ln.startswith("Search completed") and not ln.startswith("Searching database")
                and not ln.startswith("\tOutput = ")]
        return "\n".join(keep)
    assert summary(r.stdout) == summary(g.stdout)
    return a, b, info, used


def first_difference(a: bytes, b: bytes) -> str:
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return "line %d:\n  ref: %r\n  gpu: %r" % (i + 1, x[:200], y[:200])
    return "lengths differ: %d vs %d lines" % (len(la), len(lb))


def check_identical(a, b, info, min_hits):
    assert a == b, first_difference(a, b)
    assert a.count(b"name = ") >= min_hits, "too few hits for a meaningful comparison: %d" % a.count(b"name = ")
    # every driver call was answered from the batch pass: no silent detour
    assert info["prefetched"] == 1 and info["direct"] == 0 and info["table"] > 0 and info["alignments"] > 0


def test_config1_single_pcr_pair_5mbp_full(tmp_path):
    """BASELINE configs[0]: 1 PCR primer pair vs a 5 Mbp record, -e 40, in full (10 fragments + overlaps)."""
    rng = np.random.default_rng(1234)
    records = [gen.random_codes(5_000_000, rng)]
    assays = gen.make_assays(np.random.default_rng(11), records, 1, "pcr", lens=(20, 20, 25), amp=(450, 550), variants=6)
    a, b, info, used = run_pair(tmp_path, records, assays, ["-e", "40"])
    assert used == 5_000_000 and info["fragments"] >= 10   # cut by the byte length of the FASTA record
    check_identical(a, b, info, 2)


def test_config1_direct_calls(tmp_path):
    """The same run without the batch pass: every amplicon() call of the driver uploads its fragment
    and searches on the spot (what the shim does under a stock main)."""
    rng = np.random.default_rng(1234)
    records = [gen.random_codes(2_000_000, rng)]
    assays = gen.make_assays(np.random.default_rng(11), records, 2, "taqman", variants=4)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-e", "40", "-E", "45"], env_gpu={"TNTB200_NO_PREFETCH": "1"})
    assert a == b, first_difference(a, b)
    assert info["prefetched"] == 0 and info["direct"] > 0 and a.count(b"name = ") >= 2


def test_dinkelbach_flag(tmp_path):
    """`--dinkelbach T` (options.cpp:222-224 -> melt.dinkelbach, tntblast_local.cpp:367): the shim hands the
    flag to tnt_engine_create and the iterative Tm of every window comes from the device."""
    rng = np.random.default_rng(77)
    records = [gen.random_codes(1_000_000, rng)]
    assays = gen.make_assays(np.random.default_rng(5), records, 4, "taqman", variants=4)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-e", "40", "-E", "45", "--dinkelbach", "T"])
    check_identical(a, b, info, 4)
    # and the flag changes the output (it is not silently ignored by either program)
    c, d, _, _ = run_pair(tmp_path, records, assays, ["-e", "40", "-E", "45"])
    assert c == d and c != a


def test_config2_taqman_slice(tmp_path):
    """BASELINE configs[1]: 100 TaqMan triplets, -e 45 -E 50, leading 50 Mbp of the 1 Gbp database."""
    n = scale(50)
    rng = np.random.default_rng(2)
    records = [gen.random_codes(min(5_000_000, n - i), rng) for i in range(0, n, 5_000_000)]
    assays = gen.make_assays(np.random.default_rng(99), records, 100, "taqman", lens=(20, 21, 25), amp=(80, 400), variants=2)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-e", "45", "-E", "50"])
    check_identical(a, b, info, 100)


def test_config3_degenerate_probes_slice(tmp_path):
    """BASELINE configs[2]: hybridisation probes carrying an inosine and a two-fold code (expanded by
    the reference's own expand_degenerate_signatures inside both programs), database with 0.1 % IUPAC
    codes and N runs, -A PROBE -E 50, both strands."""
    n = scale(50)
    rng = np.random.default_rng(3)
    records = [gen.random_codes(min(5_000_000, n - i), rng) for i in range(0, n, 5_000_000)]
    base = gen.make_assays(np.random.default_rng(33), records, 40, "probe", lens=(20, 21, 30), amp=(80, 400), variants=3)
    assays = []
    for (_, _, P) in base:
        p = list(P)
        p[10] = "I"
        p[20] = {"A": "R", "G": "R", "C": "Y", "T": "Y"}[p[20]]
        assays.append((None, None, "".join(p)))
    for rec in records:
        gen.sprinkle_degenerate(rec, rng, frac=1e-3, n_runs_per_50kb=1.0)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-A", "PROBE", "-E", "50"])
    check_identical(a, b, info, 40)


def test_config4_padlock_slice(tmp_path):
    """BASELINE configs[3]: padlock probe pairs 20+20, -A PADLOCK -e 40, records of unequal size."""
    n = scale(50)
    rng = np.random.default_rng(4)
    sizes = [int(n * f) for f in (0.34, 0.27, 0.2, 0.12, 0.07)]
    records = [gen.random_codes(s, rng) for s in sizes]
    assays = gen.make_assays(np.random.default_rng(44), records, 40, "padlock", lens=(20, 20, 25), variants=3)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-A", "PADLOCK", "-e", "40"])
    check_identical(a, b, info, 40)


def test_config5_thousand_pcr_slice(tmp_path):
    """BASELINE configs[4]: 1000 PCR assays, -e 45, leading 50 Mbp of the 100 Gbp database."""
    n = scale(50)
    rng = np.random.default_rng(5)
    records = [gen.random_codes(min(5_000_000, n - i), rng) for i in range(0, n, 5_000_000)]
    assays = gen.make_assays(np.random.default_rng(55), records, 1000, "pcr", lens=(20, 21, 25), amp=(80, 400), variants=1)
    a, b, info, _ = run_pair(tmp_path, records, assays, ["-e", "45"])
    check_identical(a, b, info, 1000)
