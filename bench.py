#!/usr/bin/env python
"""Headline benchmark of the tntblast search hot path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbp M] [--assays A]

One "step" is one pass of the hot path (seed scan -> NucCruc Tm/dG alignment -> amplicon
assembly) over the whole workload.  Default workload = BASELINE.json configs[1]:
100 TaqMan primer+probe triplets vs a synthetic 1 Gbp multi-record database (200 records x 5 Mbp,
cut into <= 500 kbp fragments with the reference's overlap), flags -e 45 -E 50.

Metric: DB Gbp*assay/s (database bases x input assays per second), whole job over all ranks.
  value : inputs already resident in HBM when the timed region starts (search only)
  e2e   : the same through the C ABI with host buffers: clear + upload of every fragment
          (pinned staging, H2D, 2-bit pack) + search + hit read-back inside the timed region
  ingest_fasta         : the same database as 80-column FASTA text in page-locked memory -> tnt_engine_add_fasta
                         (parsed, cut and packed on the device) + search + hit read-back
  e2e_packed_snapshot  : tnt_engine_import_packed of the exported 2-bit database + search + hit read-back
Extra keys: alignments_per_s (NucCruc heterodimer evaluations), roofline (NucCruc DP kernel,
int32 issue roofline of SURVEY 8d), roofline_seed_scan (HBM), cpu_baseline (the unmodified
reference OpenMP build on the host cores on a bounded slice of the same workload).

Under torchrun (N > 1) every rank owns one GPU and a contiguous shard of the database of the same
size (weak scaling; the path has no exchange step, so NCCL only carries the barrier and the
max-over-ranks of the timings).  `--impl reference` times the reference's own CPU implementation
(oracle/_ref/tntblast, built from /root/reference by oracle/Makefile) on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gen  # noqa: E402

METRIC = "DB Gbp*assay/s (tntblast search hot path: seed scan + NucCruc Tm/dG alignment + amplicon assembly)"
UNIT = "Gbp*assay/s"
RECORD_BP = 5_000_000
FRAGMENT_BP = 500_000          # DEFAULT_FRAGMENT_TARGET_LENGTH, tntblast.h:81
MAX_LEN = 2000                 # DEFAULT_MAX_LEN
OVERLAP = MAX_LEN + 2          # opt.max_product_length() + 2, tntblast_local.cpp:174
MIN_PRIMER_TM, MIN_PROBE_TM = 45.0, 50.0
ALU_OPS_PER_CELL = 27          # 32-bit ALU ops per DP cell of align_dimer
SM_COUNT, LANES_PER_SM = 148, 128


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def build_workload(rank: int, mbp: int, n_assays: int, pinned: bool = False, kind: str = "taqman"):
    """Synthetic shard of `mbp` Mbp (records of 5 Mbp) + TaqMan assays planted into it.

    With pinned=True the bases live in one page-locked host allocation (the fragments are views
    into it), so the engine's uploads are true pinned-memory H2D copies."""
    from thermonucleotideblast_b200.sharding import fragment_record
    rng = np.random.default_rng(2 + 1000 * rank)
    total = mbp * 1_000_000
    if pinned:
        import torch
        store = torch.empty(total, dtype=torch.uint8, pin_memory=True).numpy()
    else:
        store = np.empty(total, dtype=np.uint8)
    records = []
    pos = 0
    while pos < total:
        n = min(RECORD_BP, total - pos)
        store[pos:pos + n] = gen.random_codes(n, rng)
        records.append(store[pos:pos + n])
        pos += n
    arng = np.random.default_rng(99)  # same assays on every rank
    if kind == "probe":
        # BASELINE configs[2] flavour: 30-mer probes carrying one inosine and one two-fold code
        # (expanded by the caller into two concrete oligos of degeneracy 2, like the reference's
        # expand_degenerate_signatures), database with 0.1 % IUPAC codes and N runs
        base = gen.make_assays(arng, records, n_assays, "probe", lens=(20, 21, 30), amp=(80, 400), variants=2)
        assays = []
        for (_, _, P) in base:
            p = list(P)
            i_pos, d_pos = 10, 20
            p[i_pos] = "I"
            two = {"A": "AG", "G": "AG", "C": "CT", "T": "CT"}[p[d_pos]]
            for b in two:
                q = list(p)
                q[d_pos] = b
                assays.append((None, None, "".join(q), 2))
        for rec in records:
            gen.sprinkle_degenerate(rec, rng, frac=1e-3, n_runs_per_50kb=1.0)
    else:
        assays = gen.make_assays(arng, records, n_assays, kind, lens=(20, 21, 25), amp=(80, 400), variants=2)
    fragments = []
    for rec in records:
        for (a, b) in fragment_record(len(rec), FRAGMENT_BP):
            fragments.append(rec[a:min(len(rec), b + 1 + OVERLAP)])
    return records, fragments, assays, total


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank (and allocate its page-locked host buffers) on the CPUs of the NUMA node its GPU hangs
    off: with one rank per GPU all ranks pull their gigabyte through the host memory system at the same
    moment, and remote-node pinned memory made the end-to-end leg lose 6 % at 8 GPUs.
    Best effort: any failure leaves the affinity alone.  Returns the node or None."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def final_records_equal(a: bytes, b: bytes) -> bool:
    """Two finished hit lists (bytes of tnt_final_hit records, engine.finalize_hits(raw=True)) hold the same hits in
    the same order: assay, record, coordinates, strand, Tm and mismatch counts of both primers (block / index /
    text offsets differ between a sharded and a single-engine run by construction)."""
    from thermonucleotideblast_b200.engine import FinalHit
    dt = np.dtype(FinalHit)
    x, y = np.frombuffer(a, dtype=dt), np.frombuffer(b, dtype=dt)
    if x.size != y.size:
        return False
    hx, hy = x["hit"], y["hit"]
    for f in ("assay_index", "target_id", "amp_first", "amp_last", "primer_strand"):
        if not np.array_equal(hx[f], hy[f]):
            return False
    for side in ("forward", "reverse"):
        for f in ("tm", "num_mm"):
            if not np.array_equal(hx[side][f], hy[side][f]):
                return False
    return True


def traffic_from_profile():
    """DRAM bytes of one full-size search, summed per kernel family from the newest committed ncu
    capture `profiles/dram_r*.csv` (tools/capture_profiles.sh: dram__bytes_read.sum + dram__bytes_write.sum of
    every launch between cudaProfilerStart/Stop around exactly one warm search of the record workload).
    Returns (nuccruc_bytes, scan_bytes, file name) or (None, None, None) when no capture is committed."""
    import csv
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "dram_r*.csv")),
                   key=lambda f: [int(x) for x in re.findall(r"\d+", os.path.basename(f))])
    if not files:
        return None, None, None
    path = files[-1]
    nuccruc = scan = 0.0
    with open(path) as fh:
        rows = csv.DictReader(line for line in fh if line.startswith('"'))
        for row in rows:
            if not row["Metric Name"].startswith("dram__bytes"):
                continue
            v = float(row["Metric Value"].replace(",", ""))
            unit = row["Metric Unit"].lower()
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            name = row["Kernel Name"]
            if "k_align" in name:
                nuccruc += v
            elif "k_seed_scan" in name or "k_region_scan" in name:
                scan += v
    return nuccruc or None, scan or None, os.path.basename(path)


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md) through NVML.

    One step in the middle of the timed region is sampled, exactly once, from a helper thread
    that fires a fixed delay into the step, i.e. while the alignment kernels are running; the step
    waits for its sample before it ends, so whatever the query costs lands inside the timing.  A
    continuously polling nvidia-smi / NVML thread was measured to slow the CUDA API calls of a step
    by >30 %, and on some boxes a single NVML query takes 20-50 ms and stalls concurrent CUDA
    calls for as long -- hence one query in the timed region, and further samples from an extra,
    untimed repetition of the same step right after it (`extra_step`)."""

    def __init__(self, gpu_index: int, steps: int):
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self.cost_s = 0.0
        self._h = None
        self.timed_samples = 0
        self.when = {steps // 2}
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                parts = [x for x in vis.split(",") if x.strip() != ""]
                if gpu_index < len(parts) and parts[gpu_index].strip().isdigit():
                    idx = int(parts[gpu_index])
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            # the first query of a process can take > 100 ms (lazy initialisation): not in the timed region
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            try:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        except Exception:
            self._h = None

    def start_step(self, step: int, delay_s: float):
        self._thread = None
        if self._h is None or step not in self.when:
            return

        def run():
            time.sleep(max(0.0, delay_s))
            self.sample(step)

        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()

    def end_step(self):
        if getattr(self, "_thread", None) is not None:
            self._thread.join()
            self._thread = None

    def extra_step(self, run_step, delay_s: float, n: int = 2):
        """Untimed repetition(s) of the step, sampled like the timed one."""
        if self._h is None:
            return
        for _ in range(n):
            self.when = {-1}
            self.start_step(-1, delay_s)
            run_step()
            self.end_step()

    def sample(self, step: int):
        if self._h is None or step not in self.when:
            return
        if step >= 0:
            self.timed_samples += 1
        t0 = time.perf_counter()
        nv = self._nv
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                              ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                              ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass
        self.cost_s += time.perf_counter() - t0

    def result(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
               "samples": len(self.samples), "samples_in_timed_region": self.timed_samples,
               "source": "nvml, one query per sampled step, issued while the step's kernels run; "
                         "one step of the timed region + untimed repetitions of the same step right after it",
               "sampling_ms_total": self.cost_s * 1e3}
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
        return out


def write_reference_inputs(tmp: str, records, assays, sample_bp: int):
    """FASTA (80 columns) of the leading `sample_bp` bases + tab-separated assay file."""
    fa = os.path.join(tmp, "db.fa")
    lut = np.frombuffer(b"ACGTIMRSVWYHKDBN-N", dtype=np.uint8)
    left = sample_bp
    with open(fa, "wb") as f:
        for i, rec in enumerate(records):
            if left <= 0:
                break
            n = min(len(rec), left)
            f.write(b">rec%d synthetic\n" % i)
            txt = lut[rec[:n]]
            full = (n // 80) * 80
            if full:
                body = np.empty((full // 80, 81), dtype=np.uint8)
                body[:, :80] = txt[:full].reshape(-1, 80)
                body[:, 80] = ord("\n")
                f.write(body.tobytes())
            if n > full:
                f.write(txt[full:].tobytes() + b"\n")
            left -= n
    q = os.path.join(tmp, "assays.txt")
    gen.write_assays(q, assays)
    return fa, q, sample_bp - max(left, 0)


REF_EXE = os.path.join(ROOT, "oracle", "_ref", "tntblast")
REF_COUNTED_EXE = os.path.join(ROOT, "oracle", "_ref", "tntblast_counted")   # + a call counter (oracle/count_wrap.cpp)
GPU_EXE = os.path.join(ROOT, "tests", "shim", "_build", "tntblast_gpu")     # the reference objects on the engine


def reference_assays(assays, kind: str):
    """The assay list as the reference's input file holds it: the probe flavour carries a two-fold code
    that the reference expands itself (expand_degenerate_signatures), the engine gets the two expansions."""
    if kind != "probe":
        return assays
    return [(None, None, a[2][:20] + {"A": "R", "G": "R", "C": "Y", "T": "Y"}[a[2][20]] + a[2][21:]) for a in assays[::2]]


def reference_flags(kind: str):
    if kind == "probe":
        return ["-A", "PROBE", "-E", str(MIN_PROBE_TM)]
    if kind == "padlock":
        return ["-A", "PADLOCK", "-e", "40"]
    if kind == "pcr":
        return ["-e", str(MIN_PRIMER_TM)]
    return ["-e", str(MIN_PRIMER_TM), "-E", str(MIN_PROBE_TM)]


def run_reference_once(fa, q, cores: int, tmp: str, kind: str = "taqman", exe: str = None, out: str = "out.txt") -> float:
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    t0 = time.perf_counter()
    r = subprocess.run([exe or REF_EXE, "-i", q, "-d", fa, "-o", os.path.join(tmp, out)] + reference_flags(kind),
                       check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    run_reference_once.last_stderr = r.stderr
    return time.perf_counter() - t0


def hit_blocks(path: str):
    """The hit records of a tntblast text output (one block per reported match: 'name = ...' up to the blank line)."""
    blocks, cur = [], []
    with open(path, "rb") as f:
        for ln in f:
            if ln.startswith(b"name = ") and cur:
                blocks.append(b"".join(cur))
                cur = []
            if ln.startswith(b"####"):
                continue
            cur.append(ln)
    if cur:
        blocks.append(b"".join(cur))
    return [b for b in blocks if b.startswith(b"name = ")]


def parity_at_scale(fa, q, cores: int, tmp: str, kind: str):
    """The same slice through tests/shim/_build/tntblast_gpu (the unmodified reference objects with
    amplicon()/padlock()/hybrid() resolved to the engine's C ABI): its text output against the text
    output the reference binary wrote in the timed run."""
    import re
    from collections import Counter
    if not os.path.exists(GPU_EXE):
        return {"checked": False, "why": "tests/shim/_build/tntblast_gpu not built (needs /root/reference at build time)"}
    try:
        run_reference_once(fa, q, cores, tmp, kind, exe=GPU_EXE, out="out_gpu.txt")
    except subprocess.CalledProcessError as ex:
        return {"checked": False, "why": "tntblast_gpu failed: %s" % (ex.stderr or "")[-300:]}
    shim_line = [ln for ln in run_reference_once.last_stderr.splitlines() if ln.startswith("[tntb200]")]
    a, b = os.path.join(tmp, "out.txt"), os.path.join(tmp, "out_gpu.txt")
    identical = open(a, "rb").read() == open(b, "rb").read()
    ra, rb = hit_blocks(a), hit_blocks(b)
    ca, cb = Counter(ra), Counter(rb)
    only_ref, only_gpu = sum((ca - cb).values()), sum((cb - ca).values())
    # hits with a Tm within 0.01 C of the bound it was filtered with (north_star: listed separately)
    bounds = {b"forward primer tm": MIN_PRIMER_TM, b"reverse primer tm": MIN_PRIMER_TM, b"probe tm": MIN_PROBE_TM}
    if kind == "padlock":
        bounds = {b"forward primer tm": 40.0, b"reverse primer tm": 40.0}
    near = 0
    pat = re.compile(rb"^(forward primer tm|reverse primer tm|probe tm) = ([-0-9.e+]+)$", re.M)
    for blk in ra:
        if any(abs(float(v) - bounds[k]) <= 0.0105 for k, v in pat.findall(blk) if k in bounds):
            near += 1
    return {"checked": True, "byte_identical_output": bool(identical), "hits": len(ra), "hits_engine": len(rb),
            "mismatches": int(only_ref + only_gpu), "only_in_reference": int(only_ref), "only_in_engine": int(only_gpu),
            "near_threshold": int(near),
            "how": "text output of tests/shim/_build/tntblast_gpu (reference objects + C-ABI shim) vs the reference binary's, "
                   "same command line, same slice; hit blocks compared as multisets, files byte for byte",
            "shim": shim_line[-1] if shim_line else None}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_alignment_count(fa, q, cores, tmp, kind):
    """NucCruc heterodimer evaluations of the reference on this input (untimed run of the counted binary)."""
    import re
    if not os.path.exists(REF_COUNTED_EXE):
        return None
    try:
        run_reference_once(fa, q, cores, tmp, kind, exe=REF_COUNTED_EXE, out="out_counted.txt")
    except subprocess.CalledProcessError:
        return None
    m = re.search(r"approximate_tm_heterodimer calls: (\d+)", run_reference_once.last_stderr)
    return int(m.group(1)) if m else None


def cpu_baseline(records, assays, kind: str = "taqman", seconds_target: float = 15.0):
    """(cpu_baseline, parity_at_scale): the reference binary timed on a leading slice of the same
    database, its alignment count, and its text output compared with the engine-backed program's."""
    if not os.path.exists(REF_EXE):
        return ({"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                 "sample": "oracle/_ref/tntblast missing (build it with make -C oracle ref where /root/reference exists)"}, None)
    cores = host_cores()
    # measured in the survey: ~0.016 Gbp*assay/s per core on 100 TaqMan assays
    sample_bp = int(min(sum(len(r) for r in records), max(2_000_000, seconds_target * 0.016e9 * cores / max(len(assays), 1))))
    tmp = tempfile.mkdtemp(prefix="tntref_")
    try:
        fa, q, used = write_reference_inputs(tmp, records, assays, sample_bp)
        dt = run_reference_once(fa, q, cores, tmp, kind)
        aligns = reference_alignment_count(fa, q, cores, tmp, kind)
        parity = parity_at_scale(fa, q, cores, tmp, kind)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    base = {"value": used * len(assays) / 1e9 / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "first %.1f Mbp of the same database x %d assays, OMP_NUM_THREADS=%d, %.1f s wall (file read + search + output)"
                      % (used / 1e6, len(assays), cores, dt),
            "alignments": aligns, "alignments_per_s": (aligns / dt if aligns else None),
            "alignments_how": "calls of NucCruc::approximate_tm_heterodimer in an untimed run of the same command with "
                              "oracle/_ref/tntblast_counted (link-time wrapper, oracle/count_wrap.cpp)"}
    if parity is not None:
        parity["slice"] = "first %.1f Mbp x %d assays" % (used / 1e6, len(assays))
    return base, parity


def fasta_text_pinned(records):
    """The records as one 80-column FASTA text in page-locked host memory (uint8 view, nbytes)."""
    import torch
    lut = np.frombuffer(b"ACGTIMRSVWYHKDBN-N", dtype=np.uint8)
    heads = [b">rec%d synthetic\n" % i for i in range(len(records))]
    total = sum(len(h) + len(r) + (len(r) + 79) // 80 for h, r in zip(heads, records))
    buf = torch.empty(total, dtype=torch.uint8, pin_memory=True).numpy()
    pos = 0
    for h, rec in zip(heads, records):
        buf[pos:pos + len(h)] = np.frombuffer(h, dtype=np.uint8)
        pos += len(h)
        n = len(rec)
        txt = lut[rec]
        full = (n // 80) * 80
        if full:
            body = buf[pos:pos + full // 80 * 81].reshape(-1, 81)
            body[:, :80] = txt[:full].reshape(-1, 80)
            body[:, 80] = 10
            pos += full // 80 * 81
        if n > full:
            buf[pos:pos + n - full] = txt[full:]
            buf[pos + n - full] = 10
            pos += n - full + 1
    assert pos == total
    return buf, total


def fasta_leg(eng, records, opts, units, want_hits, max_over_ranks, barrier, steps):
    text, nbytes = fasta_text_pinned(records)
    addr = int(text.ctypes.data)

    def step():
        eng.clear_targets()
        eng.add_fasta_raw(addr, nbytes, FRAGMENT_BP, OVERLAP)
        n = eng.search_raw(opts)
        eng.hit_records()
        return n

    nh = step()
    barrier()
    t0 = time.perf_counter()
    parse_ms = call_ms = 0.0
    for _ in range(steps):
        nh = step()
        ist = eng.ingest_stats()
        parse_ms += ist.parse_ms
        call_ms += ist.call_ms
    barrier()
    dt = max_over_ranks((time.perf_counter() - t0) / steps)
    ist = eng.ingest_stats()
    algo = float(ist.text_bytes + ist.bases)
    return {"value": units / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": steps,
            "h2d_bytes_per_step": int(nbytes), "text": "80-column FASTA, %d records, page-locked host memory" % len(records),
            "records": int(ist.records), "fragments": int(ist.fragments), "bases": int(ist.bases),
            "hits_equal_fragment_upload": bool(nh == want_hits),
            "add_fasta_call_ms": call_ms / steps, "parser_kernel_ms": parse_ms / steps,
            "parser_launches": int(ist.launches),
            "parse_GBps": algo / (parse_ms / steps / 1e3) / 1e9 if parse_ms else None,
            "parse_algorithmic_bytes": "text read once + 1 B per base written (the parser reads the text twice: summary pass + emission pass)"}


def packed_leg(eng, frag_list, opts, units, want_hits, max_over_ranks, barrier, steps):
    import torch
    eng.clear_targets()
    eng.add_targets(frag_list)
    snap = eng.export_packed()
    pinned = {}
    for k, v in snap.items():
        if k == "info" or v.size == 0:
            pinned[k] = v
            continue
        t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(v.dtype)
        t[:] = v
        pinned[k] = t
    nbytes = int(sum(v.nbytes for k, v in pinned.items() if k != "info"))

    def step():
        eng.clear_targets()
        eng.import_packed(pinned)
        n = eng.search_raw(opts)
        eng.hit_records()
        return n

    nh = step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        nh = step()
    barrier()
    dt = max_over_ranks((time.perf_counter() - t0) / steps)
    return {"value": units / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": steps, "h2d_bytes_per_step": nbytes,
            "hits_equal_fragment_upload": bool(nh == want_hits),
            "what": "tnt_engine_import_packed (2 bit/base + mask + non-ACGT list + fragment table, page-locked host memory) + search + hit read-back"}


# ---------------------------------------------------------------------------------------------
# BASELINE configs[4]: ONE database partitioned over the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------
def run_config5(args, rank, local_rank, world):
    """1000 PCR assays against one database of args.mbp Mbp in all, cut into the reference's fragments
    and dealt out in contiguous shards (sharding.shard_targets); every rank searches its shard on
    its GPU, the tnt_hit records travel to rank 0 over the host (gloo: no NCCL in the data path) and
    are finished there with tnt_finalize_hits (truncation filter at the cuts, record coordinates,
    uniquify_results in the overlaps)."""
    import pickle
    import torch
    import torch.distributed as dist
    from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options
    from thermonucleotideblast_b200.engine import finalize_hits
    from thermonucleotideblast_b200.sharding import fragment_record, shard_targets

    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")

    n_records = max(1, args.mbp * 1_000_000 // RECORD_BP)
    assays = gen.config5_assays(args.assays)
    # global fragment table (record, start, stop, max_stop, len) in database order
    table = []
    for r in range(n_records):
        for (a, b) in fragment_record(RECORD_BP, FRAGMENT_BP):
            table.append((r, a, b, RECORD_BP - 1, min(RECORD_BP, b + 1 + OVERLAP) - a))
    lo, hi = shard_targets([t[4] for t in table], world)[rank]
    mine = table[lo:hi]
    own_records = sorted({t[0] for t in mine})
    store = torch.empty(len(own_records) * RECORD_BP, dtype=torch.uint8, pin_memory=True).numpy()
    rec_view = {}
    for k, r in enumerate(own_records):
        rec_view[r] = store[k * RECORD_BP:(k + 1) * RECORD_BP]
        gen.config5_record(r, assays, rec_view[r])
    fragments = [rec_view[r][a:a + n] for (r, a, b, ms, n) in mine]
    frag_list = FragmentList(fragments)
    shard_bases = int(sum(len(f) for f in fragments))
    db_bases = n_records * RECORD_BP

    opts = search_options(min_primer_tm=MIN_PRIMER_TM, max_len=MAX_LEN)
    eng = Engine(device=local_rank)
    alist = [Assay(i, a[0], a[1], None) for i, a in enumerate(assays)]
    eng.set_assays(alist)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- resident ---------------------------------------------------------------------------
    eng.add_targets(frag_list)
    for _ in range(args.warmup):
        eng.search_raw(opts)
    sampler = ClockSampler(local_rank, args.steps)
    barrier()
    t0 = time.perf_counter()
    dev_ms = scan_ms = align_ms = 0.0
    launches = 0
    for step in range(args.steps):
        sampler.start_step(step, 0.010)
        nhits = eng.search_raw(opts)
        sampler.end_step()
        st = eng.stats()
        dev_ms += st.total_ms
        scan_ms += st.scan_ms
        align_ms += st.align_ms
        launches += st.kernel_launches
    barrier()
    dt = reduce((time.perf_counter() - t0) / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    sampler.extra_step(lambda: eng.search_raw(opts), 0.010, n=1)
    clocks = sampler.result()
    units = db_bases * len(assays) / 1e9

    # ---- end to end: host fragments -> per-rank hits -> gathered, finished hit list on rank 0 ------
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    gather_s = finalize_s = 0.0
    final_n, final_raw = None, b""
    for _ in range(e2e_steps):
        eng.clear_targets()
        eng.add_targets(frag_list)
        eng.search_raw(opts)
        n, raw, text = eng.hit_records()
        eng.hit_sequences_bytes()
        d2h = int(eng.stats().d2h_bytes)
        tg = time.perf_counter()
        block = (raw, n, text, [tuple(t) for t in mine])
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(block, gathered, dst=0, group=host_group)
        else:
            gathered = [block]
        gather_s += time.perf_counter() - tg
        if rank == 0:
            tf = time.perf_counter()
            final_n, final_raw = finalize_hits(gathered, alist, raw=True)
            finalize_s += time.perf_counter() - tf
    barrier()
    e2e_dt = reduce((time.perf_counter() - t0) / e2e_steps, dist.ReduceOp.MAX if world > 1 else None)

    aligns = reduce(float(st.alignments), dist.ReduceOp.SUM if world > 1 else None)
    cells = reduce(float(st.dp_cells), dist.ReduceOp.SUM if world > 1 else None)
    raw_hits = reduce(float(nhits), dist.ReduceOp.SUM if world > 1 else None)
    align_s = reduce(align_ms / args.steps / 1e3, dist.ReduceOp.MAX if world > 1 else None)
    scan_s = reduce(scan_ms / args.steps / 1e3, dist.ReduceOp.MAX if world > 1 else None)
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    alu_peak = world * SM_COUNT * LANES_PER_SM * sm_max * 1e6 / 1e12
    alu_achieved = ALU_OPS_PER_CELL * cells / align_s / 1e12 if align_s > 0 else 0.0

    # optional check (small sizes): the gathered, finished list equals what ONE engine holding the whole database gives
    verify = None
    if args.verify_shards and rank == 0:
        whole = np.empty(n_records * RECORD_BP, dtype=np.uint8)
        for r in range(n_records):
            gen.config5_record(r, assays, whole[r * RECORD_BP:(r + 1) * RECORD_BP])
        e1 = Engine(device=local_rank)
        e1.set_assays(alist)
        e1.add_targets([whole[r * RECORD_BP + a: r * RECORD_BP + a + n] for (r, a, b, ms, n) in table])
        e1.search_raw(opts)
        n1, raw1, text1 = e1.hit_records()
        e1.close()
        single_n, single_raw = finalize_hits([(raw1, n1, text1, [tuple(t) for t in table])], alist, raw=True)
        verify = {"single_engine_hits": single_n, "sharded_hits": final_n,
                  "identical": final_records_equal(single_raw, final_raw)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": units / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 (DP) + f32 (dH/dS/Tm)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: %d PCR assays (-e %g) vs ONE synthetic database of %.3g Gbp (%d records of 5 Mbp, "
                                   "<=%d kbp fragments + %d bp overlap) partitioned over %d GPU(s) in contiguous shards"
                                   % (len(assays), MIN_PRIMER_TM, db_bases / 1e9, n_records, FRAGMENT_BP // 1000, OVERLAP, world),
                       "db_bases_total": db_bases, "fragments_total": len(table), "fragments_rank0": len(mine),
                       "fragment_bases_rank0": shard_bases, "assays": len(assays),
                       "l2": "inputs larger than L2 (packed shard %.0f MB + candidate buffers)" % (shard_bases * 0.375 / 1e6)},
            "alignments_per_s": aligns / dt, "alignments_per_step": aligns, "dp_cells_per_step": cells,
            "hits_per_step_all_ranks_before_gather": int(raw_hits), "hits_after_finalize": final_n,
            "device_ms_per_step": dev_ms / args.steps,
            "kernel_ms_per_step": {"seed_scan": scan_s * 1e3, "nuccruc_align": align_s * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": units / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": shard_bases, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_dt * 1e3, "steps": e2e_steps,
                    "result": "tnt_hit records + alignment strings + amplicon text of every rank, gathered on rank 0 over the host "
                              "(gloo) and finished with tnt_finalize_hits (truncation filter, record coordinates, uniquify_results)",
                    "gather_ms_per_step": gather_s / e2e_steps * 1e3, "finalize_ms_per_step": finalize_s / e2e_steps * 1e3},
            "roofline": {"bound": "alu-int32", "achieved": alu_achieved, "peak": alu_peak, "unit": "TOP/s",
                         "frac": alu_achieved / alu_peak if alu_peak else None, "traffic": None,
                         "kernel": "k_align_fast (NucCruc DP + traceback + evaluation)",
                         "note": "27 int32 ALU ops per DP cell (SURVEY 8d) x cells of all ranks / max-over-ranks kernel time; "
                                 "peak = n_gpus x 148 SM x 128 lanes x %.0f MHz (nominal)" % sm_max},
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "measured by the default (configs[1]) run"},
        }
        if verify is not None:
            line["sharded_equals_single_engine"] = verify
        _emit(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="2: BASELINE configs[1], weak scaling (default, the metric's configuration); "
                         "5: configs[4], one database partitioned over the GPUs (strong scaling; --mbp = total size)")
    ap.add_argument("--verify-shards", action="store_true", help="--config 5: also search everything with one engine and compare")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbp", type=int, default=1000, help="database size per GPU in Mbp")
    ap.add_argument("--assays", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fasta", action="store_true", help="skip the FASTA-text ingest leg")
    ap.add_argument("--kind", default="taqman", choices=["taqman", "pcr", "probe", "padlock"],
                    help="assay type (default: BASELINE configs[1]; probe / padlock: the flavours of configs[2] / configs[3])")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    kind_txt = {"taqman": "TaqMan primer+probe triplets", "pcr": "PCR primer pairs",
                "probe": "hybridisation probes (30-mers with one inosine and one two-fold code, expanded; DB with 0.1 % IUPAC codes and N runs)",
                "padlock": "padlock probe pairs (20+20)"}[args.kind]
    workload = ("%d %s (20/21/25-mers, -e %g -E %g) vs synthetic %.3g Gbp multi-record "
                "database per GPU (%d Mbp records, <=%d kbp fragments + %d bp overlap)"
                % (args.assays, kind_txt, MIN_PRIMER_TM, MIN_PROBE_TM, args.mbp / 1000.0, RECORD_BP // 1_000_000,
                   FRAGMENT_BP // 1000, OVERLAP))

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded per-step sample so that warmup + steps end within a few minutes
        cores = host_cores()
        per_step_s = 6.0
        sample_mbp = max(1, int(per_step_s * 0.016e9 * cores / max(args.assays, 1) / 1e6))
        records, _, assays, _ = build_workload(0, min(args.mbp, max(sample_mbp, 5)), args.assays, kind=args.kind)
        tmp = tempfile.mkdtemp(prefix="tntref_")
        try:
            fa, q, used = write_reference_inputs(tmp, records, reference_assays(assays, args.kind), sample_mbp * 1_000_000)
            for _ in range(args.warmup):
                run_reference_once(fa, q, cores, tmp, args.kind)
            times = [run_reference_once(fa, q, cores, tmp, args.kind) for _ in range(args.steps)]
            ref_aligns = reference_alignment_count(fa, q, cores, tmp, args.kind)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        dt = sum(times) / len(times)
        v = used * len(assays) / 1e9 / dt
        sample = "first %.1f Mbp of the same synthetic database x %d assays per step, OMP_NUM_THREADS=%d" % (used / 1e6, len(assays), cores)
        _emit(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "alignments": ref_aligns, "alignments_per_s": (ref_aligns / dt if ref_aligns else None)},
            "alignments_per_s": (ref_aligns / dt if ref_aligns else None),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    if args.config == 5:
        if args.assays == 100:
            args.assays = 1000
        return run_config5(args, rank, local_rank, world)
    torch.cuda.set_device(local_rank)
    full_affinity = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local_rank)   # pinned host buffers next to the GPU (also at N=1: remote-node memory halves the H2D rate)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from thermonucleotideblast_b200 import Assay, Engine, FragmentList, search_options

    records, fragments, assays, db_bases = build_workload(rank, args.mbp, args.assays, pinned=True, kind=args.kind)
    frag_bases = int(sum(len(f) for f in fragments))
    opts = search_options(min_primer_tm=MIN_PRIMER_TM, min_probe_tm=MIN_PROBE_TM, max_len=MAX_LEN)
    if args.kind == "probe":
        opts.assay_format = 1   # TNT_ASSAY_PROBE
    elif args.kind == "padlock":
        opts.assay_format = 2   # TNT_ASSAY_PADLOCK
        opts.min_probe_tm = 40.0
    eng = Engine(device=local_rank)
    eng.set_assays([Assay(i, a[0], a[1], a[2], probe_degen=(a[3] if len(a) > 3 else 1)) for i, a in enumerate(assays)])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pointer / length arrays of the (pinned) host fragments: what a C++ host passes to
    # tnt_engine_add_targets; the bytes themselves cross PCIe inside the timed e2e region
    frag_list = FragmentList(fragments)

    def upload():
        eng.clear_targets()
        eng.add_targets(frag_list)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident: search only -----------------------------------------------------------
    upload()
    warm_s = 0.1
    for _ in range(args.warmup):
        tw = time.perf_counter()
        eng.search_raw(opts)
        warm_s = time.perf_counter() - tw
    sampler = ClockSampler(local_rank, args.steps)
    barrier()
    t0 = time.perf_counter()
    dev_ms = scan_ms = align_ms = 0.0
    launches = 0
    for step in range(args.steps):
        if not os.environ.get("TNT_NO_SAMPLER"):
            sampler.start_step(step, min(0.010, 0.2 * warm_s))  # the alignment kernels start a few ms into the step
        nhits = eng.search_raw(opts)
        sampler.end_step()
        st = eng.stats()
        dev_ms += st.total_ms
        scan_ms += st.scan_ms
        align_ms += st.align_ms
        launches += st.kernel_launches
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    if not os.environ.get("TNT_NO_SAMPLER"):
        sampler.extra_step(lambda: eng.search_raw(opts), min(0.010, 0.2 * warm_s))
    clocks = sampler.result()
    dt = max_over_ranks(dt)
    st = eng.stats()
    units = sum_over_ranks(db_bases * len(assays) / 1e9)
    aligns = sum_over_ranks(float(st.alignments))
    value = units / dt

    # ---- end to end: host buffers -> hits ------------------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    # warm-up of the whole leg, result reads included (their first call in a process costs ~70 ms once:
    # buffers and the first launch of the text kernels)
    upload()
    eng.search_raw(opts)
    eng.hit_records()
    eng.hit_sequences_bytes()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    upload_s = 0.0
    e2e_step_ms = []
    for _ in range(e2e_steps):
        tu = time.perf_counter()
        upload()
        upload_s += time.perf_counter() - tu
        eng.search_raw(opts)
        n_e2e_hits, hit_bytes, text_bytes = eng.hit_records()   # the step's result, read on the host ...
        n_seq, seq_bytes = eng.hit_sequences_bytes()             # ... with the amplicon / site text of every hit
        # result bytes the engine copied back (site heads of live groups, records of hit sites)
        d2h = int(eng.stats().d2h_bytes)
        e2e_step_ms.append((time.perf_counter() - tu) * 1e3)
    barrier()
    e2e_dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_value = units / e2e_dt

    # ---- FASTA text -> hits (SURVEY 8f ingest row): the same database as 80-column FASTA text in
    #      page-locked host memory, parsed / cut / packed on the device (tnt_engine_add_fasta)
    ingest = None
    if not args.no_fasta:
        try:
            ingest = fasta_leg(eng, records, opts, units, n_e2e_hits, max_over_ranks, barrier, e2e_steps)
        except Exception as ex:
            ingest = {"error": str(ex)}

    # ---- packed snapshot -> hits: the resident 2-bit database exported once, re-imported from
    #      page-locked host memory every step (0.375 B/base over PCIe instead of 1 B/base)
    packed = None
    if not args.no_fasta:
        try:
            packed = packed_leg(eng, frag_list, opts, units, n_e2e_hits, max_over_ranks, barrier, e2e_steps)
        except Exception as ex:
            packed = {"error": str(ex)}

    # ---- seed scan alone with a single assay (the HBM-bound case of SURVEY 8d) -------------------
    scan1 = None
    try:
        eng.set_assays([Assay(0, assays[0][0], assays[0][1], assays[0][2])])
        eng.scan_only(opts)
        best = None
        for _ in range(3):
            ncand, ms1 = eng.scan_only(opts)
            best = ms1 if best is None else min(best, ms1)
        scan1 = {"candidates": int(ncand), "ms": best,
                 "achieved_GBps": (0.375 * frag_bases + 8 * ncand) / (best / 1e3) / 1e9 if best else None}
    except Exception as ex:  # keep the headline line even if this extra fails
        scan1 = {"error": str(ex)}

    # ---- rooflines ----------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_max = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)
    alu_nominal = SM_COUNT * LANES_PER_SM * sm_max * 1e6 / 1e12  # int32 lane-ops/s, TOP/s
    # measured on this GPU, right now (tnt_engine_alu_peak: independent 32-bit adds / min-max / the DP's
    # subtract-then-max pairs); MEASURED_PEAKS.json carries no integer figure
    try:
        alu_measured = eng.alu_peak()
    except Exception as ex:
        alu_measured = {"error": str(ex)}
    alu_peak = alu_measured.get("iadd") or alu_nominal
    align_s = align_ms / args.steps / 1e3
    scan_s = scan_ms / args.steps / 1e3
    alu_achieved = ALU_OPS_PER_CELL * st.dp_cells / align_s / 1e12 if align_s > 0 else 0.0
    scan_achieved = st.scan_bytes / scan_s / 1e9 if scan_s > 0 else 0.0

    # DRAM traffic of the kernels behind the two rooflines: read from the committed ncu capture of one
    # search of this very workload (never typed in); only quoted when the run is that workload
    record_cfg = args.kind == "taqman" and args.mbp == 1000 and args.assays == 100
    nuccruc_traffic = scan_traffic = traffic_file = None
    if record_cfg:
        nuccruc_traffic, scan_traffic, traffic_file = traffic_from_profile()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (DP) + f32 (dH/dS/Tm)", "data": "synthetic",
        "config": {"workload": workload, "db_bases_per_gpu": db_bases, "fragments_per_gpu": len(fragments),
                   "fragment_bases_per_gpu": frag_bases, "assays": len(assays),
                   "l2": "inputs larger than L2 (packed DB %.0f MB + candidate buffers)" % (frag_bases * 0.375 / 1e6),
                   "host_numa_node_rank0": numa_node},
        "alignments_per_s": aligns / dt,
        "alignments_per_step": aligns,
        "dp_cells_per_step": float(st.dp_cells),
        "hits_per_step_rank0": int(nhits),
        "device_ms_per_step": dev_ms / args.steps,
        "kernel_ms_per_step": {"seed_scan": scan_ms / args.steps, "nuccruc_align": align_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": frag_bases, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_dt * 1e3, "steps": e2e_steps,
                "upload_call_ms_per_step": upload_s / e2e_steps * 1e3, "step_ms_rank0": e2e_step_ms,
                "result": "tnt_hit records + alignment strings + amplicon / site text of every hit (tnt_engine_hit_sequences)",
                "hit_text_bytes_per_step": int(seq_bytes)},
        "roofline": {"bound": "alu-int32", "achieved": alu_achieved, "peak": alu_peak, "unit": "TOP/s",
                     "frac": alu_achieved / alu_peak if alu_peak else None, "traffic": nuccruc_traffic,
                     "traffic_note": "DRAM bytes of all NucCruc launches of one search, summed from the committed ncu capture profiles/%s "
                                     "(window fetches are random 32-byte sectors; the kernels are ALU-bound, not memory-bound)" % traffic_file,
                     "algorithmic_bytes": float(aligns / max(world, 1)) * 24.0,
                     "kernel": "k_align (NucCruc DP + traceback + evaluation)",
                     "peak_source": "measured on this GPU in this run: independent 32-bit integer adds (tnt_engine_alu_peak)"
                                    if "iadd" in alu_measured else "nominal",
                     "peak_nominal": alu_nominal, "alu_measured_TOPs": alu_measured,
                     "note": "27 int32 ALU ops per DP cell (SURVEY 8d) x %d cells per step / CUDA-event time of the kernels; "
                             "nominal issue peak = 148 SM x 128 lanes x %.0f MHz" % (st.dp_cells, sm_max)},
        "roofline_seed_scan": {"bound": "hbm", "achieved": scan_achieved, "peak": hbm_peak, "unit": "GB/s",
                               "frac": scan_achieved / hbm_peak if hbm_peak else None, "traffic": scan_traffic,
                               "traffic_note": "k_seed_scan + k_region_scan launches of one search, profiles/%s" % traffic_file,
                               "kernel": "k_seed_scan", "peak_source": hbm_src,
                               "note": "0.375 B per base per pass + 8 B per emitted candidate (SURVEY 8d); with %d oligo "
                                       "strands the scan is bound by table walks and bucket atomics, not HBM" % (2 * len(assays)),
                               "single_assay": scan1},
    }
    if packed is not None:
        line["e2e_packed_snapshot"] = packed
    if ingest is not None:
        if "parse_GBps" in ingest:
            ingest["parse_frac_of_hbm_peak"] = ingest["parse_GBps"] / hbm_peak if hbm_peak else None
        line["ingest_fasta"] = ingest
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU arm gets every host core again (the GPU legs ran on the GPU's NUMA node)
        try:
            os.sched_setaffinity(0, full_affinity)
        except OSError:
            pass
        # the engine (and its HBM) stays alive meanwhile; the engine-backed program creates its own
        line["cpu_baseline"], parity = cpu_baseline(records, reference_assays(assays, args.kind), args.kind)
        if parity is not None:
            line["parity_at_scale"] = parity
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "only measured at N=1"}
    if rank == 0:
        _emit(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line: str) -> None:
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner, ...) also write
    there, so everything else is routed to stderr and the line goes to the original descriptor."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)

if __name__ == "__main__":
    sys.exit(main())
