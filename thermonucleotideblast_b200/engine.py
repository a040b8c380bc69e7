"""ctypes binding of the C ABI in include/tntb200.h (libtntb200.so, sm_100a).

This is the host-side entry used by tests/ and bench.py; the reference-side binding a tntblast
maintainer would add is the C++ shim in INTEGRATION.md.  There is no CPU path: creating an
Engine without a usable CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# TNT_LIB: developer hook to load an alternatively built library (kernel variants under test)
LIB_PATH = os.environ.get("TNT_LIB") or os.path.join(HERE, "libtntb200.so")

ASSAY_PCR, ASSAY_PROBE, ASSAY_PADLOCK, ASSAY_MIPS = 0, 1, 2, 3
STRAND_PLUS, STRAND_MINUS, STRAND_BOTH = 1, 2, 3
PLUS, MINUS = 0, 1
OLIGO_F, OLIGO_R, OLIGO_P, OLIGO_NONE = 0, 1, 2, -1
MAX_OLIGO_LEN = 56


class EngineParams(C.Structure):
    _fields_ = [("target_T", C.c_float), ("salt", C.c_float), ("dangle5", C.c_int32),
                ("dangle3", C.c_int32), ("dinkelbach", C.c_int32), ("word_size", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32)]


class SearchOptions(C.Structure):
    _fields_ = [
        ("assay_format", C.c_int32),
        ("forward_primer_strand", C.c_float), ("reverse_primer_strand", C.c_float),
        ("probe_strand", C.c_float),
        ("min_primer_tm", C.c_float), ("max_primer_tm", C.c_float),
        ("min_primer_dg", C.c_float), ("max_primer_dg", C.c_float),
        ("min_probe_tm", C.c_float), ("max_probe_tm", C.c_float),
        ("min_probe_dg", C.c_float), ("max_probe_dg", C.c_float),
        ("primer_clamp", C.c_uint32), ("min_max_primer_clamp", C.c_int32),
        ("probe_clamp_5", C.c_uint32), ("probe_clamp_3", C.c_uint32),
        ("max_gap", C.c_uint32), ("max_mismatch", C.c_uint32), ("max_poly_degen", C.c_uint32),
        ("max_len", C.c_uint32), ("single_primer_pcr", C.c_int32), ("target_strand", C.c_int32),
    ]


def search_options(**kw) -> SearchOptions:
    """Reference CLI defaults (tntblast.h:19-76)."""
    o = SearchOptions(
        assay_format=ASSAY_PCR,
        forward_primer_strand=9.0e-7, reverse_primer_strand=9.0e-7, probe_strand=2.5e-7,
        min_primer_tm=0.0, max_primer_tm=9999.0, min_primer_dg=-9999.0, max_primer_dg=0.0,
        min_probe_tm=0.0, max_probe_tm=9999.0, min_probe_dg=-9999.0, max_probe_dg=0.0,
        primer_clamp=0, min_max_primer_clamp=-1, probe_clamp_5=0, probe_clamp_3=0,
        max_gap=999, max_mismatch=999, max_poly_degen=3, max_len=2000,
        single_primer_pcr=1, target_strand=STRAND_BOTH)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class CAssay(C.Structure):
    _fields_ = [("id", C.c_int32), ("forward", C.c_char_p), ("reverse", C.c_char_p),
                ("probe", C.c_char_p), ("forward_degen", C.c_int32), ("reverse_degen", C.c_int32),
                ("probe_degen", C.c_int32)]


class BoundOligo(C.Structure):
    _fields_ = [("oligo", C.c_int32), ("loc_5", C.c_int32), ("loc_3", C.c_int32),
                ("tm", C.c_float), ("dH", C.c_float), ("dS", C.c_float),
                ("num_mm", C.c_int32), ("num_gap", C.c_int32),
                ("anchor_5", C.c_int32), ("anchor_3", C.c_int32), ("align_off", C.c_uint32)]


class CHit(C.Structure):
    _fields_ = [("assay_index", C.c_int32), ("assay_id", C.c_int32), ("target_id", C.c_uint32),
                ("primer_strand", C.c_int32), ("probe_strand", C.c_int32),
                ("amp_first", C.c_int32), ("amp_last", C.c_int32),
                ("probe_first", C.c_int32), ("probe_last", C.c_int32),
                ("forward", BoundOligo), ("reverse", BoundOligo), ("probe", BoundOligo),
                ("forward_clamp", C.c_int32), ("reverse_clamp", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("db_bases", C.c_uint64), ("seeds", C.c_uint64), ("alignments", C.c_uint64),
                ("dp_cells", C.c_uint64), ("bound_sites", C.c_uint64), ("hits", C.c_uint64),
                ("kernel_launches", C.c_uint64),
                ("scan_ms", C.c_double), ("align_ms", C.c_double), ("pair_ms", C.c_double),
                ("total_ms", C.c_double), ("scan_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("replayed_groups", C.c_uint64), ("undefined_dropped", C.c_uint64),
                ("nonbinding_dropped", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AlignResult(C.Structure):
    _fields_ = [("tm", C.c_float), ("dH", C.c_float), ("dS", C.c_float), ("dG", C.c_float),
                ("valid", C.c_int32), ("anchor5", C.c_int32), ("anchor3", C.c_int32),
                ("num_mismatch", C.c_int32), ("num_gap", C.c_int32), ("max_poly_degen", C.c_int32),
                ("q_first", C.c_int32), ("q_last", C.c_int32), ("t_first", C.c_int32), ("t_last", C.c_int32),
                ("target_start", C.c_int32), ("target_stop", C.c_int32),
                ("loc_5", C.c_int32), ("loc_3", C.c_int32),
                ("alignment", C.c_char * 512)]


class FastaRecord(C.Structure):
    """tnt_fasta_record (include/tntb200.h)."""
    _fields_ = [("text_offset", C.c_uint64), ("text_bytes", C.c_uint64), ("defline_offset", C.c_uint64),
                ("defline_len", C.c_uint32), ("n_fragments", C.c_uint32), ("bases", C.c_uint64),
                ("first_fragment", C.c_uint32), ("pad", C.c_uint32)]


class IngestStats(C.Structure):
    """tnt_ingest_stats (include/tntb200.h)."""
    _fields_ = [("text_bytes", C.c_uint64), ("bases", C.c_uint64), ("records", C.c_uint64), ("fragments", C.c_uint64),
                ("slabs", C.c_uint64), ("launches", C.c_uint64), ("parse_ms", C.c_double), ("call_ms", C.c_double)]


class PackedTarget(C.Structure):
    """tnt_packed_target (include/tntb200.h)."""
    _fields_ = [("base", C.c_uint64), ("len", C.c_uint32), ("pad", C.c_uint32), ("exc_begin", C.c_uint64), ("exc_end", C.c_uint64)]


class PackedInfo(C.Structure):
    """tnt_packed_info (include/tntb200.h)."""
    _fields_ = [("format", C.c_uint32), ("word_size", C.c_uint32), ("n_targets", C.c_uint64), ("n_words", C.c_uint64),
                ("n_exceptions", C.c_uint64), ("next_base", C.c_uint64), ("total_bases", C.c_uint64)]


class FastaFragment(C.Structure):
    """tnt_fasta_fragment (include/tntb200.h)."""
    _fields_ = [("record", C.c_uint32), ("start", C.c_uint32), ("stop", C.c_uint32), ("max_stop", C.c_uint32),
                ("len", C.c_uint32), ("target_id", C.c_uint32)]


class AssayStructures(C.Structure):
    """tnt_assay_structures (include/tntb200.h): index 0 / 1 / 2 = forward primer / reverse primer / probe;
    heterodimer_tm = F with R, F with F, R with R."""
    _fields_ = [("hairpin_tm", C.c_float * 3), ("homodimer_tm", C.c_float * 3), ("heterodimer_tm", C.c_float * 3)]


class Fragment(C.Structure):
    """tnt_fragment (include/tntb200.h): one piece of a database record as the driver cut it."""
    _fields_ = [("record", C.c_uint32), ("start", C.c_uint32), ("stop", C.c_uint32), ("max_stop", C.c_uint32), ("len", C.c_uint32)]


class HitBlock(C.Structure):
    """tnt_hit_block (include/tntb200.h)."""
    _fields_ = [("hits", C.POINTER(CHit)), ("n_hits", C.c_size_t), ("arena", C.c_char_p),
                ("fragments", C.POINTER(Fragment)), ("n_fragments", C.c_size_t)]


class FinalHit(C.Structure):
    """tnt_final_hit (include/tntb200.h)."""
    _fields_ = [("block", C.c_uint32), ("index", C.c_uint32), ("hit", CHit)]


@dataclass
class Assay:
    id: int
    forward: Optional[str] = None
    reverse: Optional[str] = None
    probe: Optional[str] = None
    forward_degen: int = 1
    reverse_degen: int = 1
    probe_degen: int = 1


@dataclass
class Hit:
    """Python view of one tnt_hit with the alignment strings resolved."""
    raw: CHit
    forward_align: str
    reverse_align: str
    probe_align: str

    def __getattr__(self, name):
        return getattr(self.raw, name)


_lib = None


def load_library() -> C.CDLL:
    """Load libtntb200.so; raises if it has not been built (python -m thermonucleotideblast_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libtntb200.so is missing: run `python thermonucleotideblast_b200/build.py` "
                           "(or __graft_entry__.build()); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    u8p = C.POINTER(C.c_uint8)
    u32p = C.POINTER(C.c_uint32)
    L.tnt_last_error.restype = C.c_char_p
    L.tnt_abi_version.restype = C.c_int
    L.tnt_engine_create.argtypes = [C.POINTER(EngineParams), C.POINTER(vp)]
    L.tnt_engine_destroy.argtypes = [vp]
    L.tnt_engine_destroy.restype = None
    L.tnt_engine_add_target.argtypes = [vp, u8p, C.c_uint32, u32p]
    L.tnt_engine_add_targets.argtypes = [vp, C.POINTER(C.c_void_p), u32p, C.c_uint32, u32p]
    L.tnt_engine_clear_targets.argtypes = [vp]
    L.tnt_engine_add_fasta.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                       C.POINTER(C.POINTER(FastaRecord)), C.POINTER(C.c_size_t),
                                       C.POINTER(C.POINTER(FastaFragment)), C.POINTER(C.c_size_t)]
    L.tnt_engine_hits_near_threshold.argtypes = [vp, C.c_float, C.c_float, u32p, C.c_size_t]
    L.tnt_engine_hits_near_threshold.restype = C.c_long
    L.tnt_engine_export_packed.argtypes = [vp, C.POINTER(PackedInfo), vp, vp, vp, vp, vp]
    L.tnt_engine_import_packed.argtypes = [vp, C.POINTER(PackedInfo), vp, vp, vp, vp, vp]
    L.tnt_engine_get_ingest_stats.argtypes = [vp, C.POINTER(IngestStats)]
    L.tnt_engine_target_codes.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, u8p]
    L.tnt_engine_set_assays.argtypes = [vp, C.POINTER(CAssay), C.c_int32]
    L.tnt_engine_search.argtypes = [vp, C.POINTER(SearchOptions)]
    L.tnt_engine_get_hits.argtypes = [vp, C.POINTER(C.POINTER(CHit)), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.tnt_engine_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.tnt_engine_hit_sequence.argtypes = [vp, C.POINTER(CHit), C.c_char_p, C.c_size_t]
    L.tnt_engine_hit_sequence.restype = C.c_long
    L.tnt_engine_hit_sequences.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_size_t)]
    L.tnt_finalize_hits.argtypes = [C.POINTER(HitBlock), C.c_size_t, C.POINTER(CAssay), C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.POINTER(FinalHit)), C.POINTER(C.c_size_t)]
    L.tnt_free.argtypes = [C.c_void_p]
    L.tnt_free.restype = None
    L.tnt_postprocess_error.restype = C.c_char_p
    L.tnt_engine_seeds.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_int32, u32p, u32p, C.c_long]
    L.tnt_engine_seeds.restype = C.c_long
    L.tnt_engine_align.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_int32, C.c_float, u32p, u32p,
                                   C.c_long, C.POINTER(AlignResult)]
    L.tnt_engine_oligo_dimer.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.POINTER(AlignResult)]
    L.tnt_engine_oligo_hairpin.argtypes = [vp, C.c_char_p, C.POINTER(AlignResult)]
    L.tnt_engine_assay_structures.argtypes = [vp, C.POINTER(SearchOptions), C.POINTER(AssayStructures)]
    L.tnt_engine_alu_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.tnt_engine_scan_only.argtypes = [vp, C.POINTER(SearchOptions), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


def c_assays(assays: Sequence["Assay"]):
    """(array of tnt_assay, keep-alive list) for a list of Assay."""
    arr = (CAssay * max(len(assays), 1))()
    keep = []
    for i, a in enumerate(assays):
        f = a.forward.encode() if a.forward else None
        r = a.reverse.encode() if a.reverse else None
        p = a.probe.encode() if a.probe else None
        keep.extend([f, r, p])
        arr[i] = CAssay(a.id, f, r, p, a.forward_degen, a.reverse_degen, a.probe_degen)
    return arr, keep


def finalize_hits(blocks, assays: Sequence["Assay"], best_match: bool = False, uniquify: Optional[bool] = None,
                  raw: bool = False):
    """tnt_finalize_hits: the reference driver's post-processing (truncation filter, record coordinates,
    select_best_match, uniquify_results, sort) over the hit lists of one or several engines.

    `blocks`: list of (hits_bytes, n_hits, arena_bytes, fragments) per engine, where hits_bytes / arena_bytes
    are what Engine.hit_records() returns and `fragments` a list of (record, start, stop, max_stop, len)
    by target id.  Returns a list of (block, index, CHit) in output order, or with raw=True the pair
    (n, bytes of n tnt_final_hit records).  Runs on the host; no GPU."""
    L = load_library()
    nb = len(blocks)
    cb = (HitBlock * max(nb, 1))()
    keep = []
    for i, (hit_bytes, n_hits, arena, frags) in enumerate(blocks):
        hbuf = C.create_string_buffer(hit_bytes, max(len(hit_bytes), 1))
        abuf = C.create_string_buffer(arena, max(len(arena), 1) + 1)
        fr = (Fragment * max(len(frags), 1))(*[Fragment(*f) for f in frags])
        keep.extend([hbuf, abuf, fr])
        cb[i] = HitBlock(C.cast(hbuf, C.POINTER(CHit)), n_hits, C.cast(abuf, C.c_char_p), fr, len(frags))
    arr, keep2 = c_assays(assays)
    out = C.POINTER(FinalHit)()
    n = C.c_size_t()
    mode = -1 if uniquify is None else int(bool(uniquify))
    rc = L.tnt_finalize_hits(cb, nb, arr, len(assays), int(best_match), mode, C.byref(out), C.byref(n))
    if rc < 0:
        raise EngineError(L.tnt_postprocess_error().decode())
    if raw:
        # the finished records as one byte string (n x tnt_final_hit): what a C++ host would keep; no
        # per-hit Python objects (3 us each: 7 s for the 2.3 M hits of bench.py --config 5)
        data = C.string_at(out, n.value * C.sizeof(FinalHit)) if n.value else b""
        L.tnt_free(out)
        return n.value, data
    res = [(out[i].block, out[i].index, CHit.from_buffer_copy(out[i].hit)) for i in range(n.value)]
    L.tnt_free(out)
    return res


class FragmentList:
    """Pointer and length arrays of a list of host fragments, as tnt_engine_add_targets takes them
    (what a C++ host holds anyway).  Keeps the arrays alive."""

    def __init__(self, fragments: Sequence[np.ndarray]):
        self.arrays = [np.ascontiguousarray(f, dtype=np.uint8) for f in fragments]
        self.n = len(self.arrays)
        self.ptrs = (C.c_void_p * max(self.n, 1))(*[a.ctypes.data for a in self.arrays])
        self.lens = (C.c_uint32 * max(self.n, 1))(*[a.size for a in self.arrays])


class Engine:
    """One engine per GPU (and per host thread), like the per-thread DNAHash + NucCruc pair of the
    reference (tntblast_local.cpp:345-372)."""

    def __init__(self, target_T: float = 310.15, salt: float = 50.0e-3, dangle5: bool = False,
                 dangle3: bool = False, word_size: int = 7, device: int = 0, keep_culled_sites: bool = False,
                 dinkelbach: bool = False):
        self.L = load_library()
        prm = EngineParams(target_T=target_T, salt=salt, dangle5=int(dangle5), dangle3=int(dangle3),
                           dinkelbach=int(dinkelbach), word_size=word_size, device=device,
                           reserved=1 if keep_culled_sites else 0)   # TNT_ENGINE_KEEP_CULLED_SITES
        self.h = C.c_void_p()
        self._check(self.L.tnt_engine_create(C.byref(prm), C.byref(self.h)))
        self.target_T = target_T
        self._keep = []

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.L.tnt_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.L.tnt_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- targets ---------------------------------------------------------------------------
    def add_target(self, codes: np.ndarray) -> int:
        a = np.ascontiguousarray(codes, dtype=np.uint8)
        tid = C.c_uint32()
        self._check(self.L.tnt_engine_add_target(self.h, a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size, C.byref(tid)))
        return tid.value

    def add_targets(self, fragments) -> int:
        """Register many fragments with one call; returns the id of the first.  `fragments` is a
        sequence of uint8 arrays or a FragmentList (the marshalled pointer / length arrays, reusable
        when the same host buffers are registered again)."""
        fl = fragments if isinstance(fragments, FragmentList) else FragmentList(fragments)
        first = C.c_uint32()
        self._check(self.L.tnt_engine_add_targets(self.h, fl.ptrs, fl.lens, fl.n, C.byref(first)))
        return first.value

    def add_fasta(self, text, fragment_threshold: int = 500000, overlap: int = 2002, address: int = 0, nbytes: int = 0):
        """Register the records of a FASTA text (bytes / uint8 array, or a raw `address` + `nbytes` of
        e.g. a page-locked buffer): parsed, cut into fragments and packed on the device
        (tnt_engine_add_fasta).  Returns (records, fragments) as lists of ctypes structs."""
        if address:
            ptr, n = C.c_void_p(address), nbytes
        elif isinstance(text, (bytes, bytearray)):
            keep = np.frombuffer(text, dtype=np.uint8)
            ptr, n = C.c_void_p(keep.ctypes.data if keep.size else 0), keep.size
        else:
            keep = np.ascontiguousarray(text, dtype=np.uint8)
            ptr, n = C.c_void_p(keep.ctypes.data if keep.size else 0), keep.size
        pr, pf = C.POINTER(FastaRecord)(), C.POINTER(FastaFragment)()
        nr, nf = C.c_size_t(), C.c_size_t()
        self._check(self.L.tnt_engine_add_fasta(self.h, ptr, n, fragment_threshold, overlap,
                                                C.byref(pr), C.byref(nr), C.byref(pf), C.byref(nf)))
        recs = [FastaRecord.from_buffer_copy(pr[i]) for i in range(nr.value)]
        frags = [FastaFragment.from_buffer_copy(pf[i]) for i in range(nf.value)]
        return recs, frags

    def add_fasta_raw(self, address: int, nbytes: int, fragment_threshold: int = 500000, overlap: int = 2002) -> int:
        """tnt_engine_add_fasta without materialising the tables in Python; returns the fragment count."""
        nf = C.c_size_t()
        self._check(self.L.tnt_engine_add_fasta(self.h, C.c_void_p(address), nbytes, fragment_threshold, overlap,
                                                None, None, None, C.byref(nf)))
        return nf.value

    # -- packed database snapshot ---------------------------------------------------------------
    PACKED_FIELDS = (("targets", np.uint8, 32), ("db2", np.uint64, 1), ("nmask", np.uint32, 1),
                     ("exc_pos", np.uint64, 1), ("exc_code", np.uint8, 1))

    def export_packed(self) -> dict:
        """Resident packed database -> numpy arrays (tnt_engine_export_packed); `np.savez(path, **d)`
        makes the persistent cache file, `import_packed(dict(np.load(path)))` brings it back."""
        info = PackedInfo()
        self._check(self.L.tnt_engine_export_packed(self.h, C.byref(info), None, None, None, None, None))
        n = {"targets": info.n_targets, "db2": info.n_words, "nmask": info.n_words,
             "exc_pos": info.n_exceptions, "exc_code": info.n_exceptions}
        d = {name: np.zeros(max(int(n[name]) * k, 1), dtype=dt)[:int(n[name]) * k] for name, dt, k in self.PACKED_FIELDS}
        self._check(self.L.tnt_engine_export_packed(self.h, C.byref(info), *[C.c_void_p(d[name].ctypes.data) for name, _, _ in self.PACKED_FIELDS]))
        d["info"] = np.array([info.format, info.word_size, info.n_targets, info.n_words, info.n_exceptions,
                              info.next_base, info.total_bases], dtype=np.uint64)
        return d

    def import_packed(self, d: dict):
        """tnt_engine_import_packed from the arrays of export_packed (page-locked arrays are read by DMA
        after the call returns: keep them alive and unchanged until the next search has returned)."""
        i = [int(x) for x in d["info"]]
        info = PackedInfo(i[0], i[1], i[2], i[3], i[4], i[5], i[6])
        arrs = []
        for name, dt, _ in self.PACKED_FIELDS:
            a = d[name]
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                a = np.ascontiguousarray(a, dtype=dt)
            arrs.append(a)
        self._keep = arrs
        self._check(self.L.tnt_engine_import_packed(self.h, C.byref(info), *[C.c_void_p(a.ctypes.data if a.size else 0) for a in arrs]))

    def ingest_stats(self) -> IngestStats:
        st = IngestStats()
        self._check(self.L.tnt_engine_get_ingest_stats(self.h, C.byref(st)))
        return st

    def target_codes(self, target_id: int, start: int, n: int) -> np.ndarray:
        """seq.h codes of a range of a registered fragment, read back from the packed database."""
        out = np.zeros(max(n, 1), dtype=np.uint8)
        self._check(self.L.tnt_engine_target_codes(self.h, target_id, start, n, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out[:n]

    def clear_targets(self):
        self._check(self.L.tnt_engine_clear_targets(self.h))

    # -- assays ----------------------------------------------------------------------------
    def set_assays(self, assays: Sequence[Assay]):
        arr, keep = c_assays(assays)
        self._check(self.L.tnt_engine_set_assays(self.h, arr, len(assays)))

    # -- search ----------------------------------------------------------------------------
    def search(self, opts: SearchOptions) -> List[Hit]:
        self._check(self.L.tnt_engine_search(self.h, C.byref(opts)))
        return self.hits()

    def search_raw(self, opts: SearchOptions) -> int:
        """Search without materialising Python hit objects; returns the hit count."""
        self._check(self.L.tnt_engine_search(self.h, C.byref(opts)))
        n = C.c_size_t()
        self._check(self.L.tnt_engine_get_hits(self.h, None, C.byref(n), None, None))
        return n.value

    def hits(self) -> List[Hit]:
        ph = C.POINTER(CHit)()
        n = C.c_size_t()
        arena = C.c_void_p()
        asz = C.c_size_t()
        self._check(self.L.tnt_engine_get_hits(self.h, C.byref(ph), C.byref(n), C.byref(arena), C.byref(asz)))
        buf = C.string_at(arena.value, asz.value) if asz.value else b""

        def s(off):
            end = buf.index(b"\0", off)
            return buf[off:end].decode()

        out = []
        for i in range(n.value):
            h = CHit.from_buffer_copy(ph[i])
            out.append(Hit(h, s(h.forward.align_off), s(h.reverse.align_off), s(h.probe.align_off)))
        return out

    def hit_records(self):
        """The hit array and the string arena as the C ABI hands them out: copies of the raw bytes
        (tnt_hit records, alignment strings), no per-hit Python objects."""
        ph = C.POINTER(CHit)()
        n = C.c_size_t()
        arena = C.c_void_p()
        asz = C.c_size_t()
        self._check(self.L.tnt_engine_get_hits(self.h, C.byref(ph), C.byref(n), C.byref(arena), C.byref(asz)))
        recs = C.string_at(ph, n.value * C.sizeof(CHit)) if n.value else b""
        text = C.string_at(arena.value, asz.value) if asz.value else b""
        return n.value, recs, text

    def hits_near_threshold(self, tm_tol: float = 0.01, dg_tol: float = 0.001) -> List[int]:
        """Indices of the hits of the last search that sit within the comparison tolerance of a Tm / dG
        bound (to be listed separately when hit sets of two implementations are compared)."""
        n = self.L.tnt_engine_hits_near_threshold(self.h, tm_tol, dg_tol, None, 0)
        if n < 0:
            raise EngineError(self.L.tnt_last_error().decode())
        idx = np.zeros(max(n, 1), dtype=np.uint32)
        self.L.tnt_engine_hits_near_threshold(self.h, tm_tol, dg_tol, idx.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        return idx[:n].tolist()

    def stats(self) -> Stats:
        st = Stats()
        self._check(self.L.tnt_engine_get_stats(self.h, C.byref(st)))
        return st

    def hit_sequence(self, hit: Hit) -> str:
        raw = hit.raw if isinstance(hit, Hit) else hit
        n = self.L.tnt_engine_hit_sequence(self.h, C.byref(raw), None, 0)
        if n < 0:
            raise EngineError(self.L.tnt_last_error().decode())
        buf = C.create_string_buffer(n + 1)
        self.L.tnt_engine_hit_sequence(self.h, C.byref(raw), buf, n + 1)
        return buf.value.decode()

    def hit_sequences(self) -> List[str]:
        """Amplicon / probe-site text of every hit of the last search (tnt_engine_hit_sequences: one
        kernel launch and one device-to-host copy for all of them)."""
        n, text, offs = self.hit_sequences_raw()
        return [text[offs[i]:offs[i + 1] - 1].decode() for i in range(n)]

    def hit_sequences_bytes(self):
        """tnt_engine_hit_sequences without materialising anything in Python: (n_hits, text bytes)."""
        text = C.c_void_p()
        offs = C.POINTER(C.c_uint64)()
        n = C.c_size_t()
        self._check(self.L.tnt_engine_hit_sequences(self.h, C.byref(text), C.byref(offs), C.byref(n)))
        return n.value, (int(offs[n.value]) if n.value else 0)

    def hit_sequences_raw(self):
        """(n_hits, text bytes, offsets) as the C ABI hands them out."""
        text = C.c_void_p()
        offs = C.POINTER(C.c_uint64)()
        n = C.c_size_t()
        self._check(self.L.tnt_engine_hit_sequences(self.h, C.byref(text), C.byref(offs), C.byref(n)))
        o = [offs[i] for i in range(n.value + 1)] if n.value else [0]
        buf = C.string_at(text.value, o[-1]) if o[-1] else b""
        return n.value, buf, o

    def oligo_dimer(self, query: str, target: Optional[str] = None, conc_a: float = 9.0e-7, conc_b: float = 9.0e-7) -> AlignResult:
        """Homodimer (target None) or heterodimer Tm of oligos (tntblast_local.cpp:657-686) on the device."""
        out = AlignResult()
        self._check(self.L.tnt_engine_oligo_dimer(self.h, query.encode(), target.encode() if target else None, conc_a, conc_b, C.byref(out)))
        return out

    def oligo_hairpin(self, query: str) -> AlignResult:
        """Hairpin Tm of an oligo (approximate_tm_hairpin, tntblast_local.cpp:661) on the device."""
        out = AlignResult()
        self._check(self.L.tnt_engine_oligo_hairpin(self.h, query.encode(), C.byref(out)))
        return out

    def assay_structures(self, opts: SearchOptions, n_assays: int):
        """Hairpin / homodimer / heterodimer temperatures of every registered assay (one launch)."""
        out = (AssayStructures * max(n_assays, 1))()
        self._check(self.L.tnt_engine_assay_structures(self.h, C.byref(opts), out))
        return [out[i] for i in range(n_assays)]

    def alu_peak(self):
        """Measured int32 throughput (TOP/s): independent adds, independent min / max, subtract-then-max pairs."""
        t = (C.c_double * 3)()
        self._check(self.L.tnt_engine_alu_peak(self.h, t))
        return {"iadd": t[0], "imnmx": t[1], "sub_max_pairs": t[2]}

    def scan_only(self, opts: SearchOptions):
        """Seed scan of all fragments with the stage-1 oligo strands; returns (candidates, ms)."""
        n = C.c_uint64()
        ms = C.c_double()
        self._check(self.L.tnt_engine_scan_only(self.h, C.byref(opts), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    # -- stage-level entry points ------------------------------------------------------------
    def seeds(self, target_id: int, oligo: str, plus: bool) -> List[Tuple[int, int]]:
        cap = 1 << 16
        while True:
            q = np.zeros(cap, dtype=np.uint32)
            t = np.zeros(cap, dtype=np.uint32)
            n = self.L.tnt_engine_seeds(self.h, target_id, oligo.encode(), int(plus),
                                        q.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        t.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
            if n < 0:
                raise EngineError(self.L.tnt_last_error().decode())
            if n <= cap:
                return list(zip(q[:n].tolist(), t[:n].tolist()))
            cap = n

    def align(self, target_id: int, oligo: str, plus: bool, seeds: Sequence[Tuple[int, int]],
              ct: float = 9.0e-7):
        n = len(seeds)
        q = np.array([s[0] for s in seeds], dtype=np.uint32)
        t = np.array([s[1] for s in seeds], dtype=np.uint32)
        out = (AlignResult * max(n, 1))()
        self._check(self.L.tnt_engine_align(self.h, target_id, oligo.encode(), int(plus), ct,
                                            q.ctypes.data_as(C.POINTER(C.c_uint32)),
                                            t.ctypes.data_as(C.POINTER(C.c_uint32)), n, out))
        return [out[i] for i in range(n)]
