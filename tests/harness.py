"""ctypes bindings for the TEST-ONLY checkers.

* ``ref``    -> oracle/_ref/libtntref.so : the unmodified reference compiled from /root/reference
                (recipe: oracle/Makefile).  Ground truth.
* ``oracle`` -> oracle/libtntoracle.so    : the plain-C restatement (oracle/tnt_oracle.c).

Both export the same record layouts (oracle/ref_harness.h).  Only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libtntref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "tntblast")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libtntoracle.so")

TABLE = 49 * 49

# seq.h:12-33 target codes
DB = {c: i for i, c in enumerate("ACGTIMRSVWYHKDBN")}
DB["-"] = 16
DB_UNKNOWN = 17
DB_LETTERS = "ACGTIMRSVWYHKDBN-?"

ASSAY_PCR, ASSAY_PROBE, ASSAY_PADLOCK, ASSAY_MIPS = 0, 1, 2, 3
STRAND_PLUS, STRAND_MINUS, STRAND_BOTH = 1, 2, 3


class Tables(C.Structure):
    _fields_ = [
        ("delta_g", C.c_int32 * TABLE),
        ("param_H", C.c_float * TABLE),
        ("param_S", C.c_float * TABLE),
        ("loop_terminal_H", C.c_float * TABLE),
        ("loop_terminal_S", C.c_float * TABLE),
        ("loop_S", C.c_float * 513),
        ("bulge_S", C.c_float * 513),
        ("supp", C.c_float * 12),
        ("supp_salt", C.c_float * 4),
        ("init_H", C.c_float),
        ("init_S", C.c_float),
        ("AT_closing_H", C.c_float),
        ("AT_closing_S", C.c_float),
        ("symmetry_S", C.c_float),
        ("SALT", C.c_float),
        ("asymmetric_loop_dS", C.c_float),
        ("bulge_AT_closing_S", C.c_float),
        ("watson_and_crick", C.c_uint8 * 49),
    ]


class AlignOut(C.Structure):
    _fields_ = [
        ("tm", C.c_float), ("dH", C.c_float), ("dS", C.c_float), ("dG", C.c_float),
        ("dp_dg", C.c_float),
        ("valid", C.c_int32),
        ("anchor5", C.c_int32), ("anchor3", C.c_int32),
        ("num_mismatch", C.c_int32), ("num_gap", C.c_int32), ("max_poly_degen", C.c_int32),
        ("q_first", C.c_int32), ("q_last", C.c_int32), ("t_first", C.c_int32), ("t_last", C.c_int32),
        ("target_start", C.c_int32), ("target_stop", C.c_int32),
        ("loc_5", C.c_int32), ("loc_3", C.c_int32),
        ("alignment", C.c_char * 512),
    ]

    def key(self):
        """Everything that must be bit-exact for a valid alignment."""
        if not self.valid:
            return (0,)
        return (1, self.anchor5, self.anchor3, self.num_mismatch, self.num_gap,
                self.max_poly_degen, self.q_first, self.q_last, self.t_first, self.t_last,
                self.alignment.decode())


class Options(C.Structure):
    _fields_ = [
        ("assay_format", C.c_int32), ("word_size", C.c_int32),
        ("target_T", C.c_float), ("salt", C.c_float),
        ("dangle5", C.c_int32), ("dangle3", C.c_int32),
        ("forward_primer_strand", C.c_float), ("reverse_primer_strand", C.c_float),
        ("probe_strand", C.c_float),
        ("min_primer_tm", C.c_float), ("max_primer_tm", C.c_float),
        ("min_primer_dg", C.c_float), ("max_primer_dg", C.c_float),
        ("min_probe_tm", C.c_float), ("max_probe_tm", C.c_float),
        ("min_probe_dg", C.c_float), ("max_probe_dg", C.c_float),
        ("primer_clamp", C.c_uint32), ("min_max_primer_clamp", C.c_int32),
        ("probe_clamp_5", C.c_uint32), ("probe_clamp_3", C.c_uint32),
        ("max_gap", C.c_uint32), ("max_mismatch", C.c_uint32), ("max_poly_degen", C.c_uint32),
        ("max_len", C.c_uint32),
        ("single_primer_pcr", C.c_int32), ("target_strand", C.c_int32),
    ]


def default_options(**kw) -> Options:
    """Defaults of the reference CLI (tntblast.h:19-76, options.h)."""
    o = Options(
        assay_format=ASSAY_PCR, word_size=7, target_T=310.15, salt=50.0e-3,
        dangle5=0, dangle3=0,
        forward_primer_strand=9.0e-7, reverse_primer_strand=9.0e-7, probe_strand=2.5e-7,
        min_primer_tm=0.0, max_primer_tm=9999.0, min_primer_dg=-9999.0, max_primer_dg=0.0,
        min_probe_tm=0.0, max_probe_tm=9999.0, min_probe_dg=-9999.0, max_probe_dg=0.0,
        primer_clamp=0, min_max_primer_clamp=-1, probe_clamp_5=0, probe_clamp_3=0,
        max_gap=999, max_mismatch=999, max_poly_degen=3, max_len=2000,
        single_primer_pcr=1, target_strand=STRAND_BOTH,
    )
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class Hit(C.Structure):
    _fields_ = [
        ("primer_strand", C.c_int32), ("probe_strand", C.c_int32),
        ("amp_first", C.c_int32), ("amp_last", C.c_int32),
        ("probe_first", C.c_int32), ("probe_last", C.c_int32),
        ("forward_tm", C.c_float), ("forward_dH", C.c_float), ("forward_dS", C.c_float),
        ("reverse_tm", C.c_float), ("reverse_dH", C.c_float), ("reverse_dS", C.c_float),
        ("probe_tm", C.c_float), ("probe_dH", C.c_float), ("probe_dS", C.c_float),
        ("forward_mm", C.c_int32), ("forward_gap", C.c_int32),
        ("reverse_mm", C.c_int32), ("reverse_gap", C.c_int32),
        ("probe_mm", C.c_int32), ("probe_gap", C.c_int32),
        ("forward_clamp", C.c_int32), ("reverse_clamp", C.c_int32),
        ("amplicon_len", C.c_int32),
        ("amplicon_fnv", C.c_uint64),
        ("forward_oligo", C.c_char * 128), ("reverse_oligo", C.c_char * 128),
        ("forward_align", C.c_char * 512), ("reverse_align", C.c_char * 512),
        ("probe_align", C.c_char * 512),
        ("amplicon_head", C.c_char * 256),
    ]

    def exact_key(self):
        """Integer/string part of a hit: must be bit-exact."""
        return (self.primer_strand, self.probe_strand, self.amp_first, self.amp_last,
                self.probe_first, self.probe_last,
                self.forward_mm, self.forward_gap, self.reverse_mm, self.reverse_gap,
                self.probe_mm, self.probe_gap, self.forward_clamp, self.reverse_clamp,
                self.amplicon_len, self.amplicon_fnv,
                self.forward_oligo, self.reverse_oligo,
                self.forward_align, self.reverse_align, self.probe_align, self.amplicon_head)

    def floats(self):
        return (self.forward_tm, self.forward_dH, self.forward_dS,
                self.reverse_tm, self.reverse_dH, self.reverse_dS,
                self.probe_tm, self.probe_dH, self.probe_dS)


def encode(seq: str) -> np.ndarray:
    """ASCII -> seq.h target codes (ascii_to_hash_base, seq.h:148-189)."""
    lut = np.full(256, DB_UNKNOWN, dtype=np.uint8)
    for ch, v in DB.items():
        lut[ord(ch)] = v
        lut[ord(ch.lower())] = v
    lut[ord("U")] = lut[ord("u")] = DB["T"]
    return lut[np.frombuffer(seq.encode(), dtype=np.uint8)]


def decode(codes: np.ndarray) -> str:
    return "".join(DB_LETTERS[c] for c in codes)


_COMP = str.maketrans("ACGTMRSVWYHKDBNIacgt", "TGCAKYSBWRDMHVNItgca")


def revcomp(s: str) -> str:
    return s.translate(_COMP)[::-1]


class PostHit(C.Structure):
    """ref_post_hit (oracle/ref_harness.h)."""
    _fields_ = [("id", C.c_int32), ("degen_id", C.c_int32), ("seq_id", C.c_int32),
                ("has_primers", C.c_int32), ("has_probe", C.c_int32),
                ("amp_first", C.c_int32), ("amp_last", C.c_int32), ("probe_first", C.c_int32), ("probe_last", C.c_int32),
                ("forward_tm", C.c_float), ("reverse_tm", C.c_float), ("probe_tm", C.c_float),
                ("forward_len", C.c_int32), ("reverse_len", C.c_int32),
                ("forward_align", C.c_char_p), ("reverse_align", C.c_char_p), ("probe_align", C.c_char_p)]


class _Lib:
    """Common binding: the reference harness and the C restatement export the same symbols
    with a different prefix (``ref_`` / ``orc_``)."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        u8p = C.POINTER(C.c_uint8)
        u32p = C.POINTER(C.c_uint32)

        def fn(name, res, args):
            f = getattr(L, prefix + name)
            f.restype = res
            f.argtypes = args
            return f

        self._last_error = fn("last_error", C.c_char_p, [])
        self._set_dinkelbach = fn("set_dinkelbach", None, [C.c_int])
        self._dump_tables = fn("dump_tables", C.c_int, [C.c_float, C.c_float, C.POINTER(Tables)])
        self._seeds_raw = fn("seeds_raw", C.c_long, [u8p, C.c_uint32, C.c_int, C.c_char_p, C.c_int, u32p, u32p, C.c_long])
        self._seeds_unique = fn("seeds_unique", C.c_long, [u8p, C.c_uint32, C.c_int, C.c_char_p, C.c_int, u32p, u32p, C.c_long])
        self._align = fn("align", C.c_int, [C.c_char_p, u8p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(AlignOut)])
        self._bind_window = fn("bind_window", C.c_int, [u8p, C.c_uint32, C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(AlignOut)])
        self._search = fn("search", C.c_long, [u8p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(Options)])
        self._get_hits = fn("get_hits", C.c_int, [C.POINTER(Hit), C.c_long])
        self._dimer = fn("dimer", C.c_int, [C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(AlignOut)])
        u64p = C.POINTER(C.c_uint64)
        self._seq_len_increment = fn("seq_len_increment", None, [C.c_uint32, C.c_uint32, u32p, u32p])
        if prefix == "orc_":
            self._fasta_index = fn("fasta_index", C.c_long, [C.c_char_p, C.c_uint64, u64p, C.c_long])
            self._fasta_read = fn("fasta_read", C.c_long, [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                                         u8p, C.c_long, u64p, u32p])
            self._fragments = fn("fragments", C.c_long, [C.c_uint32, C.c_uint32, u32p, u32p, C.c_long])
        else:
            self._fasta_open = fn("fasta_open", C.c_long, [C.c_char_p])
            self._fasta_approx_len = fn("fasta_approx_len", C.c_long, [C.c_long])
            self._fasta_read_file = fn("fasta_read", C.c_long, [C.c_long, C.c_uint32, C.c_uint32, C.c_int, u8p, C.c_long,
                                                              C.c_char_p, C.c_long])
            self._fasta_close = fn("fasta_close", None, [])

    def set_dinkelbach(self, on: bool) -> None:
        """NucCruc::dinkelbach(on) for every later call of this thread (nuc_cruc.h:763-766)."""
        self._set_dinkelbach(1 if on else 0)

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(self._last_error().decode())
        return rc

    @staticmethod
    def _u8(a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return a, a.ctypes.data_as(C.POINTER(C.c_uint8))

    def dump_tables(self, T=310.15, na=0.05) -> Tables:
        t = Tables()
        self._check(self._dump_tables(T, na, C.byref(t)))
        return t

    def seeds(self, codes, oligo: str, word=7, plus=False, unique=True) -> List[Tuple[int, int]]:
        a, p = self._u8(codes)
        f = self._seeds_unique if unique else self._seeds_raw
        cap = 1 << 16
        while True:
            q = np.zeros(cap, dtype=np.uint32)
            t = np.zeros(cap, dtype=np.uint32)
            n = self._check(f(p, len(a), word, oligo.encode(), int(plus),
                              q.ctypes.data_as(C.POINTER(C.c_uint32)),
                              t.ctypes.data_as(C.POINTER(C.c_uint32)), cap))
            if n <= cap:
                return list(zip(q[:n].tolist(), t[:n].tolist()))
            cap = n

    def align(self, query: str, target_bases, T=310.15, na=0.05, ct=9.0e-7, dangle5=0, dangle3=0) -> AlignOut:
        a, p = self._u8(target_bases)
        out = AlignOut()
        self._check(self._align(query.encode(), p, len(a), T, na, ct, dangle5, dangle3, C.byref(out)))
        return out

    def bind_window(self, codes, oligo: str, plus: bool, q: int, t: int, T=310.15, na=0.05,
                    ct=9.0e-7, dangle5=0, dangle3=0) -> AlignOut:
        a, p = self._u8(codes)
        out = AlignOut()
        self._check(self._bind_window(p, len(a), oligo.encode(), int(plus), q, t, T, na, ct,
                                      dangle5, dangle3, C.byref(out)))
        return out

    def search(self, codes, forward: Optional[str], reverse: Optional[str], probe: Optional[str],
               opts: Options, degen: Sequence[int] = (1, 1, 1)) -> List[Hit]:
        a, p = self._u8(codes)
        n = self._check(self._search(p, len(a), (forward or "").encode(), (reverse or "").encode(),
                                     (probe or "").encode(), degen[0], degen[1], degen[2], C.byref(opts)))
        arr = (Hit * max(n, 1))()
        got = self._get_hits(arr, n)
        return [arr[i] for i in range(got)]


    def hairpin(self, query: str, T=310.15, na=0.05) -> AlignOut:
        """approximate_tm_hairpin of an oligo (tntblast_local.cpp:661); q/t_first = bases closing the loop,
        num_gap = columns of the stem."""
        f = getattr(self.lib, self.prefix + "hairpin")
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_float, C.c_float, C.POINTER(AlignOut)]
        out = AlignOut()
        self._check(f(query.encode(), T, na, C.byref(out)))
        return out

    def finalize(self, hits: Sequence[PostHit], best_match: bool, uniquify: bool) -> List[int]:
        """One result list through the reference's select_best_match / uniquify_results / sort (only the
        compiled reference exports it): indices of the surviving records in output order."""
        f = getattr(self.lib, self.prefix + "finalize")
        f.restype = C.c_long
        f.argtypes = [C.POINTER(PostHit), C.c_long, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        n = len(hits)
        arr = (PostHit * max(n, 1))(*hits)
        out = (C.c_int32 * max(n, 1))()
        k = self._check(f(arr, n, int(best_match), int(uniquify), out))
        return [out[i] for i in range(k)]

    def dimer(self, query: str, target: Optional[str] = None, T=310.15, na=0.05, conc_a=9.0e-7, conc_b=9.0e-7) -> AlignOut:
        """Homodimer (target None) or heterodimer Tm of oligos, as tntblast_local.cpp:657-686 computes them."""
        out = AlignOut()
        self._check(self._dimer(query.encode(), target.encode() if target is not None else None, T, na, conc_a, conc_b, C.byref(out)))
        return out

    # -- FASTA reader ------------------------------------------------------------------------
    def seq_len_increment(self, length: int, max_len: int) -> Tuple[int, int]:
        d, n = C.c_uint32(), C.c_uint32()
        self._seq_len_increment(length, max_len, C.byref(d), C.byref(n))
        return d.value, n.value

    def fasta_records(self, text: bytes, path: Optional[str] = None, threshold: int = 0, overlap: int = 0):
        """Parse a FASTA text like the reference: per record (offset, approx_len, defline, codes of
        the whole record, [(start, stop, codes of [start, stop + overlap])] per driver fragment).
        The oracle works on the bytes, the compiled reference on the file `path` holding them."""
        out = []
        if self.prefix == "orc_":
            cap = max(text.count(b">"), 1)
            pos = np.zeros(cap, dtype=np.uint64)
            n = self._fasta_index(text, len(text), pos.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
            ends = list(pos[1:n]) + [len(text)]
            for i in range(n):
                b, e = int(pos[i]), int(ends[i])
                buf = np.zeros(max(e - b, 1), dtype=np.uint8)
                d0, dl = C.c_uint64(), C.c_uint32()

                def read(start, stop):
                    m = self._fasta_read(text, b, e, start, stop, buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size,
                                         C.byref(d0), C.byref(dl))
                    if m < 0:
                        raise RuntimeError("Truncated fasta file detected!")
                    return buf[:m].copy()

                whole = read(0, 0xffffffff)
                defline = text[d0.value:d0.value + dl.value].decode("latin-1")
                out.append((b, e - b, defline, whole, self._pieces(e - b, threshold, overlap, read)))
        else:
            n = self._check(self._fasta_open(path.encode()))
            try:
                for i in range(n):
                    alen = self._fasta_approx_len(i)
                    buf = np.zeros(max(alen, 1), dtype=np.uint8)
                    dbuf = C.create_string_buffer(4096)

                    def read(start, stop, whole=0):
                        m = self._check(self._fasta_read_file(i, start, stop, whole, buf.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                              buf.size, dbuf, 4096))
                        return buf[:m].copy()

                    whole = read(0, 0, 1)
                    out.append((None, alen, dbuf.value.decode("latin-1"), whole, self._pieces(alen, threshold, overlap, read)))
            finally:
                self._fasta_close()
        return out

    def _pieces(self, alen, threshold, overlap, read):
        """tntblast_local.cpp:282-289,448-468: the (start, stop) queue of one record."""
        if not threshold:
            return []
        delta, _ = self.seq_len_increment(alen, threshold)
        max_stop = alen - 1
        start, stop = 0, delta
        pieces = []
        while True:
            pieces.append((start, stop, read(start, stop + overlap)))
            if stop == max_stop:
                break
            start = stop + 1
            stop = min(stop + delta, max_stop)
        return pieces


_ref = None
_oracle = None


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def ref() -> _Lib:
    global _ref
    if _ref is None:
        _ref = _Lib(REF_LIB, "ref_")
    return _ref


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def oracle() -> _Lib:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            build_oracle()
        _oracle = _Lib(ORACLE_LIB, "orc_")
    return _oracle
