"""Seeded synthetic inputs shared by the parity tests, bench.py and the golden-fixture script.

Shapes follow SURVEY.md section 8(d): uniform random ACGT databases, assays sampled from the
database with planted exact sites plus variants carrying mismatches and a 1-base indel, optional
IUPAC codes / N runs in the target and inosine / IUPAC codes in the oligos.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = str.maketrans("ACGTMRSVWYHKDBNIacgt", "TGCAKYSBWRDMHVNItgca")


def revcomp(s: str) -> str:
    return s.translate(_COMP)[::-1]


def random_codes(n: int, rng: np.random.Generator) -> np.ndarray:
    """seq.h codes 0..3, uniform."""
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def codes_to_str(codes: np.ndarray) -> str:
    lut = np.frombuffer(b"ACGTIMRSVWYHKDBN-?", dtype=np.uint8)
    return lut[codes].tobytes().decode()


def str_to_codes(s: str) -> np.ndarray:
    lut = np.full(256, 17, dtype=np.uint8)
    for i, ch in enumerate("ACGTIMRSVWYHKDBN-"):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    return lut[np.frombuffer(s.encode(), dtype=np.uint8)]


def rand_oligo(L: int, rng: np.random.Generator) -> str:
    return ACGT[rng.integers(0, 4, size=L)].tobytes().decode()


def mutate(s: str, nmut: int, rng: np.random.Generator, indel: bool = True) -> str:
    s = list(s)
    for _ in range(nmut):
        p = int(rng.integers(0, len(s)))
        k = int(rng.integers(0, 4 if indel else 2))
        if k < 2:
            s[p] = "ACGT"[int(rng.integers(0, 4))]
        elif k == 2 and len(s) > 10:
            del s[p]
        else:
            s.insert(p, "ACGT"[int(rng.integers(0, 4))])
    return "".join(s)


def plant(codes: np.ndarray, pos: int, text: str) -> None:
    c = str_to_codes(text)
    n = min(len(c), len(codes) - pos)
    if n > 0:
        codes[pos:pos + n] = c[:n]


def sprinkle_degenerate(codes: np.ndarray, rng: np.random.Generator, frac: float = 1e-3,
                        n_runs_per_50kb: float = 1.0) -> None:
    """0.1 % two/three-fold IUPAC codes + N runs of length 1..10 (config 3 of SURVEY 8d)."""
    n = len(codes)
    k = int(n * frac)
    if k:
        idx = rng.integers(0, n, size=k)
        codes[idx] = rng.integers(5, 15, size=k).astype(np.uint8)  # M..B
    runs = int(n / 50000.0 * n_runs_per_50kb)
    for _ in range(runs):
        p = int(rng.integers(0, n))
        L = int(rng.integers(1, 11))
        codes[p:p + L] = 15


def make_pcr_case(rng: np.random.Generator, n: int, n_sites: int = 3, probe: bool = False,
                  lens: Tuple[int, int, int] = (20, 20, 25), amp: Tuple[int, int] = (80, 400)):
    """One assay + a database of n bases with planted (mutated) amplicons on either strand."""
    codes = random_codes(n, rng)
    F = rand_oligo(lens[0], rng)
    R = rand_oligo(lens[1], rng)
    P = rand_oligo(lens[2], rng) if probe else None
    for k in range(n_sites):
        alen = int(rng.integers(amp[0], amp[1]))
        inner = rand_oligo(alen, rng)
        mid = ""
        if probe:
            pp = P if rng.integers(0, 2) else revcomp(P)
            mid = rand_oligo(int(rng.integers(5, 30)), rng) + mutate(pp, int(rng.integers(0, 3)) if k else 0, rng)
        text = mutate(F, int(rng.integers(0, 4)) if k else 0, rng) + mid + inner + \
            mutate(revcomp(R), int(rng.integers(0, 4)) if k else 0, rng)
        if rng.integers(0, 3) == 0:
            text = revcomp(text)
        pos = int(rng.integers(0, max(1, n - len(text) - 1)))
        plant(codes, pos, text)
    return codes, F, R, P


def make_assays(rng: np.random.Generator, db: List[np.ndarray], n_assays: int, kind: str,
                lens=(20, 21, 25), amp=(80, 400), variants: int = 2):
    """Assays sampled from random oligos, each planted once exactly (+ `variants` mutated copies)
    into random fragments of `db` (modified in place).  kind: pcr | taqman | probe | padlock."""
    assays = []
    for a in range(n_assays):
        F = rand_oligo(int(rng.integers(lens[0] - 2, lens[0] + 3)), rng)
        R = rand_oligo(int(rng.integers(lens[1] - 2, lens[1] + 3)), rng)
        P = rand_oligo(int(rng.integers(lens[2] - 2, lens[2] + 3)), rng)
        for v in range(1 + variants):
            frag = db[int(rng.integers(0, len(db)))]
            nm = 0 if v == 0 else int(rng.integers(1, 4))
            if kind in ("pcr", "taqman"):
                alen = int(rng.integers(amp[0], amp[1]))
                mid = ""
                if kind == "taqman":
                    pp = P if rng.integers(0, 2) else revcomp(P)
                    mid = rand_oligo(int(rng.integers(3, 20)), rng) + mutate(pp, nm if v else 0, rng)
                text = mutate(F, nm, rng) + mid + rand_oligo(alen, rng) + mutate(revcomp(R), nm, rng)
            elif kind == "probe":
                text = mutate(revcomp(P), nm, rng)
            else:  # padlock: 5'-F-3' 5'-R-3' adjacent on one strand
                text = revcomp(mutate(F, nm, rng, indel=False) + mutate(R, nm, rng, indel=False))
            if rng.integers(0, 2):
                text = revcomp(text)
            if len(frag) > len(text) + 2:
                plant(frag, int(rng.integers(0, len(frag) - len(text) - 1)), text)
        if kind == "pcr":
            assays.append((F, R, None))
        elif kind == "taqman":
            assays.append((F, R, P))
        elif kind == "probe":
            assays.append((None, None, P))
        else:
            assays.append((F, R, None))
    return assays


# -- FASTA texts ------------------------------------------------------------------------------
FASTA_EDGE_CASES = [
    b">r1\nACGT\n",
    b"junk in front\nmore junk\n>rec1 first >x\nACGTNNacgu\nRY*-K \tM\n\n>  rec2\r\nAC>GT\r\nTTTT\n>r3\rACGT>AC\nGG\n>r4\nAAAA",
    b">a\nACGT\n>b\n\n>c\nNNNN\n",                       # an empty record
    b">x desc\r\nacgtuACGTU\r\nIMRSVWYHKDBN\r\n",         # CRLF, lower case, RNA, IUPAC
    b">q\nAC GT\tAC\x0bGT\x0cAC*GT-AC\n>p\nZZ..@@12\n",   # blanks, skipped and unknown characters
    b">only header line\n",
    b"\n\n>late start\nACGTACGTAC",                       # no trailing newline
    b">t >u >v\nAC\n >w\nGT\n",                           # several '>' in a defline; '>' after a blank
]


def rand_fasta(rng, n_records=6, max_len=5000, width=60, iupac=0.01, crlf=False, lower=0.1):
    """A random multi-record FASTA text (bytes) with occasional IUPAC codes, lower case and blank lines."""
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for r in range(n_records):
        n = int(rng.integers(0, max_len))
        seq = rng.choice(list(b"ACGT"), size=n).astype(np.uint8)
        m = rng.random(n) < iupac
        seq[m] = rng.choice(list(b"NRYKMSWBDHVI"), size=int(m.sum())).astype(np.uint8)
        lo = rng.random(n) < lower
        seq[lo] |= 0x20
        out.append(b">rec%d some description %d" % (r, n) + eol)
        w = width if width else max(n, 1)
        for i in range(0, n, w):
            out.append(seq[i:i + w].tobytes() + eol)
        if rng.random() < 0.3:
            out.append(eol)
    return b"".join(out)


# -- files for the reference command line ------------------------------------------------------
def write_fasta(path: str, records: List[np.ndarray], names: Optional[List[str]] = None, width: int = 80,
                limit_bp: Optional[int] = None) -> int:
    """80-column FASTA of the records (seq.h codes), optionally only the leading `limit_bp` bases.
    Returns the number of bases written."""
    lut = np.frombuffer(b"ACGTIMRSVWYHKDBN-N", dtype=np.uint8)
    left = sum(len(r) for r in records) if limit_bp is None else limit_bp
    done = 0
    with open(path, "wb") as f:
        for i, rec in enumerate(records):
            if left <= 0:
                break
            n = min(len(rec), left)
            f.write((">%s\n" % (names[i] if names else "rec%d synthetic" % i)).encode())
            txt = lut[rec[:n]]
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = txt[:full].reshape(-1, width)
                body[:, width] = ord("\n")
                f.write(body.tobytes())
            if n > full:
                f.write(txt[full:].tobytes() + b"\n")
            left -= n
            done += n
    return done


def write_assays(path: str, assays) -> None:
    """Assay file of the reference (input.cpp:43-168): name, then F R [P] or P, tab separated."""
    with open(path, "w") as f:
        for i, a in enumerate(assays):
            F, R, P = a[0], a[1], a[2]
            cols = ["assay%d" % i]
            if F and R:
                cols += [F, R]
            if P:
                cols.append(P)
            f.write("\t".join(cols) + "\n")


# -- BASELINE configs[4]: a database every rank can generate piecewise ----------------------------
def config5_assays(n_assays: int):
    arng = np.random.default_rng(55)
    return [(rand_oligo(int(arng.integers(18, 23)), arng), rand_oligo(int(arng.integers(19, 24)), arng), None)
            for _ in range(n_assays)]


def config5_record(r: int, assays, out: np.ndarray):
    """Record r of the synthetic GenBank-scale database: 5 Mbp of uniform bases, seeded by the record
    index so that every rank can make exactly the records it owns, with amplicons of a few assays
    planted (an exact copy and mutated ones; some straddle the 500 kbp cuts)."""
    rng = np.random.default_rng([5, r])
    out[:] = random_codes(len(out), rng)
    for k in range(4):
        F, R, _ = assays[int(rng.integers(0, len(assays)))]
        for v in range(3):
            nm = 0 if v == 0 else int(rng.integers(1, 4))
            text = mutate(F, nm, rng) + rand_oligo(int(rng.integers(80, 400)), rng) + mutate(revcomp(R), nm, rng)
            if rng.integers(0, 2):
                text = revcomp(text)
            pos = int(rng.integers(0, len(out) - len(text) - 1))
            if v == 2:   # on a cut of the record (reference fragment rule), so overlaps and truncations occur
                pos = max(0, min(len(out) - len(text) - 1, int(rng.integers(1, 10)) * 500_000 + int(rng.integers(-len(text), 600))))
            plant(out, pos, text)
