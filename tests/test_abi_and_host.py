"""CPU-side checks: the C ABI library loads and exports every declared symbol, the host-side
table construction matches the oracle, and the engine refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gen
import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tntb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(engine_lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(engine_lib, s), s
    assert engine_lib.tnt_abi_version() == 3


def test_no_cpu_fallback(engine_lib):
    """Without a usable CUDA device creation fails loudly; nothing routes around the kernels."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from thermonucleotideblast_b200 import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine()
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "thermonucleotideblast_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "tnt_oracle" not in txt and "libtntoracle" not in txt and "harness" not in txt, f
                assert "libtntref" not in txt, f


def test_host_thermo_table_matches_oracle(engine_lib, oracle):
    engine_lib.tnt_debug_thermo.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_uint8)]
    for T, na in [(310.15, 0.05), (298.15, 0.2), (333.15, 0.01), (310.15, 1.0)]:
        dg = (C.c_int32 * 2401)()
        bbp = (C.c_uint8 * 324)()
        assert engine_lib.tnt_debug_thermo(T, na, dg, bbp) == 0
        assert list(dg) == list(oracle.dump_tables(T, na).delta_g)
    # invalid salt is refused like NucCruc::salt does (nuc_cruc.h:852-867)
    assert engine_lib.tnt_debug_thermo(310.15, 2.0, None, None) < 0


def test_dinkelbach_rule_table_matches_oracle(engine_lib, oracle):
    """The Dinkelbach kernels re-derive delta_g entry by entry at the temperature of each iteration
    (Thermo::dg_class + the T-independent inputs): the rule table of an engine built at T, evaluated at
    any other temperature, must equal update_dp_param at that temperature."""
    engine_lib.tnt_debug_thermo_at.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int32)]
    rng = np.random.default_rng(12)
    for na in (0.05, 0.2, 1.0):
        for T_eval in [273.15, 310.15, 355.0] + [float(np.float32(273.15 + x)) for x in rng.uniform(0.0, 95.0, size=12)]:
            dg = (C.c_int32 * 2401)()
            assert engine_lib.tnt_debug_thermo_at(310.15, na, T_eval, dg) == 0
            assert list(dg) == list(oracle.dump_tables(T_eval, na).delta_g), (na, T_eval)


def test_host_word_lists_match_oracle_seeds(engine_lib, oracle):
    """The compacted word list (incl. the inosine offset quirk) reproduces the
    oracle's raw seed enumeration on a fragment that contains every word once."""
    engine_lib.tnt_debug_words.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint16)]
    rng = np.random.default_rng(3)
    for it in range(20):
        L = int(rng.integers(7, 40))
        ol = list(gen.rand_oligo(L, rng))
        if it % 2:
            ol[int(rng.integers(0, L))] = "I"
        if it % 5 == 0:
            ol[int(rng.integers(0, L))] = "N"
        ol = "".join(ol)
        W = int(rng.integers(3, 9))
        for comp in (0, 1):
            words = (C.c_uint16 * 64)()
            n = engine_lib.tnt_debug_words(ol.encode(), W, comp, words)
            assert n >= 0
            # expected: distinct consecutive word indices 0..n-1 in the raw seed list of the oracle
            plain = ol.replace("I", "A").replace("N", "A")
            text = gen.revcomp(plain) if comp else plain
            codes = gen.str_to_codes("ACGT" * 3 + text + "TGCA" * 3)
            raw = oracle.seeds(codes, ol, W, bool(comp), unique=False)
            assert (max(q for q, _ in raw) + 1 if raw else 0) <= n
            # every word of the list must be the W-mer of the seed string at its true offset
            s = gen.revcomp(ol) if comp else ol
            k = 0
            for off in range(len(s) - W + 1):
                w = s[off:off + W]
                if all(c in "ACGT" for c in w):
                    val = 0
                    for c in w:
                        val = (val << 2) | "ACGT".index(c)
                    assert words[k] == val
                    k += 1
            assert k == n


def test_min_columns_bound_against_the_oracle(engine_lib, oracle):
    """The lean alignment tier does not evaluate gapless alignments that are too short to reach the
    Tm threshold (lean_min_columns, thermo.cpp).  Check the bound against the oracle: the most
    stable gapless duplex of n columns is a perfect match, so for every n below the bound and every
    window of the oligo, the oracle's Tm of oligo vs (perfect complement of the window, blocked on
    both sides) must stay below the threshold."""
    import numpy as np
    import gen
    f = engine_lib.tnt_debug_min_columns
    f.argtypes = [C.c_float, C.c_float, C.c_char_p, C.c_float, C.c_float]
    f.restype = C.c_int
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rng = np.random.default_rng(8128)
    checked = 0
    for (T, na, ct, min_tm) in [(310.15, 0.05, 9.0e-7, 45.0), (310.15, 0.05, 2.5e-7, 50.0), (330.15, 0.1, 9.0e-7, 40.0)]:
        for _ in range(12):
            L = int(rng.integers(18, 31))
            ol = gen.rand_oligo(L, rng)
            b = f(T, na, ol.encode(), ct, min_tm)
            assert 3 <= b <= min(L, 24) + 1
            # thresholds that accept Tm = 0 outcomes switch the shortcut off
            assert f(T, na, ol.encode(), ct, 0.0) == 0
            for n in range(max(3, b - 4), b):
                for x in range(0, L - n + 1):
                    window = ol[x:x + n]
                    target = "".join(comp[c] for c in reversed(window))
                    # a base never pairs with itself: block the extension on both sides
                    left = ol[x + n] if x + n < L else "A"     # faces the oligo base after the window
                    right = ol[x - 1] if x > 0 else "A"        # faces the oligo base before it
                    tb = np.array([code[c] for c in (left + target + right)], dtype=np.uint8)
                    a = oracle.align(ol, tb, T=T, na=na, ct=ct)
                    assert (not a.valid) or a.tm < min_tm, (ol, n, x, a.tm, b)
                    checked += 1
    assert checked > 1000


def test_replay_array_form_equals_list_form(engine_lib):
    """The replay of the reference's staged PCR search (assemble.cpp) exists twice: literally with
    std::list, and on arrays with the cheap route wherever the list's comparator is a strict weak
    ordering.  Random match lists with overlapping bound sites through both: identical hit lists."""
    engine_lib.tnt_debug_replay_selftest.argtypes = [C.c_uint32, C.c_int32, C.POINTER(C.c_long)]
    engine_lib.tnt_debug_replay_selftest.restype = C.c_long
    hits = C.c_long()
    for seed in (1, 2, 3):
        assert engine_lib.tnt_debug_replay_selftest(seed, 20000, C.byref(hits)) == 0
        assert hits.value > 100000
