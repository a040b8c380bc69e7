#pragma once
#include "tnt_types.h"

namespace tnt {

// update_dp_param + pair tables for one (T, [Na+]) -- reference nuc_cruc.cpp:226-487
void build_thermo(Thermo &th, float T, float na, bool dangle5, bool dangle3);

// the same table at another temperature, entry by entry from Thermo::dg_class (Dinkelbach mode)
void dg_at_temperature(const Thermo &th, float T, int32_t *out);

// penalty tables of the fast alignment kernel: out[len][72] and out[20]
void build_row_tables(const Thermo &th, const OligoStrand &os, int32_t *out);
void build_p5_table(const Thermo &th, int32_t *out);
// lean-tier rows out[len][64] from rows[len][72]; false: structure not present, use the full-trace tier
bool build_lean_tables(const Thermo &th, const OligoStrand &os, const int32_t *rows, int32_t *out);

// NC_R*log(Ct) with the host libm (reference nuc_cruc.cpp:2291)
float r_log_ct(float ct);
// param_symmetry_S (nuc_cruc_santa_lucia.cpp): joins the initiation entropy of a homodimer (nuc_cruc.cpp:1632)
float symmetry_S();

// compacted seed word list of an oligo (complement = reverse complement, plus-strand search)
// Smallest number of columns a trimmed gapless alignment of this oligo (first and last column
// Watson-Crick, the only kind the lean alignment tier evaluates) needs to reach a melting
// temperature of min_tm, whatever the target is.  Shorter alignments fail the Tm filter and are
// not evaluated.
int lean_min_columns(const Thermo &th, const OligoStrand &os, float min_tm);

int build_words(const char *oligo, int W, bool complement, uint16_t *words);

int base_from_ascii(char c);
bool complementary(int q, int t);
uint8_t base_set(int b);

} // namespace tnt
