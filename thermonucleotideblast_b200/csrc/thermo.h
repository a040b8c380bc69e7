#pragma once
#include "tnt_types.h"

namespace tnt {

// update_dp_param + pair tables for one (T, [Na+]) -- reference nuc_cruc.cpp:226-487
void build_thermo(Thermo &th, float T, float na, bool dangle5, bool dangle3);

// NC_R*log(Ct) with the host libm (reference nuc_cruc.cpp:2291)
float r_log_ct(float ct);

// compacted seed word list of an oligo (complement = reverse complement, plus-strand search)
int build_words(const char *oligo, int W, bool complement, uint16_t *words);

int base_from_ascii(char c);
bool complementary(int q, int t);
uint8_t base_set(int b);

} // namespace tnt
