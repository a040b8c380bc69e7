// Host-side construction of the thermodynamic tables the kernels consume.
//
// Mirrors what the reference computes once per thread in NucCruc::NucCruc + Salt() +
// update_dp_param() (nuc_cruc.cpp:226-487): the T/salt-dependent integer penalty table and the
// pair-resolution table for IUPAC bases (nuc_cruc.cpp:14-213).  The parameter values themselves
// are data (santalucia_tables.inc, exported from the reference by oracle/gen_tables.py).
//
// Everything here is IEEE binary32 evaluated in the same order as the reference so that the
// integer truncation `int(x*10000.0f)` lands on the same value; the file is compiled without
// FP contraction.

#include "thermo.h"

#include <cmath>
#include <vector>
#include <cstring>
#include <stdexcept>

#include "santalucia_tables.inc"

namespace tnt {

namespace {

constexpr int pair_of(int x, int y) { return x*7 + y; }
constexpr int sidx(int prev, int cur) { return prev*NPAIR + cur; }

// A/C/G/T membership of each IUPAC code as a 4-bit set (A=1, C=2, G=4, T=8).
// `B` shares N's set and default: the reference's switch falls through from B into N
// (nuc_cruc.cpp:163-197).
struct DegenRule { uint8_t set; uint8_t fallback; };
const DegenRule kDegen[NB] = {
	{0, bA}, {0, bC}, {0, bG}, {0, bT}, {0, bI}, {0, bE}, {0, bGAP},
	{1 | 2, bA},      // M
	{1 | 4, bA},      // R
	{2 | 4, bG},      // S
	{1 | 2 | 4, bA},  // V
	{1 | 8, bA},      // W
	{2 | 8, bT},      // Y
	{1 | 2 | 8, bA},  // H
	{4 | 8, bT},      // K
	{1 | 4 | 8, bA},  // D
	{15, bA},         // B (== N)
	{15, bA},         // N
};

// The most optimistic concrete base for `x` when it faces `other`.
int resolve(int x, int other)
{
	if (x < bM) return x;
	if (other <= bT) {
		const int wanted = 3 - other; // Watson-Crick partner in the 0..3 code
		if (kDegen[x].set & (1u << wanted)) return wanted;
	}
	return kDegen[x].fallback;
}

inline int32_t scaled(float x) { return (int32_t)(x*10000.0f); } // NC_SCORE_SCALE, nuc_cruc.h:166
inline int32_t unfavourable(float x) { const int32_t v = scaled(x); return v > 0 ? v : 0; }

} // namespace

uint8_t base_set(int b)
{
	static const uint8_t kSet[NB] = {1, 2, 4, 8, 15, 0, 0, 1 | 2, 1 | 4, 2 | 4, 1 | 2 | 4, 1 | 8, 2 | 8,
		1 | 2 | 8, 4 | 8, 1 | 4 | 8, 2 | 4 | 8, 15};
	return kSet[b];
}

bool complementary(int q, int t) // BASE::is_complemetary_base, nuc_cruc_anchor.cpp:8-139
{
	const unsigned ts = base_set(t);
	const unsigned tc = ((ts & 1) << 3) | ((ts & 8) >> 3) | ((ts & 2) << 1) | ((ts & 4) >> 1);
	return (base_set(q) & tc) != 0;
}

int base_from_ascii(char c) // BASE::char_to_nucleic_acid, nuc_cruc.h:190-231
{
	switch (c) {
	case 'A': case 'a': return bA;
	case 'C': case 'c': return bC;
	case 'G': case 'g': return bG;
	case 'T': case 't': return bT;
	case 'I': case 'i': return bI;
	case 'M': case 'm': return bM;
	case 'R': case 'r': return bR;
	case 'S': case 's': return bS;
	case 'V': case 'v': return bV;
	case 'W': case 'w': return bW;
	case 'Y': case 'y': return bY;
	case 'H': case 'h': return bH;
	case 'K': case 'k': return bK;
	case 'D': case 'd': return bD;
	case 'B': case 'b': return bB;
	case 'N': case 'n': return bN;
	default: return -1;
	}
}

void build_thermo(Thermo &th, float T, float na, bool dangle5, bool dangle3)
{
	if (T < 0.0f) throw std::runtime_error("update_dp_param: target_T < 0");
	if (na < 1.0e-6f) throw std::runtime_error(":salt: [Na+] < 1.0e-6f");
	if (na > 1.0f) throw std::runtime_error(":salt: [Na+] > 1.0f");

	std::memset(&th, 0, sizeof(th));
	th.T = T;
	th.log_na = std::log(na); // float overload == logf, as in the reference translation unit
	th.init_H = SL_INIT_H;
	th.init_S = SL_INIT_S;
	th.at_H = SL_AT_CLOSING_H;
	th.at_S = SL_AT_CLOSING_S;
	th.salt = SL_SALT;
	th.asym_loop_dS = SL_ASYMMETRIC_LOOP_DS;
	th.bulge_at_S = SL_BULGE_AT_CLOSING_S;
	th.dangle5 = dangle5;
	th.dangle3 = dangle3;
	std::memcpy(th.H, SL_PARAM_H, sizeof(th.H));
	std::memcpy(th.S, SL_PARAM_S, sizeof(th.S));
	std::memcpy(th.loop_S, SL_LOOP_S, sizeof(th.loop_S));
	std::memcpy(th.bulge_S, SL_BULGE_S, sizeof(th.bulge_S));
	static_assert(SL_NUM_HAIRPIN_LOOP == NUM_HAIRPIN_LOOP, "special hairpin loops");
	std::memcpy(th.hairpin_S, SL_HAIRPIN_S, sizeof(th.hairpin_S));
	std::memcpy(th.hairpin_loop, SL_HAIRPIN_LOOP, sizeof(th.hairpin_loop));
	std::memcpy(th.hairpin_special_H, SL_HAIRPIN_SPECIAL_H, sizeof(th.hairpin_special_H));
	std::memcpy(th.hairpin_special_S, SL_HAIRPIN_SPECIAL_S, sizeof(th.hairpin_special_S));

	for (int x = 0; x < NB; ++x)
		for (int y = 0; y < NB; ++y)
			th.bbp[x*NB + y] = (uint8_t)pair_of(resolve(x, y), resolve(y, x));

	// Watson-Crick pairs, inosine pairs with everything (nuc_cruc.cpp:229-238)
	th.wc[pair_of(bA, bT)] = th.wc[pair_of(bT, bA)] = 1;
	th.wc[pair_of(bC, bG)] = th.wc[pair_of(bG, bC)] = 1;
	for (int b = bA; b <= bI; ++b) th.wc[pair_of(b, bI)] = th.wc[pair_of(bI, b)] = 1;

	const float salt_correction = SL_SALT*th.log_na;
	for (int i = 0; i < TABLE; ++i) th.dg[i] = scaled(SL_PARAM_H[i] - T*(SL_PARAM_S[i] + salt_correction));
	th.salt_correction = salt_correction;
	std::memcpy(th.supp, SL_SUPP, sizeof(th.supp));
	for (int k = 0; k < 4; ++k) th.supp_sc[k] = salt_correction*SL_SUPP_SALT[k];

	// Supplementary terms (nuc_cruc.cpp:271-300, :379-486): pairs next to a gap, double
	// mismatches and gap extension; never favourable.
	const float loop_sc = salt_correction*SL_SUPP_SALT[0];
	const float bulge_sc = salt_correction*SL_SUPP_SALT[1];
	const float match_sc = salt_correction*SL_SUPP_SALT[2];
	const float mismatch_sc = salt_correction*SL_SUPP_SALT[3];
	const int32_t pen_loop = unfavourable(SL_SUPP[0] - T*(SL_SUPP[1] + loop_sc));
	const int32_t pen_bulge = unfavourable(SL_SUPP[2] - T*(SL_SUPP[3] + bulge_sc));
	const int32_t pen_at = unfavourable(SL_SUPP[4] - T*(SL_SUPP[5] + match_sc));
	const int32_t pen_gc = unfavourable(SL_SUPP[6] - T*(SL_SUPP[7] + match_sc));
	const int32_t pen_ino = unfavourable(SL_SUPP[8] - T*(SL_SUPP[9] + match_sc));
	const int32_t pen_mm = unfavourable(SL_SUPP[10] - T*(SL_SUPP[11] + mismatch_sc));

	for (int x = bA; x <= bI; ++x)
		for (int y = bA; y <= bI; ++y) {
			const int cur = pair_of(x, y);
			int32_t next_to_gap = pen_mm;
			uint8_t cls = DG_TERM_MM;
			if (th.wc[cur]) {
				const bool at = (cur == pair_of(bA, bT)) || (cur == pair_of(bT, bA));
				const bool gc = (cur == pair_of(bG, bC)) || (cur == pair_of(bC, bG));
				next_to_gap = at ? pen_at : (gc ? pen_gc : pen_ino);
				cls = at ? DG_TERM_AT : (gc ? DG_TERM_GC : DG_TERM_INO);
			}
			for (int k = bA; k <= bI; ++k) {
				const int g1 = pair_of(k, bGAP), g2 = pair_of(bGAP, k);
				th.dg[sidx(cur, g1)] = th.dg[sidx(g1, cur)] = next_to_gap;
				th.dg[sidx(cur, g2)] = th.dg[sidx(g2, cur)] = next_to_gap;
				th.dg_class[sidx(cur, g1)] = th.dg_class[sidx(g1, cur)] = cls;
				th.dg_class[sidx(cur, g2)] = th.dg_class[sidx(g2, cur)] = cls;
			}
			if (!th.wc[cur])
				for (int k = bA; k <= bI; ++k)
					for (int l = bA; l <= bI; ++l)
						if (!th.wc[pair_of(k, l)]) {
							th.dg[sidx(cur, pair_of(k, l))] = pen_loop;
							th.dg_class[sidx(cur, pair_of(k, l))] = DG_LOOP;
						}
		}
	for (int x = bA; x <= bI; ++x)
		for (int y = bA; y <= bI; ++y) {
			th.dg[sidx(pair_of(x, bGAP), pair_of(y, bGAP))] = pen_bulge;
			th.dg[sidx(pair_of(bGAP, x), pair_of(bGAP, y))] = pen_bulge;
			th.dg_class[sidx(pair_of(x, bGAP), pair_of(y, bGAP))] = DG_BULGE;
			th.dg_class[sidx(pair_of(bGAP, x), pair_of(bGAP, y))] = DG_BULGE;
		}
}

// update_dp_param at another temperature from the rule table (what the kernels do entry by entry in
// the Dinkelbach mode); used by the host-side self check only
void dg_at_temperature(const Thermo &th, float T, int32_t *out)
{
	const int32_t pen[7] = {0,
		unfavourable(th.supp[0] - T*(th.supp[1] + th.supp_sc[0])),
		unfavourable(th.supp[2] - T*(th.supp[3] + th.supp_sc[1])),
		unfavourable(th.supp[4] - T*(th.supp[5] + th.supp_sc[2])),
		unfavourable(th.supp[6] - T*(th.supp[7] + th.supp_sc[2])),
		unfavourable(th.supp[8] - T*(th.supp[9] + th.supp_sc[2])),
		unfavourable(th.supp[10] - T*(th.supp[11] + th.supp_sc[3]))};
	for (int i = 0; i < TABLE; ++i)
		out[i] = th.dg_class[i] ? pen[th.dg_class[i]] : scaled(th.H[i] - T*(th.S[i] + th.salt_correction));
}

// Per-row penalty tables of the fast alignment kernel.  Row i (1-based) of the DP matrix pairs
// the reversed oligo base qb = q[L-i] (previous: pq = q[L-i+1], GAP for i == 1) with the target;
// every delta_g lookup of align_dimer (nuc_cruc.cpp:529-646) is a function of (row, target
// dinucleotide) only.  Target dinucleotide index td = 4*pt + tb with pt in {A,C,G,T,GAP(=4)}.
void build_row_tables(const Thermo &th, const OligoStrand &os, int32_t *out)
{
	auto bbp = [&](int x, int y) { return (int)th.bbp[x*NB + y]; };
	const int L = os.len;
	for (int i = 1; i <= L; ++i) {
		int32_t *row = out + (size_t)(i - 1)*72;
		const int qb = os.seq[L - i];
		const int pq = (i == 1) ? (int)bGAP : (int)os.seq[L - i + 1];
		for (int td = 0; td < 20; ++td) {
			const int pt = (td/4 == 4) ? (int)bGAP : td/4;
			const int tb = td%4;
			const int cur = bbp(tb, qb);
			row[0 + td] = th.dg[sidx(bbp(pt, pq), cur)];              // M from M
			row[20 + td] = th.dg[sidx(bbp(pt, bGAP), cur)];           // M from I_query
			row[40 + td] = th.dg[sidx(bbp(pt, qb), bbp(tb, bGAP))];   // I_query from M
		}
		for (int tb = 0; tb < 4; ++tb) {
			row[60 + tb] = th.dg[sidx(bbp(bGAP, pq), bbp(tb, qb))];   // M from I_target
			row[64 + tb] = th.dg[sidx(bbp(tb, pq), bbp(bGAP, qb))];   // I_target from M
		}
		row[68] = th.dg[sidx(bbp(bGAP, pq), bbp(bGAP, qb))];          // I_target from I_target
		row[69] = row[70] = row[71] = 0;
	}
}

// Lean-tier rows (align_core.cuh, LEAN_*), derived from the rows of build_row_tables:
//   {P1,P4}[20] interleaved | V[4] = M-from-I_query for a real previous target base | column-1
//   M-from-I_query[4] | M-from-I_target[4] | I_target-from-M[4] | I_target-from-I_target
// All values are multiplied by 64 (LEAN_SCALE).
// Returns false when the table lacks the structure the lean fill relies on (see align_core.cuh);
// such an oligo strand is aligned by the full-trace tier instead.
bool build_lean_tables(const Thermo &th, const OligoStrand &os, const int32_t *rows, int32_t *out)
{
	const int len = os.len;
	bool ok = len >= 2;
	for (int r = 0; r < len; ++r) {
		const int32_t *row = rows + (size_t)r*72;
		int32_t *o = out + (size_t)r*64;
		for (int k = 0; k < 64; ++k) o[k] = 0;
		for (int td = 0; td < 20; ++td) { o[2*td] = row[0 + td]; o[2*td + 1] = row[40 + td]; }
		for (int tb = 0; tb < 4; ++tb) {
			o[40 + tb] = row[20 + tb];
			o[44 + tb] = row[20 + 16 + tb];
			o[48 + tb] = row[60 + tb];
			o[52 + tb] = row[64 + tb];
			// M from I_query must not depend on the previous target base
			for (int pt = 1; pt < 4; ++pt) ok = ok && row[20 + pt*4 + tb] == row[20 + tb];
			if (r >= 1) {
				ok = ok && row[60 + tb] == row[20 + tb];                     // M from I_target == V(r, tb)
				ok = ok && row[64 + tb] == rows[(size_t)(r - 1)*72 + 20 + tb]; // I_target from M == V(r-1, tb)
			}
		}
		o[56] = row[68];
		if (r >= 2) ok = ok && row[68] == rows[72 + 68];
		// scores are kept in units of 1/64 (the low six bits of a score carry the row index)
		for (int k = 0; k < 57; ++k) {
			ok = ok && o[k] > -(1 << 24) && o[k] < (1 << 24);
			o[k] = (int32_t)((int64_t)o[k]*64);
		}
		// evaluation side: pair code 7*query + target (best_base_pair) and the Watson-Crick flag of
		// this row's oligo base against each target base, one byte each
		const int qb = os.seq[len - 1 - r];
		uint32_t pairs = 0;
		for (int tb = 0; tb < 4; ++tb) {
			const unsigned code = th.bbp[qb*NB + tb];
			if (code >= 64) ok = false;
			pairs |= ((code & 63u) | (th.wc[code] ? 0x80u : 0u)) << (8*tb);
		}
		o[57] = (int32_t)pairs;
	}
	return ok;
}

// I_query from I_query: the only penalty that does not depend on the oligo row
void build_p5_table(const Thermo &th, int32_t *out)
{
	for (int td = 0; td < 20; ++td) {
		const int pt = (td/4 == 4) ? (int)bGAP : td/4;
		const int tb = td%4;
		out[td] = th.dg[sidx(th.bbp[pt*NB + bGAP], th.bbp[tb*NB + bGAP])];
	}
}

// Tm >= theta  <=>  dH/(R ln Ct + dS_total) >= K  (K = theta + 273.15, denominator negative)
//              <=>  G_K := dH - K*dS_total <= K * R ln Ct,
// and every other outcome of the evaluation (dH >= 0, non-negative denominator) reports Tm = 0.
// G_K of a gapless alignment is a sum over its columns (lean_evaluate, align_core.cuh):
//   initiation; A.T closing at both ends; per column outside an internal loop the stack term
//   H - K*S of (previous pair, pair) and one salt increment; per closed internal loop of m > 1
//   mismatches -K*loop_S[2m] and one salt increment (the subtract / add pairs cancel).
// For one oligo, the minimum over all start positions and all target strings of n columns is a
// shortest path over (oligo position, target base of the last column, open mismatch run).  The
// bound is taken 0.05 K below min_tm: binary32 evaluation of a few dozen terms differs from exact
// arithmetic by ~1e-3 K.  Alignments of n >= the returned value are evaluated normally.
int lean_min_columns(const Thermo &th, const OligoStrand &os, float min_tm)
{
	if (!(min_tm > 0.1f)) return 0; // Tm = 0 outcomes would pass: evaluate everything
	const int L = os.len;
	const int NMAX = std::min(L, 24);
	const double K = (double)min_tm - 0.05 + 273.15;
	const double limit = K*(double)os.r_log_ct;
	const double c_s = -K*(double)th.salt*(double)th.log_na; // salt term per unit of 0.5*num_base
	const double at_g = (double)th.at_H - K*(double)th.at_S;
	const double INF = 1e300;
	auto code_of = [&](int x, int t) { return (int)th.bbp[os.seq[x]*NB + t]; };
	auto is_at = [&](int c) { return c == 7*bA + bT || c == 7*bT + bA; };
	const int M = NMAX + 1;
	// f[(x*4 + t)*M + m]: cheapest G_K of an alignment whose last column is (oligo base x, target
	// base t) with an open run of m mismatch columns
	std::vector<double> f((size_t)L*4*M, INF), g((size_t)L*4*M, INF);
	for (int x = 0; x < L; ++x)
		for (int t = 0; t < 4; ++t) {
			const int c = code_of(x, t);
			if (th.wc[c]) f[(size_t)(x*4 + t)*M] = (double)th.init_H - K*(double)th.init_S + (is_at(c) ? at_g : 0.0);
		}
	for (int n = 2; n <= NMAX; ++n) {
		std::fill(g.begin(), g.end(), INF);
		for (int x = 0; x + 1 < L; ++x)
			for (int t1 = 0; t1 < 4; ++t1) {
				const int last = code_of(x, t1);
				for (int m = 0; m < n; ++m) {
					const double base = f[(size_t)(x*4 + t1)*M + m];
					if (base >= INF) continue;
					for (int t2 = 0; t2 < 4; ++t2) {
						const int cur = code_of(x + 1, t2);
						double v = base;
						if (th.wc[last] || th.wc[cur]) v += (double)th.H[last*NPAIR + cur] - K*(double)th.S[last*NPAIR + cur] + c_s;
						int m2;
						if (th.wc[cur]) {
							if (m > 1) v += -K*(double)th.loop_S[2*m] + c_s;
							m2 = 0;
						}
						else m2 = m + 1;
						double &slot = g[(size_t)((x + 1)*4 + t2)*M + m2];
						if (v < slot) slot = v;
					}
				}
			}
		f.swap(g);
		if (n >= 3) {
			double best = INF;
			for (int x = 0; x < L; ++x)
				for (int t = 0; t < 4; ++t) {
					const int c = code_of(x, t);
					if (th.wc[c]) best = std::min(best, f[(size_t)(x*4 + t)*M] + (is_at(c) ? at_g : 0.0));
				}
			if (best <= limit) return n;
		}
	}
	return NMAX + 1;
}

float symmetry_S() { return SL_SYMMETRY_S; }

float r_log_ct(float ct)
{
	return 1.9872e-3f*std::log(ct*1.0f); // NC_R*log(strand*alpha), nuc_cruc.cpp:2291
}

int build_words(const char *oligo, int W, bool complement, uint16_t *words)
{
	// DNAHash_iterator::build_word_list<std::string> (seq_hash.h:287-374): letters other than
	// ACGT break the run without shifting the accumulator; valid words are appended densely.
	const int L = (int)std::strlen(oligo);
	if (W > L) return 0;
	const unsigned mask = (1u << (2*W)) - 1u;
	unsigned acc = 0;
	int run = 0, n = 0;
	for (int s = 0; s < L; ++s) {
		const char c = complement ? oligo[L - 1 - s] : oligo[s];
		int code;
		switch (c) {
		case 'A': case 'a': code = 0; break;
		case 'C': case 'c': code = 1; break;
		case 'G': case 'g': code = 2; break;
		case 'T': case 't': code = 3; break;
		default: code = -1; break;
		}
		if (code < 0) run = 0;
		else {
			++run;
			acc = ((acc << 2) | (unsigned)(complement ? 3 - code : code)) & 0xffffu;
		}
		if (code >= 0 && run >= W) words[n++] = (uint16_t)(acc & mask);
	}
	return n;
}

} // namespace tnt
