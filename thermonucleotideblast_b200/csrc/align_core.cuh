// NucCruc heterodimer evaluation of one (oligo, target window) pair, one CUDA thread per pair.
//
// What the reference does per candidate in NucCruc::approximate_tm_heterodimer
// (nuc_cruc.cpp:2441-2454): the 3-state integer local alignment (align_dimer, :492-696), the
// enumeration of co-optimal tracebacks (tm_dimer :2517-2540, enumerate_dimer_alignments
// :973-1170, trace_back :1409-1618) and the nearest-neighbour dH/dS/Tm evaluation of every
// enumerated alignment (evaluate_alignment :1620-2299), keeping the lowest dG.
//
// Layout choices (not the reference's):
//  * the three DP rows live in shared memory, clamped at zero, indexed [column][thread] so a warp
//    touches 32 consecutive banks;
//  * the 15-byte NC_Elem of the reference shrinks to a 12-bit trace word per cell (trace masks +
//    the sign facts the traceback tests), written [cell][thread] to a per-CTA scratch in global
//    memory (coalesced 64 B per warp and cell);
//  * maximal cells are not collected in a list: a "candidate" bit marks cells that tied or raised
//    the running maximum, and everything from the last strict raise onwards is the row-major
//    list the reference builds;
//  * all floating point uses explicit round-to-nearest intrinsics in the reference's statement
//    order (no FMA contraction), so dH/dS/Tm are bit-identical to the host computation.
#pragma once

#include <cuda_runtime.h>
#include "tnt_types.h"

namespace tnt {

// trace word layout
constexpr unsigned TW_IQ_INS = 1u << 3;  // I_query came from M (insert)
constexpr unsigned TW_IQ_EXT = 1u << 4;  // I_query came from I_query (extend)
constexpr unsigned TW_IT_INS = 1u << 5;
constexpr unsigned TW_IT_EXT = 1u << 6;
constexpr unsigned TW_M_NEG = 1u << 7;
constexpr unsigned TW_M_ZERO = 1u << 8;
constexpr unsigned TW_IQ_NEG = 1u << 9;
constexpr unsigned TW_IT_NEG = 1u << 10;
constexpr unsigned TW_CAND = 1u << 11;   // M >= running maximum when the cell was computed
// word of a never-written cell (row 0 / column 0 of the reference matrix, nuc_cruc.h:531-536)
constexpr unsigned TW_BORDER = TW_M_NEG | TW_IQ_NEG | TW_IT_NEG;

constexpr int P_AT = bA*7 + bT, P_TA = bT*7 + bA, P_GT = bG*7 + bT, P_TG = bT*7 + bG;
constexpr int P_EE = bE*7 + bE, P_NONE = bGAP*7 + bGAP;

// IUPAC membership sets, bit0=A bit1=C bit2=G bit3=T (nuc_cruc_anchor.cpp:8-139)
__device__ __constant__ uint8_t c_base_set[NB] = {1, 2, 4, 8, 15, 0, 0, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15};

__device__ __forceinline__ bool is_virtual(int b) { return b == bE || b == bGAP; }

__device__ __forceinline__ bool dev_complementary(int q, int t)
{
	const unsigned ts = c_base_set[t];
	const unsigned tc = ((ts & 1u) << 3) | ((ts & 8u) >> 3) | ((ts & 2u) << 1) | ((ts & 4u) >> 1);
	return (c_base_set[q] & tc) != 0;
}

// target base j (0-based, NucCruc orientation) from the 2-bit packed window
__device__ __forceinline__ int packed_base(uint64_t lo, uint64_t hi, int j)
{
	return (int)(((j < 32) ? (lo >> (2*j)) : (hi >> (2*(j - 32)))) & 3u);
}

// ACGT-only target window (<= 64 bases) held in two registers; indexes like the byte array the
// generic kernel uses (NucCruc codes A..T == 0..3)
struct PackedTgt {
	uint64_t lo, hi;
	__device__ __forceinline__ int operator[](int j) const { return packed_base(lo, hi, j); }
};

struct AlnState {
	uint8_t q[MAX_COLS];
	uint8_t t[MAX_COLS];
	int b, e;            // live columns are [b, e)
	int fm_q, fm_t, lm_q, lm_t;
	float dH, dS, tm;
};

// Block-wide constants staged in shared memory.
struct DpShared {
	const int32_t *dg;   // [TABLE]
	const uint8_t *bbp;  // [NB*NB]
	const uint8_t *wc;   // [NPAIR]
	const uint8_t *q;    // oligo, 5'->3'
	int Lq;
	float T;             // target_T: ranks co-optimal alignments by dH - T*dS (the engine's temperature, or the
	                     // temperature of the current Dinkelbach iteration)
};

// delta_g table as the fill reads it: a resident table ...
struct DgTable {
	const int32_t *__restrict__ p;
	__device__ __forceinline__ int operator[](int i) const { return p[i]; }
};

// ... or update_dp_param (nuc_cruc.cpp:340-487) evaluated entry by entry at a temperature of the
// thread's own: the Dinkelbach iteration (nuc_cruc.cpp:2399-2440) re-derives the table at the Tm of
// the previous pass.  Same binary32 operations as the host (thermo.cpp; no FMA contraction).
struct DgAtT {
	const float *__restrict__ H;
	const float *__restrict__ S;
	const uint8_t *__restrict__ cls;
	float T, sc;
	int32_t pen[7];
	static __device__ __forceinline__ int32_t scaled(float x) { return __float2int_rz(__fmul_rn(x, 10000.0f)); }
	__device__ void set(const Thermo *__restrict__ th, float temperature)
	{
		H = th->H; S = th->S; cls = th->dg_class;
		T = temperature;
		sc = th->salt_correction;
		pen[0] = 0;
#define TNT_PEN(k, h, s, c) do { const int32_t v = scaled(__fsub_rn(th->supp[h], __fmul_rn(T, __fadd_rn(th->supp[s], th->supp_sc[c])))); pen[k] = v > 0 ? v : 0; } while (0)
		TNT_PEN(DG_LOOP, 0, 1, 0);
		TNT_PEN(DG_BULGE, 2, 3, 1);
		TNT_PEN(DG_TERM_AT, 4, 5, 2);
		TNT_PEN(DG_TERM_GC, 6, 7, 2);
		TNT_PEN(DG_TERM_INO, 8, 9, 2);
		TNT_PEN(DG_TERM_MM, 10, 11, 3);
#undef TNT_PEN
	}
	__device__ __forceinline__ int operator[](int i) const
	{
		const unsigned c = __ldg(cls + i);
		if (c) return pen[c];
		return scaled(__fsub_rn(__ldg(H + i), __fmul_rn(T, __fadd_rn(__ldg(S + i), sc))));
	}
};

struct DpResult {
	int runmax;
	int last_raise;  // linear cell index (i-1)*Lt + (j-1) of the last strict raise of the maximum
	int nmax;        // cells equal to the maximum from there on
};

// ------------------------------------------------------------------------------------------
// Stage 1: fill.  Rows follow the reversed oligo, columns the target 5'->3'
// (align_dimer, nuc_cruc.cpp:508-693).  NT = threads per block (row/trace stride).
// ------------------------------------------------------------------------------------------
// `tri` > 0: align_hairpin (nuc_cruc.cpp:771-971) -- the same recurrence for the oligo against itself
// over the triangle i + j <= tri + 1 only (tri = the longest stem the steric limit allows); cells
// inside the triangle depend on cells inside it alone.  The caller clears the trace beforehand
// (cells outside the triangle are never written).
template <int NT, class DG>
__device__ __forceinline__ DpResult nc_fill(const DpShared &sh, const DG &dg, const uint8_t *tgt, int Lt,
	int32_t *rowM, int32_t *rowIq, int32_t *rowIt, uint16_t *trace, int tri = 0)
{
	const uint8_t *__restrict__ bbp = sh.bbp;
	const int Lq = sh.Lq;

	for (int j = 0; j <= Lt; ++j) { rowM[j*NT] = 0; rowIq[j*NT] = 0; rowIt[j*NT] = 0; }

	DpResult res;
	res.runmax = -1;
	res.last_raise = -1;
	res.nmax = 0;

	const int rows = tri > 0 ? tri : Lq;
	for (int i = 1; i <= rows; ++i) {
		int cell = (i - 1)*Lt;
		const int cols = tri > 0 ? tri - (i - 1) : Lt;
		const int qb = sh.q[Lq - i];
		const int pq = (i == 1) ? (int)bGAP : (int)sh.q[Lq - i + 1];
		const int gap_pq = bbp[bGAP*NB + pq]*NPAIR;   // previous pair (GAP, pq)
		const int cur_it = bbp[bGAP*NB + qb];          // current pair of the I_target state
		const int pen_ext_it = dg[gap_pq + cur_it];

		int dM = 0, dIq = 0, dIt = 0;   // clamped states of (i-1, j-1)
		int lM = 0, lIq = 0;            // clamped states of (i, j-1)
		int pt_pq = bbp[bGAP*NB + pq];  // (prev target, prev query)
		int pt_qb = bbp[bGAP*NB + qb];  // (prev target, query)
		int pt_gap = bbp[bGAP*NB + bGAP];

#pragma unroll 2
		for (int j = 1; j <= cols; ++j, ++cell) {
			const int tb = tgt[j - 1];
			const int cur = bbp[tb*NB + qb];
			const int tb_pq = bbp[tb*NB + pq];
			const int tb_gap = bbp[tb*NB + bGAP];

			const int uM = rowM[j*NT], uIq = rowIq[j*NT], uIt = rowIt[j*NT];

			// match / mismatch state
			const int d1 = dM - dg[pt_pq*NPAIR + cur];
			const int d2 = dIq - dg[pt_gap*NPAIR + cur];
			const int d3 = dIt - dg[gap_pq + cur];
			int M;
			unsigned w;
			if (d1 >= d2) {
				if (d1 >= d3) { M = d1; w = T_DIAG | (d1 == d2 ? T_LEFT : 0u) | (d1 == d3 ? T_UP : 0u); }
				else { M = d3; w = T_UP; }
			}
			else {
				if (d2 >= d3) { M = d2; w = T_LEFT | (d2 == d3 ? T_UP : 0u); }
				else { M = d3; w = T_UP; }
			}

			// gap in the query: extend along the row
			const int qi = lM - dg[pt_qb*NPAIR + tb_gap];
			const int qe = lIq - dg[pt_gap*NPAIR + tb_gap];
			int Iq;
			if (qi >= qe) { Iq = qi; w |= TW_IQ_INS | (qi == qe ? TW_IQ_EXT : 0u); }
			else { Iq = qe; w |= TW_IQ_EXT; }

			// gap in the target: extend down the column
			const int ti = uM - dg[tb_pq*NPAIR + cur_it];
			const int te = uIt - pen_ext_it;
			int It;
			if (ti >= te) { It = ti; w |= TW_IT_INS | (ti == te ? TW_IT_EXT : 0u); }
			else { It = te; w |= TW_IT_EXT; }

			if (M < 0) w |= TW_M_NEG;
			if (M == 0) w |= TW_M_ZERO;
			if (Iq < 0) w |= TW_IQ_NEG;
			if (It < 0) w |= TW_IT_NEG;

			if (M >= res.runmax) {
				w |= TW_CAND;
				if (M > res.runmax) { res.runmax = M; res.last_raise = cell; res.nmax = 1; }
				else ++res.nmax;
			}
			trace[(size_t)cell*NT] = (uint16_t)w;

			dM = uM; dIq = uIq; dIt = uIt;
			lM = max(M, 0);
			lIq = max(Iq, 0);
			rowM[j*NT] = lM;
			rowIq[j*NT] = lIq;
			rowIt[j*NT] = max(It, 0);
			pt_pq = tb_pq;
			pt_qb = cur;
			pt_gap = tb_gap;
		}
	}
	return res;
}

template <int NT>
__device__ __forceinline__ DpResult nc_fill(const DpShared &sh, const uint8_t *tgt, int Lt,
	int32_t *rowM, int32_t *rowIq, int32_t *rowIt, uint16_t *trace, int tri = 0)
{
	return nc_fill<NT, DgTable>(sh, DgTable{sh.dg}, tgt, Lt, rowM, rowIq, rowIt, trace, tri);
}

// Read access to the stored trace of one alignment.  get(i, j) returns the trace word of DP cell
// (i, j), 1 <= i <= Lq, 1 <= j <= Lt, in the layout of the TW_* constants above.
template <int NT>
struct RowMajorTrace {
	static constexpr bool kHasGapStates = true;
	const uint16_t *trace;
	int Lt;
	__device__ __forceinline__ unsigned get(int i, int j) const { return trace[(size_t)((i - 1)*Lt + (j - 1))*NT]; }
};

// ------------------------------------------------------------------------------------------
// Stage 2: traceback of one path (trace_back, nuc_cruc.cpp:1409-1618)
// ------------------------------------------------------------------------------------------
struct Branch {
	uint16_t id;     // (cell index + 1)*3 + state; identifies the trace byte the reference points at
	uint8_t mask;
	uint8_t cur;
};
constexpr int MAX_BRANCH = 3*(MAX_OLIGO + MAX_WINDOW);

__device__ __forceinline__ bool path_split(unsigned m) { return __popc(m & 7u) > 1; }

template <class TV, class TG>
__device__ void nc_trace_back(const DpShared &sh, const TG &tgt, int Lt, const TV &tv,
	int start_cell, Branch *stack, int &nstack, int &zero_count, AlnState &a, unsigned &flags)
{
	const int Lq = sh.Lq;
	int last_i = start_cell/Lt + 1, last_j = start_cell%Lt + 1;
	a.fm_q = Lq - last_i;
	a.fm_t = last_j - 1;

	int truncate_at_zero = 0;
	bool count_zeros = false;
	if (zero_count < 0) { zero_count = 0; count_zeros = true; }
	else truncate_at_zero = zero_count--;

	unsigned cur_id = 0;        // 0 == the static first_match byte of the reference
	unsigned cur_mask = T_DIAG;

	for (;;) {
		bool valid = true;
		unsigned local;
		if (path_split(cur_mask)) {
			int k = 0;
			for (; k < nstack; ++k) if (stack[k].id == cur_id) break;
			if (k == nstack) {
				if (nstack == MAX_BRANCH) { flags |= F_STACK; return; }
				stack[k].id = (uint16_t)cur_id;
				stack[k].mask = (uint8_t)cur_mask;
				stack[k].cur = (uint8_t)((cur_mask & T_DIAG) ? T_DIAG : ((cur_mask & T_UP) ? T_UP : T_LEFT));
				++nstack;
			}
			local = stack[k].cur;
		}
		else local = cur_mask;

		const bool inside = (last_i >= 1 && last_j >= 1);
		const int cell = (last_i - 1)*Lt + (last_j - 1);
		const unsigned w = inside ? tv.get(last_i, last_j) : TW_BORDER;

		if (local == T_DIAG) {
			if (last_i > Lq || last_j < 1) valid = false;
			else {
				if (w & TW_M_NEG) valid = false;
				else if (w & TW_M_ZERO) {
					if (count_zeros) ++zero_count;
					else if (--truncate_at_zero == 0) valid = false;
				}
				if (last_i < 1) { flags |= F_OOB; return; } // the reference reads query[len] here
				if (a.e < MAX_COLS) { a.q[a.e] = sh.q[Lq - last_i]; a.t[a.e] = (uint8_t)tgt[last_j - 1]; ++a.e; }
				else flags |= F_TRUNC;
				a.lm_q = Lq - last_i;
				a.lm_t = last_j - 1;
				cur_id = (unsigned)(cell + 1)*3u + 0u;
				cur_mask = inside ? (w & 7u) : T_INVALID;
				--last_i;
				--last_j;
			}
		}
		else if (!TV::kHasGapStates && (local == T_LEFT || local == T_UP)) { flags |= F_NEEDGENERIC; return; }
		else if (local == T_LEFT) { // a gap goes into the query
			if (last_j < 1) valid = false;
			else {
				if (w & TW_IQ_NEG) valid = false;
				if (a.e < MAX_COLS) { a.q[a.e] = bGAP; a.t[a.e] = (uint8_t)tgt[last_j - 1]; ++a.e; }
				else flags |= F_TRUNC;
				a.lm_q = Lq - last_i + 1;
				a.lm_t = last_j - 1;
				cur_id = (unsigned)(cell + 1)*3u + 1u;
				cur_mask = inside ? (((w & TW_IQ_INS) ? T_DIAG : 0u) | ((w & TW_IQ_EXT) ? T_LEFT : 0u)) : T_INVALID;
				--last_j;
			}
		}
		else if (local == T_UP) { // a gap goes into the target
			if (last_i > Lq) valid = false;
			else {
				if (w & TW_IT_NEG) valid = false;
				if (last_i < 1) { flags |= F_OOB; return; }
				if (a.e < MAX_COLS) { a.q[a.e] = sh.q[Lq - last_i]; a.t[a.e] = bGAP; ++a.e; }
				else flags |= F_TRUNC;
				a.lm_q = Lq - last_i;
				a.lm_t = last_j;
				cur_id = (unsigned)(cell + 1)*3u + 2u;
				cur_mask = inside ? (((w & TW_IT_INS) ? T_DIAG : 0u) | ((w & TW_IT_EXT) ? T_UP : 0u)) : T_INVALID;
				--last_i;
			}
		}
		else { flags |= F_OOB; return; } // "invalid_match in trace back"

		if (!valid) break;
		if (last_i < 0 || last_j < 0) { flags |= F_OOB; return; }
	}
}

// ------------------------------------------------------------------------------------------
// Stage 3: nearest-neighbour evaluation (evaluate_alignment, nuc_cruc.cpp:1620-2299)
// ------------------------------------------------------------------------------------------
#define TNT_ADD(a, b) __fadd_rn((a), (b))
#define TNT_SUB(a, b) __fsub_rn((a), (b))
#define TNT_MUL(a, b) __fmul_rn((a), (b))

__device__ __forceinline__ bool has_at_initiation(const DpShared &sh, const uint8_t *q, const uint8_t *t, int k)
{
	do { --k; } while (k != 0 && (q[k] == bGAP || t[k] == bGAP)); // nuc_cruc.cpp:2888-2905
	const int bp = sh.bbp[q[k]*NB + t[k]];
	return bp == P_AT || bp == P_TA;
}

// `hairpin`: HAIRPIN mode of evaluate_alignment (nuc_cruc.cpp:1627-1633, :2286-2289) -- no initiation
// term (the caller preloaded a.dH / a.dS with the loop terms), Tm = dH/dS without a strand concentration.
__device__ bool nc_evaluate(const DpShared &sh, const Thermo *__restrict__ th, float r_log_ct, AlnState &a, bool hairpin = false)
{
	const uint8_t *q = a.q + a.b, *t = a.t + a.b;
	const int n = a.e - a.b;
	const uint8_t *bbp = sh.bbp;
	const uint8_t *wc = sh.wc;
	const float *__restrict__ H = th->H;
	const float *__restrict__ S = th->S;

#define ADD_HS(idx) do { const int _x = (idx); dH = TNT_ADD(dH, __ldg(H + _x)); dS = TNT_ADD(dS, __ldg(S + _x)); } while (0)
#define SUB_HS(idx) do { const int _x = (idx); dH = TNT_SUB(dH, __ldg(H + _x)); dS = TNT_SUB(dS, __ldg(S + _x)); } while (0)
#define NONVIRT_PAIR(p) (((p)%7 < bE) && ((p)/7 < bE))

	int terminal = P_NONE, last_last = P_NONE, last = P_NONE;
	float dH = th->init_H, dS = TNT_ADD(th->init_S, 0.0f);
	if (hairpin) { dH = a.dH; dS = a.dS; }
	unsigned nqgap = 0, ntgap = 0, nmm = 0, num_base = 0;
	bool terminal_5 = false;

	int cur = bbp[q[0]*NB + t[0]];
	if (wc[cur]) {
		terminal_5 = true;
		if (cur == P_AT || cur == P_TA) { dH = TNT_ADD(dH, th->at_H); dS = TNT_ADD(dS, th->at_S); }
	}
	num_base += is_virtual(q[0]) ? 0u : 1u;
	num_base += is_virtual(t[0]) ? 0u : 1u;

	for (int k = 1; k < n; ++k) {
		last_last = last;
		last = cur;
		cur = bbp[q[k]*NB + t[k]];
		const bool in_loop = (q[k] == bGAP) || (t[k] == bGAP) || (!wc[last] && !wc[cur]);

		if (!in_loop) {
			if (k == 1 && !wc[last] && NONVIRT_PAIR(last)) {
				ADD_HS((int)bbp[(last/7)*NB + bE]*NPAIR + cur);
				ADD_HS((int)bbp[bE*NB + (last%7)]*NPAIR + cur);
			}
			else if (k == n - 1 && !wc[cur] && NONVIRT_PAIR(cur)) {
				ADD_HS(last*NPAIR + (int)bbp[q[k]*NB + bE]);
				ADD_HS(last*NPAIR + (int)bbp[bE*NB + t[k]]);
			}
			else ADD_HS(last*NPAIR + cur);
			num_base += is_virtual(q[k]) ? 0u : 1u;
			num_base += is_virtual(t[k]) ? 0u : 1u;
		}

		if (wc[cur] || cur == P_EE) {
			terminal = cur;
			if (!terminal_5) {
				terminal_5 = true;
				if (cur == P_AT || cur == P_TA) { dH = TNT_ADD(dH, th->at_H); dS = TNT_ADD(dS, th->at_S); }
			}
			const unsigned max_gap = max(nqgap, ntgap);

			if (nmm > 1 || (max_gap > 0 && nmm == 1)) {
				// an internal loop closes on this column
				const unsigned gap_diff = nqgap > ntgap ? nqgap - ntgap : ntgap - nqgap;
				const unsigned loop_size = nmm*2 + gap_diff;
				if (loop_size == 2 && (last == P_GT || last == P_TG) && (last_last == P_GT || last_last == P_TG)) {
					ADD_HS(last_last*NPAIR + last);
					num_base += 2;
				}
				else {
					dS = TNT_ADD(dS, __ldg(th->loop_S + loop_size));
					dS = TNT_ADD(dS, TNT_MUL((float)gap_diff, th->asym_loop_dS));
					int rhs_q = k - 1, rhs_t = k - 1;
					SUB_HS(last*NPAIR + cur);
					const bool last_has_gap = (last%7 == bGAP) || (last/7 >= bGAP);
					if (!last_has_gap) ADD_HS(last*NPAIR + cur); // loop-terminal table == NN table
					else {
						int mm = P_NONE;
						if (last/7 == bGAP) {
							for (;;) {
								if (!is_virtual(q[rhs_q])) { mm = bbp[q[rhs_q]*NB + (last%7)]; break; }
								if (rhs_q == 0) break;
								--rhs_q;
							}
						}
						else {
							for (;;) {
								if (!is_virtual(t[rhs_t])) { mm = bbp[(last/7)*NB + t[rhs_t]]; break; }
								if (rhs_t == 0) break;
								--rhs_t;
							}
						}
						ADD_HS(mm*NPAIR + cur);
					}
					int lhs_q = k - 1, lhs_t = k - 1;
					for (;;) {
						const int pm = bbp[q[lhs_q]*NB + t[lhs_t]];
						if (wc[pm]) {
							++lhs_q; ++lhs_t;
							if (q[lhs_q] != bGAP && t[lhs_t] != bGAP) SUB_HS(pm*NPAIR + (int)bbp[q[lhs_q]*NB + t[lhs_t]]);
							else {
								num_base += 2;
								while (q[lhs_q] == bGAP) ++lhs_q;
								while (t[lhs_t] == bGAP) ++lhs_t;
							}
							ADD_HS(pm*NPAIR + (int)bbp[q[lhs_q]*NB + t[lhs_t]]);
							break;
						}
						if (lhs_q == 0) break;
						--lhs_q; --lhs_t;
					}
					if (rhs_q != lhs_q) ++num_base;
					if (rhs_t != lhs_t) ++num_base;
				}
			}
			else if (nqgap || ntgap) {
				const unsigned bulge = max(nqgap, ntgap);
				if (bulge == 1) ADD_HS(last_last*NPAIR + cur);
				dS = TNT_ADD(dS, __ldg(th->bulge_S + bulge));
				if (bulge != 1 && (q[k] == bA || q[k] == bT)) dS = TNT_ADD(dS, th->bulge_at_S);
				if (bulge != 1 && has_at_initiation(sh, q, t, k)) dS = TNT_ADD(dS, th->bulge_at_S);
			}
			nqgap = ntgap = nmm = 0;
		}
		else nmm += (!is_virtual(q[k]) && !is_virtual(t[k])) ? 1u : 0u;

		nqgap += (q[k] == bGAP) ? 1u : 0u;
		ntgap += (t[k] == bGAP) ? 1u : 0u;
	}

	if (terminal == P_AT || terminal == P_TA) { dH = TNT_ADD(dH, th->at_H); dS = TNT_ADD(dS, th->at_S); }

	a.dH = dH;
	a.dS = dS;
	if (dH >= 0.0f) return false;

	// dS += SALT*(0.5f*num_base - 1)*log[Na+]
	dS = TNT_ADD(dS, TNT_MUL(TNT_MUL(th->salt, TNT_SUB(TNT_MUL(0.5f, (float)num_base), 1.0f)), th->log_na));
	a.dS = dS;
	const float tm = hairpin ? TNT_SUB(__fdiv_rn(dH, dS), 273.15f) : TNT_SUB(__fdiv_rn(dH, TNT_ADD(r_log_ct, dS)), 273.15f);
	a.tm = fmaxf(0.0f, tm);
	return true;
#undef ADD_HS
#undef SUB_HS
#undef NONVIRT_PAIR
}

// ------------------------------------------------------------------------------------------
// Stage 2+3 driver: every maximal cell, every enumerated path
// (tm_dimer :2517-2540, enumerate_dimer_alignments :973-1170)
// ------------------------------------------------------------------------------------------
struct Best {
	bool valid;
	float dH, dS, tm;
};

// frayed ends: drop columns until both ends are Watson-Crick (nuc_cruc.cpp:1022-1054)
__device__ __forceinline__ void nc_trim_frayed_ends(const DpShared &sh, AlnState &a)
{
	while (a.e > a.b && !sh.wc[sh.bbp[a.q[a.e - 1]*NB + a.t[a.e - 1]]]) {
		if (!is_virtual(a.q[a.e - 1])) --a.lm_q;
		if (!is_virtual(a.t[a.e - 1])) ++a.lm_t;
		--a.e;
	}
	while (a.e > a.b && !sh.wc[sh.bbp[a.q[a.b]*NB + a.t[a.b]]]) {
		if (!is_virtual(a.q[a.b])) ++a.fm_q;
		if (!is_virtual(a.t[a.b])) --a.fm_t;
		++a.b;
	}
}

// optional dangling-end virtual bases (nuc_cruc.cpp:1088-1137); false: F_OOB raised
template <class TG>
__device__ __forceinline__ bool nc_dangling_ends(const DpShared &sh, const Thermo *__restrict__ th, const TG &tgt, int Lt,
	AlnState &a, unsigned &flags)
{
	const int Lq = sh.Lq;
	if (th->dangle5 && (a.fm_q != 0 || a.fm_t != Lt - 1)) {
		int qb, tb;
		if (a.fm_q == 0) qb = bE;
		else { --a.fm_q; if (a.fm_q < 0 || a.fm_q >= Lq) { flags |= F_OOB; return false; } qb = sh.q[a.fm_q]; }
		if (a.fm_t == Lt - 1) tb = bE;
		else { ++a.fm_t; if (a.fm_t < 0 || a.fm_t >= Lt) { flags |= F_OOB; return false; } tb = tgt[a.fm_t]; }
		--a.b;
		a.q[a.b] = (uint8_t)qb;
		a.t[a.b] = (uint8_t)tb;
	}
	if (th->dangle3 && (a.lm_q != Lq - 1 || a.lm_t != 0)) {
		int qb, tb;
		if (a.lm_q == Lq - 1) qb = bE;
		else { ++a.lm_q; if (a.lm_q < 0 || a.lm_q >= Lq) { flags |= F_OOB; return false; } qb = sh.q[a.lm_q]; }
		if (a.lm_t == 0) tb = bE;
		else { --a.lm_t; if (a.lm_t < 0 || a.lm_t >= Lt) { flags |= F_OOB; return false; } tb = tgt[a.lm_t]; }
		if (a.e < MAX_COLS) { a.q[a.e] = (uint8_t)qb; a.t[a.e] = (uint8_t)tb; ++a.e; }
		else flags |= F_TRUNC;
	}
	return true;
}

// `cells` lists the maximal DP cells as linear indices (i-1)*Lt + (j-1) in row-major order,
// i.e. the order of the reference's max_ptr vector.
// `fresh` false: continue with the best alignment found so far (the maximal cells of one window
// arrive in several chunks when there are more than MAX_MAXCELLS of them).
template <class TV, class TG>
__device__ void nc_enumerate(const DpShared &sh, const Thermo *__restrict__ th, float r_log_ct,
	const TG &tgt, int Lt, const TV &tv, const uint16_t *cells, int ncells,
	AlnState &work, AlnState &best_aln, Best &best, unsigned &flags, bool fresh = true)
{
	const float T = sh.T;
	if (fresh) {
		best.valid = false;
		best.dH = best.dS = best.tm = 0.0f;
	}
	if (ncells == 0) return;

	Branch stack[MAX_BRANCH];

	for (int ci = 0; ci < ncells; ++ci) {
		const int cell = cells[ci];

		bool first_time = true;
		int nstack = 0, zero_count = -1;
		unsigned trace_count = 0;
		float best_dg = TNT_SUB(best.dH, TNT_MUL(T, best.dS));

		for (;;) {
			if (!first_time && nstack == 0 && zero_count <= 0) break;
			if (16u < trace_count) break; // max_dp_path_enum = 16 (nuc_cruc.cpp:332, :1005)
			++trace_count;
			first_time = false;

			AlnState &a = work;
			a.b = a.e = 2; // room for a dangling-end column in front
			a.fm_q = a.fm_t = a.lm_q = a.lm_t = 0;
			a.dH = a.dS = a.tm = 0.0f;
			nc_trace_back(sh, tgt, Lt, tv, cell, stack, nstack, zero_count, a, flags);
			if (flags & (F_OOB | F_STACK | F_NEEDGENERIC)) return;

			nc_trim_frayed_ends(sh, a);

			if (zero_count == 0 && nstack > 0) {
				while (nstack > 0) {
					Branch &br = stack[nstack - 1];
					bool more = false;
					while ((br.cur = (uint8_t)(br.cur << 1)) < T_INVALID) if (br.cur & br.mask) { more = true; break; }
					if (more) break;
					--nstack;
				}
				zero_count = -1;
			}

			if (!nc_dangling_ends(sh, th, tgt, Lt, a, flags)) return;

			if (a.e - a.b < 3) continue;

			if (nc_evaluate(sh, th, r_log_ct, a)) {
				const float local_dg = TNT_SUB(a.dH, TNT_MUL(T, a.dS));
				if (!best.valid || local_dg < best_dg) {
					best.valid = true;
					best.dH = a.dH;
					best.dS = a.dS;
					best.tm = a.tm;
					best_dg = local_dg;
					// keep the winning columns
					best_aln.b = a.b; best_aln.e = a.e;
					best_aln.fm_q = a.fm_q; best_aln.fm_t = a.fm_t;
					best_aln.lm_q = a.lm_q; best_aln.lm_t = a.lm_t;
					for (int c = a.b; c < a.e; ++c) { best_aln.q[c] = a.q[c]; best_aln.t[c] = a.t[c]; }
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// Hairpins (approximate_tm_hairpin, nuc_cruc.cpp:2542-2618): the oligo against itself
// ------------------------------------------------------------------------------------------
// find_loop_index (nuc_cruc.cpp:2620-2860): the loop with its closing pair in the table of special loops
__device__ inline int nc_find_loop(const DpShared &sh, const Thermo *__restrict__ th, int start, int len)
{
	char text[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (int k = 0; k < len; ++k) {
		const int b = (start + k >= 0 && start + k < sh.Lq) ? sh.q[start + k] : (int)bGAP;
		text[k] = b <= bT ? "ACGT"[b] : (b <= bE ? 'E' : '?');
	}
	for (int i = 0; i < NUM_HAIRPIN_LOOP; ++i) {
		bool same = true;
		for (int k = 0; k < 7 && same; ++k) same = th->hairpin_loop[i][k] == text[k];
		if (same) return i;
	}
	return -1;
}

// evaluate_hairpin_alignment (nuc_cruc.cpp:2301-2394)
__device__ inline bool nc_evaluate_hairpin(const DpShared &sh, const Thermo *__restrict__ th, AlnState &a, unsigned &flags)
{
	const int last_3 = a.fm_q, last_5 = a.fm_t;
	const int loop_len = last_3 - last_5 - 1;
	if (loop_len < 0 || loop_len > MAX_LOOP || last_5 < 0 || last_3 >= sh.Lq) { flags |= F_OOB; return false; }
	a.dH = 0.0f;
	a.dS = TNT_ADD(0.0f, th->hairpin_S[loop_len]);
	const int last_pair = sh.bbp[sh.q[last_5]*NB + sh.q[last_3]];
	if (loop_len == 3) {
		const int idx = nc_find_loop(sh, th, last_5, 5);
		if (idx >= 0) { a.dH = TNT_ADD(a.dH, th->hairpin_special_H[idx]); a.dS = TNT_ADD(a.dS, th->hairpin_special_S[idx]); }
		if (last_pair == P_AT || last_pair == P_TA) a.dS = TNT_ADD(a.dS, th->bulge_at_S);
	}
	else {
		if (loop_len == 4) {
			const int idx = nc_find_loop(sh, th, last_5, 6);
			if (idx >= 0) { a.dH = TNT_ADD(a.dH, th->hairpin_special_H[idx]); a.dS = TNT_ADD(a.dS, th->hairpin_special_S[idx]); }
		}
		if (last_5 + 1 >= sh.Lq || last_3 - 1 < 0) { flags |= F_OOB; return false; }
		const int cur = sh.bbp[sh.q[last_5 + 1]*NB + sh.q[last_3 - 1]];
		// param_hairpin_terminal_* are copies of the stacking tables (nuc_cruc_santa_lucia.cpp:594-595)
		a.dH = TNT_ADD(a.dH, __ldg(th->H + last_pair*NPAIR + cur));
		a.dS = TNT_ADD(a.dS, __ldg(th->S + last_pair*NPAIR + cur));
	}
	return nc_evaluate(sh, th, 0.0f, a, true);
}

__device__ __forceinline__ void nc_hairpin_keep(const DpShared &sh, const AlnState &a, AlnState &best_aln, Best &best, float &best_dg)
{
	const float local_dg = TNT_SUB(a.dH, TNT_MUL(sh.T, a.dS));
	if (!best.valid || local_dg < best_dg) {
		best.valid = true;
		best.dH = a.dH; best.dS = a.dS; best.tm = a.tm;
		best_dg = local_dg;
		best_aln.b = a.b; best_aln.e = a.e;
		best_aln.fm_q = a.fm_q; best_aln.fm_t = a.fm_t;
		best_aln.lm_q = a.lm_q; best_aln.lm_t = a.lm_t;
	}
}

// enumerate_hairpin_alignments (nuc_cruc.cpp:1172-1407) over the maximal cells of a chunk
template <class TV>
__device__ void nc_enumerate_hairpin(const DpShared &sh, const Thermo *__restrict__ th, const uint8_t *tgt, int Lt, const TV &tv,
	const uint16_t *cells, int ncells, AlnState &work, AlnState &best_aln, Best &best, unsigned &flags, bool fresh)
{
	const int Lq = sh.Lq;
	if (fresh) {
		best.valid = false;
		best.dH = best.dS = best.tm = 0.0f;
	}
	Branch stack[MAX_BRANCH];
	for (int ci = 0; ci < ncells; ++ci) {
		const int cell = cells[ci];
		bool first_time = true;
		int nstack = 0, zero_count = -1;
		unsigned trace_count = 0;
		float best_dg = TNT_SUB(best.dH, TNT_MUL(sh.T, best.dS));
		for (;;) {
			if (!first_time && nstack == 0 && zero_count <= 0) break;
			if (16u < trace_count) break;
			++trace_count;
			first_time = false;
			AlnState &a = work;
			a.b = a.e = 2;
			a.fm_q = a.fm_t = a.lm_q = a.lm_t = 0;
			a.dH = a.dS = a.tm = 0.0f;
			nc_trace_back(sh, tgt, Lt, tv, cell, stack, nstack, zero_count, a, flags);
			if (flags & (F_OOB | F_STACK)) return;
			nc_trim_frayed_ends(sh, a);
			if (zero_count == 0 && nstack > 0) {
				while (nstack > 0) {
					Branch &br = stack[nstack - 1];
					bool more = false;
					while ((br.cur = (uint8_t)(br.cur << 1)) < T_INVALID) if (br.cur & br.mask) { more = true; break; }
					if (more) break;
					--nstack;
				}
				zero_count = -1;
			}
			// the stem as it is (:1265-1286)
			if (a.e - a.b >= 3 && nc_evaluate_hairpin(sh, th, a, flags)) nc_hairpin_keep(sh, a, best_aln, best, best_dg);
			if (flags & F_OOB) return;
			// one more column at the open end: the next bases or a dangling-end virtual base (:1307-1326)
			if (a.lm_t != 0 || a.lm_q != Lq - 1) {
				int tb, qb;
				if (a.lm_t == 0) tb = bE;
				else { --a.lm_t; if (a.lm_t < 0 || a.lm_t >= Lq) { flags |= F_OOB; return; } tb = sh.q[a.lm_t]; }
				if (a.lm_q == Lq - 1) qb = bE;
				else { ++a.lm_q; if (a.lm_q < 0 || a.lm_q >= Lq) { flags |= F_OOB; return; } qb = sh.q[a.lm_q]; }
				if (a.e < MAX_COLS) { a.q[a.e] = (uint8_t)qb; a.t[a.e] = (uint8_t)tb; ++a.e; }
				else flags |= F_TRUNC;
			}
			const int align_size = a.e - a.b;
			if (align_size < 3) continue;
			if (nc_evaluate_hairpin(sh, th, a, flags)) nc_hairpin_keep(sh, a, best_aln, best, best_dg);
			if (flags & F_OOB) return;
			// without the closing pair, unless it is G-C / C-G (:1360-1406)
			if (align_size <= 3) continue;
			if (a.fm_t < 0 || a.fm_q >= Lq) { flags |= F_OOB; return; }
			const int last_pair = sh.bbp[sh.q[a.fm_t]*NB + sh.q[a.fm_q]];
			if (last_pair == bG*7 + bC || last_pair == bC*7 + bG) continue;
			++a.fm_q;
			--a.fm_t;
			++a.b;
			if (nc_evaluate_hairpin(sh, th, a, flags)) nc_hairpin_keep(sh, a, best_aln, best, best_dg);
			if (flags & F_OOB) return;
		}
	}
}

// ------------------------------------------------------------------------------------------
// Per-oligo filters' inputs (nuc_cruc_anchor.cpp:143-192, :249-298; nuc_cruc.h:389-483)
// ------------------------------------------------------------------------------------------
template <class TG>
__device__ inline unsigned nc_anchor5(const DpShared &sh, const TG &tgt, int Lt, const AlnState &a)
{
	unsigned anchor = 0;
	int qi = 0, ti = a.fm_q + a.fm_t;
	if (a.e > a.b && a.t[a.b] == bE) return 0;
	if (a.e > a.b && a.q[a.b] == bE) --ti;
	if (ti >= Lt) return 0;
	while (qi < sh.Lq && ti >= 0 && dev_complementary(sh.q[qi], tgt[ti])) { ++anchor; ++qi; --ti; }
	return anchor;
}

template <class TG>
__device__ inline unsigned nc_anchor3(const DpShared &sh, const TG &tgt, int Lt, const AlnState &a)
{
	unsigned anchor = 0;
	int qi = sh.Lq - 1, ti = (a.lm_q + a.lm_t + 1) - sh.Lq;
	if (a.e > a.b && a.t[a.e - 1] == bE) return 0;
	if (a.e > a.b && a.q[a.e - 1] == bE) ++ti;
	if (ti >= Lt || ti < 0) return 0;
	while (qi >= 0 && ti < Lt && dev_complementary(sh.q[qi], tgt[ti])) { ++anchor; --qi; ++ti; }
	return anchor;
}

__device__ inline void nc_counts(const DpShared &sh, const AlnState &a, unsigned &mm, unsigned &gaps, unsigned &poly)
{
	unsigned aligned = 0, run = 0;
	mm = gaps = poly = 0;
	for (int c = a.b; c < a.e; ++c) {
		const int q = a.q[c], t = a.t[c];
		if (!is_virtual(q)) {
			if (!is_virtual(t) && !dev_complementary(q, t)) ++mm;
			++aligned;
		}
		gaps += (q == bGAP) + (t == bGAP);
		if (t >= bM && t <= bN) { if (++run > poly) poly = run; }
		else run = 0;
	}
	mm += (unsigned)sh.Lq - aligned;
}

// ------------------------------------------------------------------------------------------
// Fast fill (ACGT-only windows): column sweep with the oligo rows held in registers.
//
// The same recurrence as nc_fill, reorganised for the SM:
//  * one column of the DP matrix (all LQ oligo rows, three states) lives in registers; the row
//    loop is fully unrolled, so there is no indexed scratch and no shared-memory traffic for
//    DP state at all;
//  * every penalty the reference looks up through best_base_pair + delta_g depends only on
//    (oligo row, target dinucleotide); the host tabulates them per oligo strand
//    (build_row_tables, thermo.cpp) and the kernel reads them from shared memory as
//    tab[row][kind][dinucleotide] -- the row is a compile-time offset, lanes differ only in the
//    dinucleotide (<= 20 consecutive words), hence conflict-free;
//  * a trace word is assembled with one funnel shift per fact (sign bits of differences) and two
//    rows share one 32-bit store.
// Rows >= Lq are padding: their penalties are huge, they never reach the running maximum and no
// real row depends on them.
// ------------------------------------------------------------------------------------------
constexpr int ROW_WORDS = 72;      // P1[20] P2[20] P4[20] P3[4] P6[4] P7 pad[3]
constexpr int ROW_P1 = 0, ROW_P2 = 20, ROW_P4 = 40, ROW_P3 = 60, ROW_P6 = 64, ROW_P7 = 68;
constexpr int32_t ROW_PAD_PENALTY = 1 << 28;
#ifndef TNT_MAX_MAXCELLS
#define TNT_MAX_MAXCELLS 64
#endif
constexpr int MAX_MAXCELLS = TNT_MAX_MAXCELLS;   // tied maximal cells handled per chunk (a test build uses 2)

// ------------------------------------------------------------------------------------------
// Lean tier of the fast fill: what nearly every window needs and nothing more.
//
// A traceback that never leaves the match state only needs, per cell, (a) "the diagonal
// predecessor alone gives the maximum" and (b) "M < 0"; the value of M along such a path can be
// rebuilt backwards from the maximal score, because there M(i,j) = max(M(i-1,j-1), 0) - P1(i,j).
// So the trace shrinks to two bits per cell (16 rows per 32-bit store).  Anything else -- a tie
// or a gap state on the optimal path, several cells tied for the maximum -- hands the candidate
// to the full-trace tier below.
//
// The gap penalties of the reference's table have structure (update_dp_param,
// nuc_cruc.cpp:340-487: a gap pair next to a real pair costs a "terminal" penalty that depends on
// the real pair only; gap next to gap costs the bulge constant).  For rows >= 2 and columns >= 2
//   M from I_query  ==  M from I_target  ==: V(row, tb)
//   I_target from M (row)  ==  V(row - 1, tb)
//   I_target from I_target  ==  one constant
// which the host *verifies* per oligo strand on the actual table (lean_tables_ok, thermo.cpp);
// an oligo for which any identity fails goes to the full-trace tier.  Row 1 and column 1 (GAP
// neighbours) use their own entries.  Per cell that leaves three 32-bit shared-memory words
// (P1 and P4 as one 64-bit load, V) instead of six.
// ------------------------------------------------------------------------------------------
constexpr int LEAN_WORDS = 64;     // {P1,P4}[20] V[4] P2c1[4] P3[4] P6[4] P7 PAIR pad[6]
constexpr int LEAN_T = 0, LEAN_V = 40, LEAN_P2C1 = 44, LEAN_P3 = 48, LEAN_P6 = 52, LEAN_P7 = 56, LEAN_PAIR = 57;

struct FastDp {
	unsigned runkey;   // score << 12 | (63 - row) << 6 | (63 - col) of the first maximal cell (row-major)
	unsigned lastkey;  // score << 12 | row << 6 | col of the last maximal cell
};

template <int LQ>
struct LeanGeom {
	static constexpr int kWordsPerCol = (LQ + 15)/16;
};

// Trace view of the lean fill (trace in shared memory, [word][thread]).
template <int LQ, int NT>
struct ColMajorLean {
	const uint32_t *trace32;
	const int32_t *tab;      // lean rows in shared memory
	PackedTgt tgt;
	int maxscore;
	// bit 1: the diagonal alone gives the maximum, bit 0: M < 0
	__device__ __forceinline__ unsigned cell_bits(int i, int j) const
	{
		const int g = (i - 1) >> 4, p = (i - 1) & 15;
		const int n = (LQ - 16*g) < 16 ? (LQ - 16*g) : 16;
		const uint32_t word = trace32[((j - 1)*LeanGeom<LQ>::kWordsPerCol + g)*NT];
		return (word >> (2*(n - 1 - p))) & 3u;
	}
};

// One column of the lean sweep.  All lean-table values (hence all scores) are multiples of 64
// (LEAN_SCALE): the low six bits of max(M,0) carry the row, so "column maximum + first / last row
// attaining it" costs one fused add-max each; the column results are folded into the running keys
//   key1 = score << 12 | (63 - row) << 6 | (63 - col)      key2 = score << 12 | row << 6 | col
// (first / last maximal cell in row-major order; they name the same cell iff the maximum is unique).
// The inputs of row r+1 that come from column j-1 (d1, max of the two gap states) are formed
// while row r still holds its old values, so the state arrays are updated in place.
constexpr int LEAN_SCALE = 64;

template <int LQ, int NT, bool FIRST>
__device__ __forceinline__ void lean_column(const int32_t *__restrict__ tab, int tb, int td, int p5, int p7c, unsigned jj,
	uint32_t *__restrict__ col, int (&cM)[LQ], int (&cIq)[LQ], int (&cIt)[LQ], unsigned &key1, unsigned &key2)
{
	unsigned ck1 = 0, ck2 = 0;
	int uM = 0, uIt = 0;          // (i-1, j)
	int prevV = 0;
	unsigned acc = 0;
	int2 tcur = *reinterpret_cast<const int2 *>(tab + LEAN_T + 2*td);
	int d1 = -tcur.x;             // nothing above the first row
	int gmax = 0;                 // max(I_query, I_target) of (i-1, j-1)
#pragma unroll
	for (int r = 0; r < LQ; ++r) {
		const int32_t *__restrict__ row = tab + r*LEAN_WORDS;
		const int oM = cM[r], oIq = cIq[r], oIt = cIt[r]; // (i, j-1)
		int m23, It;
		if (FIRST || r == 0) {
			// both diagonal gap states are zero here (column 0 / row 0)
			const int p2 = FIRST ? row[LEAN_P2C1 + tb] : row[LEAN_V + tb];
			m23 = -min(p2, row[LEAN_P3 + tb]);
			It = max(uM - row[LEAN_P6 + tb], uIt - row[LEAN_P7]);
			if (!FIRST) prevV = p2;
		}
		else {
			const int v = row[LEAN_V + tb];
			m23 = gmax - v;
			It = max(uM - prevV, uIt - p7c);
			prevV = v;
		}
		const int M = max(d1, m23);
		const int Iq = max(oM - tcur.y, oIq - p5);
		const int mM = max(M, 0);

		ck1 = max(ck1, (unsigned)mM + (unsigned)(63 - r));
		ck2 = max(ck2, (unsigned)mM + (unsigned)r);

		acc = __funnelshift_l((unsigned)(m23 - d1), acc, 1); // set <=> the diagonal alone is the maximum
		acc = __funnelshift_l((unsigned)M, acc, 1);
		if ((r & 15) == 15 || r == LQ - 1) col[(r >> 4)*NT] = acc;

		if (r + 1 < LQ) {
			tcur = *reinterpret_cast<const int2 *>(row + LEAN_WORDS + LEAN_T + 2*td);
			d1 = oM - tcur.x;
			gmax = max(oIq, oIt);
		}
		cM[r] = mM;
		cIq[r] = max(Iq, 0);
		cIt[r] = max(It, 0);
		uM = mM;
		uIt = cIt[r];
	}
	key1 = max(key1, ck1*64u + (63u - jj));
	key2 = max(key2, ck2*64u + jj);
}

#ifndef TNT_FILL_INLINE
#define TNT_FILL_INLINE __forceinline__
#endif
template <int LQ, int NT>
__device__ TNT_FILL_INLINE FastDp nc_fill_lean(const int32_t *__restrict__ tab, const int32_t *__restrict__ p5tab,
	uint64_t tlo, uint64_t thi, int Lt, uint32_t *__restrict__ trace32)
{
	static_assert(LQ >= 2 && LQ <= 64, "rows fit six key bits");
	constexpr int WPC = LeanGeom<LQ>::kWordsPerCol;
	int cM[LQ], cIq[LQ], cIt[LQ];
#pragma unroll
	for (int r = 0; r < LQ; ++r) { cM[r] = 0; cIq[r] = 0; cIt[r] = 0; }

	unsigned key1 = 0, key2 = 0;
	const int p7c = tab[LEAN_WORDS + LEAN_P7];
	int pt = packed_base(tlo, thi, 0);
	// column 1: GAP in front of it
	lean_column<LQ, NT, true>(tab, pt, 16 + pt, p5tab[16 + pt], p7c, 0u, trace32, cM, cIq, cIt, key1, key2);
	for (int j = 2; j <= Lt; ++j) {
		const int tb = packed_base(tlo, thi, j - 1);
		const int td = pt*4 + tb;
		lean_column<LQ, NT, false>(tab, tb, td, p5tab[td], p7c, (unsigned)(j - 1),
			trace32 + (size_t)(j - 1)*WPC*NT, cM, cIq, cIt, key1, key2);
		pt = tb;
	}
	FastDp res;
	res.runkey = key1;
	res.lastkey = key2;
	return res;
}

// The maximal cell of a lean fill.  Returns 1 (cells[0] set), -1 when no cell is positive (the
// reference's ">= -1" rule then decides: generic kernel) or -2 when several cells tie for the
// maximum (full-trace tier).
__device__ __forceinline__ int lean_max_cell(const FastDp &dp, int Lt, uint16_t *cells)
{
	if ((dp.runkey >> 12) == 0) return -1;
	const int r1 = 63 - (int)((dp.runkey >> 6) & 63u), j1 = 63 - (int)(dp.runkey & 63u);
	const int r2 = (int)((dp.lastkey >> 6) & 63u), j2 = (int)(dp.lastkey & 63u);
	if (r1 != r2 || j1 != j2) return -2;
	cells[0] = (uint16_t)(r1*Lt + j1);
	return 1;
}

// Column k of a gapless alignment that starts at (query fm_q, target fm_t): the oligo base
// fm_q + k against the target base fm_t - k.  Each lean row carries, per target base, the
// reference's pair code 7*query + target (best_base_pair) in bits 0..5 and "Watson-Crick" in bit 7.
__device__ __forceinline__ unsigned lean_pair(const int32_t *__restrict__ tab, int Lq, const PackedTgt &tgt, int fm_q, int fm_t, int k)
{
	const int tb = tgt[fm_t - k];
	return ((unsigned)tab[(Lq - 1 - (fm_q + k))*LEAN_WORDS + LEAN_PAIR] >> (8*tb)) & 0xffu;
}

// evaluate_alignment (nuc_cruc.cpp:1620-2299) restricted to what a trimmed gapless alignment can
// reach: no gap columns, no virtual bases, first and last column Watson-Crick.  Same float
// operations in the same order as nc_evaluate (including the subtract-then-add pairs of the
// internal-loop branch, which do not cancel in binary32).
__device__ __forceinline__ bool lean_evaluate(const int32_t *__restrict__ tab, int Lq, const Thermo *__restrict__ th, float r_log_ct,
	const PackedTgt &tgt, int fm_q, int fm_t, int n, float &out_dH, float &out_dS, float &out_tm)
{
	const float *__restrict__ H = th->H;
	const float *__restrict__ S = th->S;
#define ADD_HS(idx) do { const int _x = (idx); dH = TNT_ADD(dH, __ldg(H + _x)); dS = TNT_ADD(dS, __ldg(S + _x)); } while (0)
#define SUB_HS(idx) do { const int _x = (idx); dH = TNT_SUB(dH, __ldg(H + _x)); dS = TNT_SUB(dS, __ldg(S + _x)); } while (0)
	float dH = th->init_H, dS = TNT_ADD(th->init_S, 0.0f);
	unsigned c = lean_pair(tab, Lq, tgt, fm_q, fm_t, 0);
	int cur = (int)(c & 63u);
	bool cur_wc = (c & 0x80u) != 0; // true after trimming
	if (cur_wc && (cur == P_AT || cur == P_TA)) { dH = TNT_ADD(dH, th->at_H); dS = TNT_ADD(dS, th->at_S); }
	unsigned num_base = 2, nmm = 0;
	int terminal = P_NONE, last = P_NONE;
	int loop_pm = P_NONE, loop_mm = P_NONE; // the pair in front of the open mismatch run and its first mismatch

	for (int k = 1; k < n; ++k) {
		const bool last_wc = cur_wc;
		last = cur;
		c = lean_pair(tab, Lq, tgt, fm_q, fm_t, k);
		cur = (int)(c & 63u);
		cur_wc = (c & 0x80u) != 0;
		if (last_wc || cur_wc) { // not inside a loop
			ADD_HS(last*NPAIR + cur);
			num_base += 2;
		}
		if (cur_wc) {
			terminal = cur;
			if (nmm > 1) {
				// an internal loop of nmm mismatches closes on this column
				dS = TNT_ADD(dS, __ldg(th->loop_S + 2*nmm));
				dS = TNT_ADD(dS, TNT_MUL(0.0f, th->asym_loop_dS));
				SUB_HS(last*NPAIR + cur);
				ADD_HS(last*NPAIR + cur);
				SUB_HS(loop_pm*NPAIR + loop_mm);
				ADD_HS(loop_pm*NPAIR + loop_mm);
				num_base += 2;
			}
			nmm = 0;
		}
		else {
			if (nmm == 0) { loop_pm = last; loop_mm = cur; }
			++nmm;
		}
	}
	if (terminal == P_AT || terminal == P_TA) { dH = TNT_ADD(dH, th->at_H); dS = TNT_ADD(dS, th->at_S); }
	out_dH = dH;
	out_dS = dS;
	if (dH >= 0.0f) return false;
	dS = TNT_ADD(dS, TNT_MUL(TNT_MUL(th->salt, TNT_SUB(TNT_MUL(0.5f, (float)num_base), 1.0f)), th->log_na));
	out_dS = dS;
	const float tm = TNT_SUB(__fdiv_rn(dH, TNT_ADD(r_log_ct, dS)), 273.15f);
	out_tm = fmaxf(0.0f, tm);
	return true;
#undef ADD_HS
#undef SUB_HS
}

#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 6
__device__ uint32_t *tnt_dbg_why; // timing / statistics experiment only
#endif
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 7
__device__ uint32_t *tnt_dbg_skip; // [0] evaluations skipped, [1] skipped although Tm >= min_tm (must stay 0)
#endif
// The whole post-fill work of a lean candidate: one traceback down the diagonal of the single
// maximal cell, frayed ends, dangling ends, evaluation.  Handles the ordinary case only -- a run of
// cells with M > 0 whose maximum comes from the diagonal alone, closed by a cell with M < 0 inside
// the matrix (the first pair of the duplex: it carries a terminal penalty, no stack).  Anything
// else (a tie or gap on the path, M == 0 on the path, the matrix border) returns false: the
// full-trace tier then follows the reference's trace_back rules literally.
// The alignment is a diagonal segment, so nothing is materialised unless the record can be
// written (`keep_all`, or Tm inside [min_tm, max_tm], the first filter of finish_alignment).
template <int LQ, int NT>
__device__ __forceinline__ bool lean_finish(const DpShared &sh, const Thermo *__restrict__ th, float r_log_ct,
	const ColMajorLean<LQ, NT> &tv, int Lt, int cell, bool keep_all, float min_tm, float max_tm, int min_cols,
	AlnState &a, Best &best, unsigned &flags)
{
	const int Lq = sh.Lq;
	int i = cell/Lt + 1, j = cell%Lt + 1;
	int fm_q = Lq - i, fm_t = j - 1;
	int n = 0;
	int m = tv.maxscore;
	for (;;) {
		const unsigned bits = tv.cell_bits(i, j);
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 6
		if (m <= 0 && !(bits & 1u)) { atomicAdd(tnt_dbg_why + 2, 1u); return false; }
		++n;
		if (m <= 0) break;
		if (!(bits & 2u)) { atomicAdd(tnt_dbg_why + 1, 1u); return false; }
		if (i == 1 || j == 1) { atomicAdd(tnt_dbg_why + 3, 1u); return false; }
#else
		if (m <= 0 && !(bits & 1u)) return false;      // M == 0 on the path
		++n;
		if (m <= 0) break;                             // M < 0: the closing pair
		if (!(bits & 2u)) return false;                // tie or gap state
		if (i == 1 || j == 1) return false;            // border next
#endif
		const int tb = tv.tgt[j - 1], pt = tv.tgt[j - 2];
		m += tv.tab[(i - 1)*LEAN_WORDS + LEAN_T + 2*(pt*4 + tb)];
		--i;
		--j;
	}
	best.valid = false;
	a.b = a.e = 2;
	a.fm_q = a.fm_t = a.lm_q = a.lm_t = 0;
	a.dH = a.dS = a.tm = 0.0f;

	if (th->dangle5 || th->dangle3) {
		// dangling-end columns can hold virtual bases: materialise and use the general code
		a.fm_q = fm_q; a.fm_t = fm_t;
		for (int k = 0; k < n; ++k) { a.q[a.e] = sh.q[fm_q + k]; a.t[a.e] = (uint8_t)tv.tgt[fm_t - k]; ++a.e; }
		a.lm_q = fm_q + n - 1; a.lm_t = fm_t - (n - 1);
		nc_trim_frayed_ends(sh, a);
		if (!nc_dangling_ends(sh, th, tv.tgt, Lt, a, flags)) return true; // F_OOB is reported, not retried
		if (a.e - a.b >= 3 && nc_evaluate(sh, th, r_log_ct, a)) {
			best.valid = true;
			best.dH = a.dH; best.dS = a.dS; best.tm = a.tm;
		}
		else { a.b = a.e = 2; a.fm_q = a.fm_t = a.lm_q = a.lm_t = 0; }
		return true;
	}

	// frayed ends (nuc_cruc.cpp:1022-1054): back, then front
	while (n > 0 && !(lean_pair(tv.tab, Lq, tv.tgt, fm_q, fm_t, n - 1) & 0x80u)) --n;
	while (n > 0 && !(lean_pair(tv.tab, Lq, tv.tgt, fm_q, fm_t, 0) & 0x80u)) { ++fm_q; --fm_t; --n; }
	if (n < 3) return true;
	// too short to reach min_tm whatever the bases are (lean_min_columns, thermo.cpp): the Tm
	// filter would reject it, so it is not evaluated (unless every result is wanted)
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 7
	// verification build: evaluate what would have been skipped and count contradictions
	if (!keep_all && n < min_cols) {
		float vH, vS, vT = 0.0f;
		atomicAdd(tnt_dbg_skip, 1u);
		if (lean_evaluate(tv.tab, Lq, th, r_log_ct, tv.tgt, fm_q, fm_t, n, vH, vS, vT) && !(vT < min_tm)) atomicAdd(tnt_dbg_skip + 1, 1u);
		return true;
	}
#else
	if (!keep_all && n < min_cols) return true;
#endif
	float dH, dS, tm = 0.0f;
	if (!lean_evaluate(tv.tab, Lq, th, r_log_ct, tv.tgt, fm_q, fm_t, n, dH, dS, tm)) return true;
	best.valid = true;
	best.dH = dH; best.dS = dS; best.tm = tm;
	a.dH = dH; a.dS = dS; a.tm = tm;
	a.fm_q = fm_q; a.fm_t = fm_t;
	a.lm_q = fm_q + n - 1; a.lm_t = fm_t - (n - 1);
	a.e = 2 + n;
	if (keep_all || !(tm < min_tm || tm > max_tm))
		for (int k = 0; k < n; ++k) { a.q[2 + k] = sh.q[fm_q + k]; a.t[2 + k] = (uint8_t)tv.tgt[fm_t - k]; }
	return true;
}

// ------------------------------------------------------------------------------------------
// Full-trace variant of the fast fill: same sweep, 12-bit trace words that also record the gap
// states (two rows per 32-bit store).  Used for the few candidates whose optimal path enters a
// gap state, which the 8-bit trace above cannot follow.
// ------------------------------------------------------------------------------------------
struct FastDpFull {
	int runmax, last_i, last_j, nmax;
};

// packed trace bits of the fast fill, most significant first (order of insertion)
//   11 d1!=M  10 d2!=M  9 d3!=M  8 qi!=Iq  7 qe!=Iq  6 ti!=It  5 te!=It  4 M<0  3 M>0  2 Iq<0  1 It<0  0 M<runmax
__device__ __forceinline__ unsigned decode_full_trace(unsigned raw)
{
	unsigned w = 0;
	if (!(raw & (1u << 11))) w |= T_DIAG;
	if (!(raw & (1u << 10))) w |= T_LEFT;
	if (!(raw & (1u << 9))) w |= T_UP;
	if (!(raw & (1u << 8))) w |= TW_IQ_INS;
	if (!(raw & (1u << 7))) w |= TW_IQ_EXT;
	if (!(raw & (1u << 6))) w |= TW_IT_INS;
	if (!(raw & (1u << 5))) w |= TW_IT_EXT;
	if (raw & (1u << 4)) w |= TW_M_NEG;
	else if (!(raw & (1u << 3))) w |= TW_M_ZERO;
	if (raw & (1u << 2)) w |= TW_IQ_NEG;
	if (raw & (1u << 1)) w |= TW_IT_NEG;
	if (!(raw & 1u)) w |= TW_CAND;
	return w;
}

template <int LQ, int NT>
struct ColMajorTraceFull {
	static constexpr bool kHasGapStates = true;
	const uint32_t *trace32;
	__device__ __forceinline__ unsigned raw(int i, int j) const
	{
		const uint32_t w = trace32[(size_t)((j - 1)*(LQ/2) + ((i - 1) >> 1))*NT];
		return ((i - 1) & 1) ? (w >> 16) : (w & 0xffffu);
	}
	__device__ __forceinline__ unsigned get(int i, int j) const { return decode_full_trace(raw(i, j)); }
};

template <int LQ, int NT>
__device__ __forceinline__ FastDpFull nc_fill_fast_full(const int32_t *__restrict__ tab, const int32_t *__restrict__ p5tab,
	uint64_t tlo, uint64_t thi, int Lt, uint32_t *__restrict__ trace32)
{
	static_assert(LQ % 2 == 0, "two rows share a trace store");
	int cM[LQ], cIq[LQ], cIt[LQ];
#pragma unroll
	for (int r = 0; r < LQ; ++r) { cM[r] = 0; cIq[r] = 0; cIt[r] = 0; }

	FastDpFull res;
	res.runmax = -1;
	res.last_i = res.last_j = 0;
	res.nmax = 0;

	int pt = 4; // GAP in front of the first column
	for (int j = 1; j <= Lt; ++j) {
		const int tb = packed_base(tlo, thi, j - 1);
		const int td = pt*4 + tb;
		const int32_t *__restrict__ ptd = tab + td;
		const int32_t *__restrict__ ptb = tab + tb;
		const int p5 = p5tab[td];
		uint32_t *__restrict__ col = trace32 + (size_t)(j - 1)*(LQ/2)*NT;

		int dM = 0, dIq = 0, dIt = 0; // (i-1, j-1)
		int uM = 0, uIt = 0;          // (i-1, j)
		unsigned pair_word = 0;
#pragma unroll
		for (int r = 0; r < LQ; ++r) {
			const int oM = cM[r], oIq = cIq[r], oIt = cIt[r]; // (i, j-1)

			const int d1 = dM - ptd[r*ROW_WORDS + ROW_P1];
			const int d2 = dIq - ptd[r*ROW_WORDS + ROW_P2];
			const int d3 = dIt - ptb[r*ROW_WORDS + ROW_P3];
			const int M = max(max(d1, d2), d3);

			const int qi = oM - ptd[r*ROW_WORDS + ROW_P4];
			const int qe = oIq - p5;
			const int Iq = max(qi, qe);

			const int ti = uM - ptb[r*ROW_WORDS + ROW_P6];
			const int te = uIt - tab[r*ROW_WORDS + ROW_P7];
			const int It = max(ti, te);

			const int rr = M - res.runmax;
			unsigned acc = 0;
			acc = __funnelshift_l((unsigned)(d1 - M), acc, 1);
			acc = __funnelshift_l((unsigned)(d2 - M), acc, 1);
			acc = __funnelshift_l((unsigned)(d3 - M), acc, 1);
			acc = __funnelshift_l((unsigned)(qi - Iq), acc, 1);
			acc = __funnelshift_l((unsigned)(qe - Iq), acc, 1);
			acc = __funnelshift_l((unsigned)(ti - It), acc, 1);
			acc = __funnelshift_l((unsigned)(te - It), acc, 1);
			acc = __funnelshift_l((unsigned)M, acc, 1);
			acc = __funnelshift_l((unsigned)(-M), acc, 1);
			acc = __funnelshift_l((unsigned)Iq, acc, 1);
			acc = __funnelshift_l((unsigned)It, acc, 1);
			acc = __funnelshift_l((unsigned)rr, acc, 1);

			if (rr > 0) { res.nmax = 1; res.last_i = r + 1; res.last_j = j; }
			else if (rr == 0) ++res.nmax;
			res.runmax = max(res.runmax, M);

			if (r & 1) col[(r >> 1)*NT] = pair_word | (acc << 16);
			else pair_word = acc;

			dM = oM; dIq = oIq; dIt = oIt;
			cM[r] = max(M, 0);
			cIq[r] = max(Iq, 0);
			cIt[r] = max(It, 0);
			uM = cM[r];
			uIt = cIt[r];
		}
		pt = tb;
	}
	return res;
}

// Maximal cells in the reference's (row-major) order from a fast fill.  The sweep visits the
// matrix column by column, so the cells tied with the maximum are gathered from the sweep
// position of the last strict raise onwards and then ordered by (row, column).
template <int LQ, int NT>
__device__ inline int collect_max_cells_full(const ColMajorTraceFull<LQ, NT> &tv, const FastDpFull &dp, int Lq, int Lt,
	uint16_t *cells, unsigned &flags)
{
	if (dp.nmax == 0) return 0;
	if (dp.nmax == 1 && dp.last_j > 0) {
		cells[0] = (uint16_t)((dp.last_i - 1)*Lt + (dp.last_j - 1));
		return 1;
	}
	int n = 0;
	const int j0 = dp.last_j > 0 ? dp.last_j : 1;
	for (int j = j0; j <= Lt && n < dp.nmax; ++j) {
		const int i0 = (j == j0 && dp.last_j > 0) ? dp.last_i : 1;
		for (int i = i0; i <= Lq && n < dp.nmax; ++i) {
			if (tv.raw(i, j) & 1u) continue; // below the running maximum
			if (n == MAX_MAXCELLS) return n; // not reached: the caller hands windows with more tied cells to the generic kernel
			// insertion sort by row-major index
			const uint16_t key = (uint16_t)((i - 1)*Lt + (j - 1));
			int k = n++;
			while (k > 0 && cells[k - 1] > key) { cells[k] = cells[k - 1]; --k; }
			cells[k] = key;
		}
	}
	return n;
}

// The same list for the row-major (generic) fill, at most MAX_MAXCELLS cells per call: `cursor`
// (start it at dp.last_raise, or 0) remembers where the scan stopped, `remaining` (start it at
// dp.nmax) how many cells are still to come.  Low-complexity windows tie hundreds of cells; the
// reference enumerates all of them (nuc_cruc.cpp:2531-2536), so does the caller, chunk by chunk.
template <int NT>
__device__ inline int collect_max_cells(const RowMajorTrace<NT> &tv, int Lq, int Lt, int &cursor, int &remaining, uint16_t *cells)
{
	int n = 0;
	const int ncell = Lq*Lt;
	for (; cursor < ncell && remaining > 0 && n < MAX_MAXCELLS; ++cursor) {
		if (!(tv.trace[(size_t)cursor*NT] & TW_CAND)) continue;
		cells[n++] = (uint16_t)cursor;
		--remaining;
	}
	if (cursor >= ncell) remaining = 0;
	return n;
}

} // namespace tnt
