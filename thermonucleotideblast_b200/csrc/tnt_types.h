// Shared host/device record layouts of the tntb200 engine (sm_100a).
#pragma once
#include <stdint.h>

namespace tnt {

// NucCruc base alphabet (reference nuc_cruc.h:179-188)
enum : int { bA = 0, bC, bG, bT, bI, bE, bGAP, bM, bR, bS, bV, bW, bY, bH, bK, bD, bB, bN, NB = 18 };

constexpr int NPAIR = 49;
constexpr int TABLE = NPAIR*NPAIR;
constexpr int MAX_OLIGO = 56;            // TNT_MAX_OLIGO_LEN
constexpr int NUM_FLANK = 4;             // tntblast.h:76
constexpr int MAX_WINDOW = MAX_OLIGO + 2*NUM_FLANK;   // 64 target bases at most
constexpr int MAX_COLS = MAX_OLIGO + MAX_WINDOW + 4;  // alignment columns incl. dangling ends
constexpr int MAX_LOOP = 512;
constexpr int NUM_HAIRPIN_LOOP = 131;     // NUM_SPECIAL_HAIRPIN_LOOP, nuc_cruc.h:587

// trace bits (reference nuc_cruc.h:62-65)
constexpr unsigned T_DIAG = 1, T_UP = 2, T_LEFT = 4, T_INVALID = 8;

// Thermodynamic tables for one (T, [Na+]); built on the host (thermo.cpp), resident in HBM.
struct Thermo {
	int32_t dg[TABLE];        // delta_g of update_dp_param (nuc_cruc.cpp:340-487)
	float H[TABLE];           // param_H / param_S (== loop terminal tables in the public reference)
	float S[TABLE];
	float loop_S[MAX_LOOP + 1];
	float bulge_S[MAX_LOOP + 1];
	uint8_t bbp[NB*NB];       // best_base_pair(x, y) = 7*resolve(x|y) + resolve(y|x)
	uint8_t wc[NPAIR + 3];    // watson_and_crick
	float T, log_na;
	float init_H, init_S, at_H, at_S, salt, asym_loop_dS, bulge_at_S;
	int32_t dangle5, dangle3;
	// hairpins (nuc_cruc.cpp:2301-2394): loop entropy by loop length, special tri- / tetra-loops
	float hairpin_S[MAX_LOOP + 1];
	char hairpin_loop[NUM_HAIRPIN_LOOP][8];
	float hairpin_special_H[NUM_HAIRPIN_LOOP];
	float hairpin_special_S[NUM_HAIRPIN_LOOP];
	// --dinkelbach (nuc_cruc.cpp:2399-2440): the penalty table is re-derived at a temperature of the
	// window's own (the Tm of the previous iteration), so the kernels evaluate update_dp_param entry by
	// entry: dg_class says which rule fills an entry (0: H - T*(S + salt correction); 1..6: the
	// supplementary terms below, clamped at zero), the rest are the T-independent inputs of the rules.
	uint8_t dg_class[TABLE + 3];
	float supp[12];           // param_supp (nuc_cruc.h:640)
	float supp_sc[4];         // salt correction * param_supp_salt[k]
	float salt_correction;    // param_SALT * log [Na+]
	int32_t dinkelbach;
};

// dg_class values
enum : int { DG_NN = 0, DG_LOOP = 1, DG_BULGE = 2, DG_TERM_AT = 3, DG_TERM_GC = 4, DG_TERM_INO = 5, DG_TERM_MM = 6 };

// One (oligo, strand) search unit.  `seq` is the oligo 5'->3' in NucCruc codes (the NucCruc
// "query"); the seed words are those of the oligo (minus strand) or of its reverse complement
// (plus strand), exactly as DNAHash_iterator::build_word_list makes them (seq_hash.h:287-374).
struct OligoStrand {
	uint8_t seq[MAX_OLIGO];
	uint16_t words[MAX_OLIGO];   // compacted word list
	int32_t len;
	int32_t nwords;
	int32_t plus;                // 1: binds the plus strand (target pushed as is)
	int32_t assay;               // index into the assay array
	int32_t role;                // TNT_OLIGO_F / _R / _P
	float r_log_ct;              // NC_R*logf(Ct) (nuc_cruc.cpp:2291), computed with the host libm
	float min_tm, max_tm, min_dg, max_dg;
	uint32_t clamp5, clamp3, max_gap, max_mismatch, max_poly_degen;
	int32_t lean_min_cols;       // gapless alignments with fewer columns cannot reach min_tm (thermo.cpp, lean_min_columns)
	int32_t pad_;
};

// Resident fragment descriptor.
struct Target {
	uint64_t base;       // global base index of position 0 (multiple of 64)
	uint32_t len;
	uint32_t pad;
	uint64_t exc_begin;  // range in the sparse non-ACGT list
	uint64_t exc_end;
};

// Seed candidate: 8 bytes.  target < 2^24 per engine, k < 256.
struct alignas(8) Candidate {   // 8-byte aligned: one 64-bit load / store per candidate
	uint32_t target_k;   // target | k << 24
	uint32_t t;          // seed position in the fragment
};

// Result of one alignment that passed the per-oligo filters (or every alignment in debug mode).
// The 48-byte head is all the assembly stage needs; the tail (aligned columns + window) is only
// fetched for the sites that end up in a reported hit, to render the alignment text.
struct BoundHead {
	uint32_t os;                 // global oligo-strand index (stage-1 set first, then stage-2 set)
	uint32_t target;
	int32_t loc5, loc3;
	float tm, dH, dS;
	int16_t anchor5, anchor3, num_mm, num_gap;
	uint32_t t;                  // seed position (oligo_info::target_loc)
	uint8_t k;                   // seed word index (oligo_info::query_loc)
	uint8_t flags;
	uint16_t align_len;          // length of the rendered alignment text
	uint32_t pad;
};
static_assert(sizeof(BoundHead) == 48, "BoundHead layout");

struct BoundRec {
	BoundHead h;
	float dG;
	int16_t poly_degen, valid;
	int32_t win_start, win_stop; // window in fragment coordinates
	int16_t fm_q, fm_t, lm_q, lm_t;
	uint8_t ncols, Lt, pad0, pad1;
	uint8_t cols_q[MAX_COLS];    // aligned columns (NucCruc codes), 5'->3' along the query
	uint8_t cols_t[MAX_COLS];
	uint8_t win[MAX_WINDOW];     // the NucCruc target (5'->3') the alignment ran against
};

constexpr uint8_t F_OOB = 1;        // the reference would have read out of range here
constexpr uint8_t F_STACK = 2;      // branch stack overflow (never expected)
constexpr uint8_t F_TRUNC = 4;      // alignment longer than MAX_COLS
constexpr uint8_t F_NEEDGENERIC = 8; // fast kernel only: this candidate needs the generic kernel (gap on an optimal path)

// Scan region of stage 2 (partner / probe search around a bound site)
struct Region {
	uint32_t target;
	uint32_t start, stop;    // fragment positions [start, stop)
	int32_t assay;
};

// Pairing output
struct PairRec {
	int32_t f, r, p;         // indices into the sorted bound-site array (-1: none)
};

} // namespace tnt
