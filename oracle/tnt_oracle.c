/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the tntblast v2.77 search hot path.
 *
 * Parity status: PINNED.  Every entry point is checked in tests/ against
 *   (1) the README known-answer vector of the reference (README.md:138,162-203), and
 *   (2) the unmodified reference compiled into oracle/_ref/libtntref.so by oracle/Makefile
 *       (differential runs over seeded random and planted inputs; fixtures generated from it
 *       are committed under tests/golden/).
 *
 * Only tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke() may
 * load this library, and only as the checker.  The product (thermonucleotideblast_b200/) never
 * links, imports or executes it.
 *
 * Each function cites the reference file:line it restates.  Data structures are deliberately
 * plain (flat arrays, index arithmetic) -- this is a specification of behaviour, including the
 * accidental behaviours listed in SURVEY.md section 8(a), not a translation of the C++.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tnt_oracle.h"
#include "santalucia_tables.inc"

/* ------------------------------------------------------------------------------------------
 * Alphabets
 * ---------------------------------------------------------------------------------------- */
/* nuc_cruc.h:179-188 */
enum { bA = 0, bC, bG, bT, bI, bE, bGAP, bM, bR, bS, bV, bW, bY, bH, bK, bD, bB, bN, NB = 18 };

#define NPAIR 49
#define PAIR(x, y) ((x)*7 + (y))         /* nuc_cruc.h:44 */
#define SIDX(prev, cur) ((prev)*NPAIR + (cur)) /* nuc_cruc.h:39 */
#define P_AT PAIR(bA, bT)
#define P_TA PAIR(bT, bA)
#define P_GT PAIR(bG, bT)
#define P_TG PAIR(bT, bG)
#define P_EE PAIR(bE, bE)
#define P_NONE PAIR(bGAP, bGAP)

#define IS_VIRTUAL(b) ((b) == bE || (b) == bGAP) /* nuc_cruc.h:52 */

/* trace bits, nuc_cruc.h:62-65 */
#define T_DIAG 1 /* im1_jm1 : query_target */
#define T_UP 2   /* im1_j   : query_gap (gap in target) */
#define T_LEFT 4 /* i_jm1   : gap_target (gap in query) */
#define T_INVALID 8

#define NUM_FLANK 4         /* tntblast.h:76 */
#define MAX_DP_PATH_ENUM 16 /* nuc_cruc.cpp:332 */
#define NC_ZERO_C 273.15f
#define NC_R 1.9872e-3f

static __thread char g_err[256];
static __thread long g_align_count;

const char *orc_last_error(void) { return g_err; }
long orc_last_alignment_count(void) { return g_align_count; }

static int fail(const char *msg)
{
	snprintf(g_err, sizeof(g_err), "%s", msg);
	return -1;
}

/* ------------------------------------------------------------------------------------------
 * Degenerate-base resolution (nuc_cruc.cpp:14-213)
 * A degenerate base opposite a plain A/C/G/T resolves to that base's complement when the
 * degeneracy allows it, otherwise to a fixed default.  `B` falls through into the `N` case in
 * the reference (missing break, nuc_cruc.cpp:163-197) and therefore behaves exactly like N.
 * ---------------------------------------------------------------------------------------- */
static const uint8_t DEGEN_ALLOWED[NB] = {
	/* bit0=A bit1=C bit2=G bit3=T */
	0, 0, 0, 0, 0, 0, 0,
	/* M */ 1 | 2, /* R */ 1 | 4, /* S */ 2 | 4, /* V */ 1 | 2 | 4, /* W */ 1 | 8, /* Y */ 2 | 8,
	/* H */ 1 | 2 | 8, /* K */ 4 | 8, /* D */ 1 | 4 | 8, /* B (acts as N) */ 15, /* N */ 15};
static const uint8_t DEGEN_DEFAULT[NB] = {
	0, 0, 0, 0, 0, 0, 0, bA, bA, bG, bA, bA, bT, bA, bT, bA, bA, bA};

static int resolve_base(int x, int other)
{
	if (x < bM) return x;
	if (other <= bT) {
		const int comp = 3 - other; /* A<->T, C<->G */
		if ((DEGEN_ALLOWED[x] >> comp) & 1) return comp;
	}
	return DEGEN_DEFAULT[x];
}

/* best_base_pair(first, second) = 7*resolve(first|second) + resolve(second|first) */
static uint8_t BBP[NB][NB];
static uint8_t WC[NPAIR];
static int g_static_init;

static void static_init(void)
{
	if (g_static_init) return;
	for (int x = 0; x < NB; ++x)
		for (int y = 0; y < NB; ++y)
			BBP[x][y] = (uint8_t)PAIR(resolve_base(x, y), resolve_base(y, x));
	memset(WC, 0, sizeof(WC));
	/* nuc_cruc.cpp:229-238 */
	WC[PAIR(bA, bT)] = WC[PAIR(bT, bA)] = WC[PAIR(bC, bG)] = WC[PAIR(bG, bC)] = 1;
	for (int b = bA; b <= bT; ++b) WC[PAIR(b, bI)] = WC[PAIR(bI, b)] = 1;
	WC[PAIR(bI, bI)] = 1;
	g_static_init = 1;
}

/* is_complemetary_base (nuc_cruc_anchor.cpp:8-139): IUPAC-aware; I == N; virtual bases match nothing */
static const uint8_t BASE_SET[NB] = {
	/* bit0=A bit1=C bit2=G bit3=T */
	1, 2, 4, 8, 15, 0, 0, 1 | 2, 1 | 4, 2 | 4, 1 | 2 | 4, 1 | 8, 2 | 8, 1 | 2 | 8, 4 | 8, 1 | 4 | 8, 2 | 4 | 8, 15};

static int is_complementary(int q, int t)
{
	const unsigned ts = BASE_SET[t];
	/* complement of a base set: A<->T (bit0<->bit3), C<->G (bit1<->bit2) */
	const unsigned tc = ((ts & 1) << 3) | ((ts & 8) >> 3) | ((ts & 2) << 1) | ((ts & 4) >> 1);
	return (BASE_SET[q] & tc) != 0;
}

static int ascii_to_base(char c) /* nuc_cruc.h:190-231 */
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
	}
	return -1;
}

/* seq.h DB_* code -> NucCruc base, and its complement (bind_oligo.cpp:524-591, :1227-1294).
 * Codes above DB_N (GAP=16, UNKNOWN=17) are silently skipped by the window loader. */
static const int8_t DB_TO_BASE[18] = {bA, bC, bG, bT, bI, bM, bR, bS, bV, bW, bY, bH, bK, bD, bB, bN, -1, -1};
static const int8_t DB_TO_COMP[18] = {bT, bG, bC, bA, bI, bK, bY, bS, bB, bW, bR, bD, bM, bH, bV, bN, -1, -1};
static const char BASE_CHAR[] = "ACGTI$-MRSVWYHKDBN"; /* nuc_cruc_output.cpp:11 */

/* ------------------------------------------------------------------------------------------
 * Thermodynamic context for one (T, [Na+])
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	float T, na, log_na;
	int32_t dg[NPAIR*NPAIR];
} thermo_t;

/* update_dp_param (nuc_cruc.cpp:340-487) */
static void thermo_init(thermo_t *th, float T, float na)
{
	static_init();
	th->T = T;
	th->na = na;
	th->log_na = logf(na);

	const float salt_correction = SL_SALT*th->log_na;
	const float loop_sc = salt_correction*SL_SUPP_SALT[0];
	const float bulge_sc = salt_correction*SL_SUPP_SALT[1];
	const float term_match_sc = salt_correction*SL_SUPP_SALT[2];
	const float term_mismatch_sc = salt_correction*SL_SUPP_SALT[3];

#define SCALE(X) ((int32_t)((X)*10000.0f))
#define CLAMP0(v) ((v) > 0 ? (v) : 0)
	for (int i = 0; i < NPAIR*NPAIR; ++i)
		th->dg[i] = SCALE(SL_PARAM_H[i] - T*(SL_PARAM_S[i] + salt_correction));

	const int32_t term_at = CLAMP0(SCALE(SL_SUPP[4] - T*(SL_SUPP[5] + term_match_sc)));
	const int32_t term_gc = CLAMP0(SCALE(SL_SUPP[6] - T*(SL_SUPP[7] + term_match_sc)));
	const int32_t term_i = CLAMP0(SCALE(SL_SUPP[8] - T*(SL_SUPP[9] + term_match_sc)));
	const int32_t term_mm = CLAMP0(SCALE(SL_SUPP[10] - T*(SL_SUPP[11] + term_mismatch_sc)));
	const int32_t loop = CLAMP0(SCALE(SL_SUPP[0] - T*(SL_SUPP[1] + loop_sc)));
	const int32_t bulge = CLAMP0(SCALE(SL_SUPP[2] - T*(SL_SUPP[3] + bulge_sc)));

	for (int i = bA; i <= bI; ++i) {
		for (int j = bA; j <= bI; ++j) {
			const int curr = PAIR(i, j);
			int32_t v;
			if (WC[curr]) {
				if (curr == P_AT || curr == P_TA) v = term_at;
				else if (curr == PAIR(bG, bC) || curr == PAIR(bC, bG)) v = term_gc;
				else v = term_i;
			}
			else v = term_mm;
			for (int k = bA; k <= bI; ++k) {
				const int prev1 = PAIR(k, bGAP), prev2 = PAIR(bGAP, k);
				th->dg[SIDX(curr, prev1)] = th->dg[SIDX(prev1, curr)] = v;
				th->dg[SIDX(curr, prev2)] = th->dg[SIDX(prev2, curr)] = v;
			}
			for (int k = bA; k <= bI; ++k)
				for (int l = bA; l <= bI; ++l) {
					const int prev = PAIR(k, l);
					if (!WC[curr] && !WC[prev]) th->dg[SIDX(curr, prev)] = loop;
				}
		}
	}
	for (int i = bA; i <= bI; ++i)
		for (int j = bA; j <= bI; ++j) {
			th->dg[SIDX(PAIR(i, bGAP), PAIR(j, bGAP))] = bulge;
			th->dg[SIDX(PAIR(bGAP, i), PAIR(bGAP, j))] = bulge;
		}
#undef SCALE
#undef CLAMP0
}

int orc_dump_tables(float T, float na, ref_tables *out)
{
	thermo_t th;
	thermo_init(&th, T, na);
	memset(out, 0, sizeof(*out));
	memcpy(out->delta_g, th.dg, sizeof(th.dg));
	memcpy(out->param_H, SL_PARAM_H, sizeof(SL_PARAM_H));
	memcpy(out->param_S, SL_PARAM_S, sizeof(SL_PARAM_S));
	memcpy(out->loop_terminal_H, SL_PARAM_H, sizeof(SL_PARAM_H)); /* nuc_cruc_santa_lucia.cpp:591-603 */
	memcpy(out->loop_terminal_S, SL_PARAM_S, sizeof(SL_PARAM_S));
	memcpy(out->loop_S, SL_LOOP_S, sizeof(SL_LOOP_S));
	memcpy(out->bulge_S, SL_BULGE_S, sizeof(SL_BULGE_S));
	memcpy(out->supp, SL_SUPP, sizeof(SL_SUPP));
	memcpy(out->supp_salt, SL_SUPP_SALT, sizeof(SL_SUPP_SALT));
	out->init_H = SL_INIT_H;
	out->init_S = SL_INIT_S;
	out->AT_closing_H = SL_AT_CLOSING_H;
	out->AT_closing_S = SL_AT_CLOSING_S;
	out->symmetry_S = SL_SYMMETRY_S;
	out->SALT = SL_SALT;
	out->asymmetric_loop_dS = SL_ASYMMETRIC_LOOP_DS;
	out->bulge_AT_closing_S = SL_BULGE_AT_CLOSING_S;
	memcpy(out->watson_and_crick, WC, NPAIR);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * NucCruc heterodimer alignment
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	int32_t M, Iq, It;
	uint8_t Mt, Iqt, Itt;
} cell_t;

#define MAXSEQ 1024
#define ALN_CAP (2*MAXSEQ + 16)

typedef struct {
	int valid;
	float dH, dS, tm;
	/* aligned columns live in col[b..e) so that both ends can grow */
	uint8_t q[ALN_CAP], t[ALN_CAP];
	int b, e;
	int fm_q, fm_t; /* first_match (query idx, target idx) */
	int lm_q, lm_t; /* last_match */
} aln_t;

static void aln_clear(aln_t *a)
{
	a->valid = 0;
	a->dH = a->dS = a->tm = 0.0f;
	a->b = a->e = ALN_CAP/2;
	a->fm_q = a->fm_t = a->lm_q = a->lm_t = 0;
}
#define ALN_N(a) ((a)->e - (a)->b)

typedef struct {
	const thermo_t *th;
	float strand;
	int dangle5, dangle3;
	int Lq, Lt;
	uint8_t q[MAXSEQ], t[MAXSEQ];
	cell_t *dp; /* (Lq+1) x (Lt+1) */
	int stride;
	int *max_cells;
	int n_max, cap_max;
	aln_t best;
	int oob; /* an unchecked out-of-range read of the reference would have happened */
	int homo; /* HOMO_DIMER mode: the symmetry entropy joins the initiation term (nuc_cruc.cpp:1632) */
	int tri;  /* > 0: align_hairpin (nuc_cruc.cpp:771-971): the same recurrence over the triangle
	           * i + j <= tri + 1 (tri = max_stem_len) of the query against itself */
	int hairpin; /* HAIRPIN mode of evaluate_alignment: no initiation term (the caller preloads dH / dS), Tm = dH/dS */
} nc_t;

/* align_dimer (nuc_cruc.cpp:492-696).  Rows follow the *reversed* query, columns the target. */
static int32_t align_dimer(nc_t *nc)
{
	const int Lq = nc->Lq, Lt = nc->Lt, st = Lt + 1;
	const int32_t *dg = nc->th->dg;
	nc->stride = st;
	nc->dp = (cell_t *)realloc(nc->dp, sizeof(cell_t)*(size_t)(Lq + 1)*(size_t)st);
	for (long k = 0; k < (long)(Lq + 1)*st; ++k) {
		/* NC_Elem default state (nuc_cruc.h:531-536); row 0 / column 0 keep it */
		nc->dp[k].M = nc->dp[k].Iq = nc->dp[k].It = -1;
		nc->dp[k].Mt = nc->dp[k].Iqt = nc->dp[k].Itt = T_INVALID;
	}
	nc->n_max = 0;
	int32_t max_score = -1;

#define POS(v) ((v) > 0 ? (v) : 0)
	const int rows = nc->tri > 0 ? nc->tri : (nc->tri < 0 ? 0 : Lq);
	for (int i = 1; i <= rows; ++i) {
		const int qb = nc->q[Lq - i];
		const int pq = (i == 1) ? bGAP : nc->q[Lq - (i - 1)];
		const int cols = nc->tri > 0 ? nc->tri - (i - 1) : Lt;
		for (int j = 1; j <= cols; ++j) {
			const int tb = nc->t[j - 1];
			const int pt = (j == 1) ? bGAP : nc->t[j - 2];
			cell_t *X = &nc->dp[i*st + j];
			const cell_t *A = &nc->dp[(i - 1)*st + (j - 1)];
			const cell_t *B = &nc->dp[(i - 1)*st + j];
			const cell_t *Cc = &nc->dp[i*st + (j - 1)];

			int cur = BBP[tb][qb];
			const int32_t dg1 = POS(A->M) - dg[SIDX(BBP[pt][pq], cur)];
			const int32_t dg2 = POS(A->Iq) - dg[SIDX(BBP[pt][bGAP], cur)];
			const int32_t dg3 = POS(A->It) - dg[SIDX(BBP[bGAP][pq], cur)];

			if (dg1 >= dg2) {
				if (dg1 >= dg3) {
					X->M = dg1;
					X->Mt = T_DIAG;
					if (dg1 == dg2) X->Mt |= T_LEFT;
					if (dg1 == dg3) X->Mt |= T_UP;
				}
				else { X->M = dg3; X->Mt = T_UP; }
			}
			else {
				if (dg2 >= dg3) {
					X->M = dg2;
					X->Mt = T_LEFT;
					if (dg2 == dg3) X->Mt |= T_UP;
				}
				else { X->M = dg3; X->Mt = T_UP; }
			}

			/* gap in the query (target base opposite a gap): comes from the left cell */
			cur = BBP[tb][bGAP];
			int32_t ins = POS(Cc->M) - dg[SIDX(BBP[pt][qb], cur)];
			int32_t ext = POS(Cc->Iq) - dg[SIDX(BBP[pt][bGAP], cur)];
			if (ins >= ext) { X->Iq = ins; X->Iqt = T_DIAG | (ins == ext ? T_LEFT : 0); }
			else { X->Iq = ext; X->Iqt = T_LEFT; }

			/* gap in the target (query base opposite a gap): comes from the upper cell */
			cur = BBP[bGAP][qb];
			ins = POS(B->M) - dg[SIDX(BBP[tb][pq], cur)];
			ext = POS(B->It) - dg[SIDX(BBP[bGAP][pq], cur)];
			if (ins >= ext) { X->It = ins; X->Itt = T_DIAG | (ins == ext ? T_UP : 0); }
			else { X->It = ext; X->Itt = T_UP; }

			/* all cells that reach the maximum, in row-major order (nuc_cruc.cpp:670-691) */
			if (X->M >= max_score) {
				if (X->M > max_score) { max_score = X->M; nc->n_max = 0; }
				if (nc->n_max == nc->cap_max) {
					nc->cap_max = nc->cap_max ? 2*nc->cap_max : 64;
					nc->max_cells = (int *)realloc(nc->max_cells, sizeof(int)*nc->cap_max);
				}
				nc->max_cells[nc->n_max++] = i*st + j;
			}
		}
	}
#undef POS
	return max_score;
}

/* One pending branch point of the traceback (trace_branch, nuc_cruc.h:279-341).  `id`
 * identifies which trace byte of which cell the branch belongs to (the reference compares
 * addresses). */
typedef struct {
	int id;
	uint8_t mask, cur;
} branch_t;

static int path_split(unsigned m) /* nuc_cruc.h:73 */
{
	return ((m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1)) > 1;
}

static int branch_next(branch_t *b) /* trace_branch::next_trace */
{
	while ((b->cur = (uint8_t)(b->cur << 1)) < T_INVALID)
		if (b->cur & b->mask) return 1;
	return 0;
}

static uint8_t q_at(nc_t *nc, int idx)
{
	if (idx < 0 || idx >= nc->Lq) { nc->oob = 1; return bGAP; }
	return nc->q[idx];
}

static uint8_t t_at(nc_t *nc, int idx)
{
	if (idx < 0 || idx >= nc->Lt) { nc->oob = 1; return bGAP; }
	return nc->t[idx];
}

static void aln_push_back(aln_t *a, int q, int t)
{
	if (a->e < ALN_CAP) { a->q[a->e] = (uint8_t)q; a->t[a->e] = (uint8_t)t; a->e++; }
}

/* trace_back (nuc_cruc.cpp:1409-1618) */
static int trace_back(nc_t *nc, int cell, branch_t *stack, int *nstack, int *zero_count, aln_t *a)
{
	const int Lq = nc->Lq, st = nc->stride;
	int last_i = cell/st, last_j = cell%st;
	a->fm_q = Lq - last_i;
	a->fm_t = last_j - 1;

	int truncate_at_zero = 0, count_zeros = 0;
	if (*zero_count < 0) { *zero_count = 0; count_zeros = 1; }
	else truncate_at_zero = (*zero_count)--;

	int cur_id = -1;           /* the static `first_match` byte */
	unsigned cur_mask = T_DIAG;

	for (;;) {
		int valid = 1;
		unsigned local;
		if (path_split(cur_mask)) {
			int k;
			for (k = 0; k < *nstack; ++k)
				if (stack[k].id == cur_id) break;
			if (k == *nstack) {
				stack[k].id = cur_id;
				stack[k].mask = (uint8_t)cur_mask;
				stack[k].cur = (cur_mask & T_DIAG) ? T_DIAG : ((cur_mask & T_UP) ? T_UP : T_LEFT);
				(*nstack)++;
			}
			local = stack[k].cur;
		}
		else local = cur_mask;

		const cell_t *c = &nc->dp[last_i*st + last_j];
		switch (local) {
		case T_DIAG:
			if (last_i > Lq || last_j < 1) valid = 0;
			else {
				if (c->M < 0) valid = 0;
				else if (c->M == 0) {
					if (count_zeros) (*zero_count)++;
					else if (--truncate_at_zero == 0) valid = 0;
				}
				aln_push_back(a, q_at(nc, Lq - last_i), t_at(nc, last_j - 1));
				a->lm_q = Lq - last_i;
				a->lm_t = last_j - 1;
				cur_id = (last_i*st + last_j)*3 + 0;
				cur_mask = c->Mt;
				--last_i;
				--last_j;
			}
			break;
		case T_LEFT: /* gap_target: a gap is inserted in the query */
			if (last_j < 1) valid = 0;
			else {
				if (c->Iq < 0) valid = 0;
				aln_push_back(a, bGAP, t_at(nc, last_j - 1));
				a->lm_q = Lq - last_i + 1;
				a->lm_t = last_j - 1;
				cur_id = (last_i*st + last_j)*3 + 1;
				cur_mask = c->Iqt;
				--last_j;
			}
			break;
		case T_UP: /* query_gap: a gap is inserted in the target */
			if (last_i > Lq) valid = 0;
			else {
				if (c->It < 0) valid = 0;
				aln_push_back(a, q_at(nc, Lq - last_i), bGAP);
				a->lm_q = Lq - last_i;
				a->lm_t = last_j;
				cur_id = (last_i*st + last_j)*3 + 2;
				cur_mask = c->Itt;
				--last_i;
			}
			break;
		default:
			return fail("invalid_match in trace back");
		}
		if (!valid) break;
		if (last_i < 0 || last_j < 0) { nc->oob = 1; break; }
	}
	return 0;
}

/* has_AT_initiation (nuc_cruc.cpp:2888-2905); k indexes into the alignment columns */
static int has_AT_initiation(const uint8_t *q, const uint8_t *t, int k)
{
	do { --k; } while (k != 0 && (q[k] == bGAP || t[k] == bGAP));
	const int bp = BBP[q[k]][t[k]];
	return bp == P_AT || bp == P_TA;
}

/* evaluate_alignment (nuc_cruc.cpp:1620-2299), HETERO_DIMER mode.  Pair index here is
 * 7*query + target, stepping 5'->3' along the query.  Summation order follows the reference
 * statement by statement; all arithmetic is IEEE binary32 without contraction. */
static int evaluate_alignment(const nc_t *nc, aln_t *a)
{
	const uint8_t *q = a->q + a->b, *t = a->t + a->b;
	const int n = ALN_N(a);
	const float *H = SL_PARAM_H, *S = SL_PARAM_S;
	const float *LTH = SL_PARAM_H, *LTS = SL_PARAM_S; /* loop-terminal tables are copies */

	int terminal = P_NONE, last_last = P_NONE, last = P_NONE, cur;
	float dH = SL_INIT_H, dS = SL_INIT_S + (nc->homo ? SL_SYMMETRY_S : 0.0f);
	if (nc->hairpin) { dH = a->dH; dS = a->dS; } /* :1627-1633: hairpins do not pay the initiation cost */
	unsigned nqgap = 0, ntgap = 0, nmm = 0, num_base = 0;
	int terminal_5 = 0;

	cur = BBP[q[0]][t[0]];
	if (WC[cur]) {
		terminal_5 = 1;
		if (cur == P_AT || cur == P_TA) { dH += SL_AT_CLOSING_H; dS += SL_AT_CLOSING_S; }
	}
	num_base += IS_VIRTUAL(q[0]) ? 0 : 1;
	num_base += IS_VIRTUAL(t[0]) ? 0 : 1;

	for (int k = 1; k < n; ++k) {
		last_last = last;
		last = cur;
		cur = BBP[q[k]][t[k]];
		const int align_start = (k == 1), align_stop = (k == n - 1);
		const int in_loop = (q[k] == bGAP) || (t[k] == bGAP) || (!WC[last] && !WC[cur]);
#define NONVIRT_PAIR(p) (((p)%7 < bE) && ((p)/7 < bE))
		if (!in_loop) {
			if (align_start && !WC[last] && NONVIRT_PAIR(last)) {
				/* frayed 5' end == two dangling ends */
				int tmp = BBP[last/7][bE];
				dH += H[SIDX(tmp, cur)]; dS += S[SIDX(tmp, cur)];
				tmp = BBP[bE][last%7];
				dH += H[SIDX(tmp, cur)]; dS += S[SIDX(tmp, cur)];
			}
			else if (align_stop && !WC[cur] && NONVIRT_PAIR(cur)) {
				int tmp = BBP[q[k]][bE];
				dH += H[SIDX(last, tmp)]; dS += S[SIDX(last, tmp)];
				tmp = BBP[bE][t[k]];
				dH += H[SIDX(last, tmp)]; dS += S[SIDX(last, tmp)];
			}
			else { dH += H[SIDX(last, cur)]; dS += S[SIDX(last, cur)]; }
			num_base += IS_VIRTUAL(q[k]) ? 0 : 1;
			num_base += IS_VIRTUAL(t[k]) ? 0 : 1;
		}

		if (WC[cur] || cur == P_EE) {
			terminal = cur;
			if (!terminal_5) {
				terminal_5 = 1;
				if (cur == P_AT || cur == P_TA) { dH += SL_AT_CLOSING_H; dS += SL_AT_CLOSING_S; }
			}
			const unsigned max_gap = nqgap > ntgap ? nqgap : ntgap;

			if (nmm > 1 || (max_gap > 0 && nmm == 1)) {
				/* closing an internal loop */
				const unsigned gap_diff = nqgap > ntgap ? nqgap - ntgap : ntgap - nqgap;
				const unsigned loop_size = nmm*2 + gap_diff;
				if (loop_size == 2 && (last == P_GT || last == P_TG) && (last_last == P_GT || last_last == P_TG)) {
					dH += H[SIDX(last_last, last)]; dS += S[SIDX(last_last, last)];
					num_base += 2;
				}
				else {
					dS += SL_LOOP_S[loop_size];
					dS += gap_diff*SL_ASYMMETRIC_LOOP_DS;
					int rhs_q = k - 1, rhs_t = k - 1;
					dH -= H[SIDX(last, cur)]; dS -= S[SIDX(last, cur)];
					const int last_has_gap = (last%7 == bGAP) || (last/7 >= bGAP);
					if (!last_has_gap) { dH += LTH[SIDX(last, cur)]; dS += LTS[SIDX(last, cur)]; }
					else {
						int mm = P_NONE;
						if (last/7 == bGAP) {
							for (;;) {
								if (!IS_VIRTUAL(q[rhs_q])) { mm = BBP[q[rhs_q]][last%7]; break; }
								if (rhs_q == 0) break;
								--rhs_q;
							}
						}
						else {
							for (;;) {
								if (!IS_VIRTUAL(t[rhs_t])) { mm = BBP[last/7][t[rhs_t]]; break; }
								if (rhs_t == 0) break;
								--rhs_t;
							}
						}
						dH += LTH[SIDX(mm, cur)]; dS += LTS[SIDX(mm, cur)];
					}
					/* left terminal mismatch: walk back to the closest Watson-Crick column */
					int lhs_q = k - 1, lhs_t = k - 1;
					for (;;) {
						const int pm = BBP[q[lhs_q]][t[lhs_t]];
						if (WC[pm]) {
							++lhs_q; ++lhs_t;
							if (q[lhs_q] != bGAP && t[lhs_t] != bGAP) {
								const int mm = BBP[q[lhs_q]][t[lhs_t]];
								dH -= H[SIDX(pm, mm)]; dS -= S[SIDX(pm, mm)];
							}
							else {
								num_base += 2;
								while (q[lhs_q] == bGAP) ++lhs_q;
								while (t[lhs_t] == bGAP) ++lhs_t;
							}
							const int mm = BBP[q[lhs_q]][t[lhs_t]];
							dH += LTH[SIDX(pm, mm)]; dS += LTS[SIDX(pm, mm)];
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
				const unsigned bulge = nqgap > ntgap ? nqgap : ntgap;
				if (bulge == 1) { dH += H[SIDX(last_last, cur)]; dS += S[SIDX(last_last, cur)]; }
				dS += SL_BULGE_S[bulge];
				/* UNAFOLD_COMPATIBILITY (nuc_cruc.h:28): no AT penalty for single-base bulges */
				if (bulge != 1 && (q[k] == bA || q[k] == bT)) dS += SL_BULGE_AT_CLOSING_S;
				if (bulge != 1 && has_AT_initiation(q, t, k)) dS += SL_BULGE_AT_CLOSING_S;
			}
			nqgap = ntgap = nmm = 0;
		}
		else nmm += (!IS_VIRTUAL(q[k]) && !IS_VIRTUAL(t[k])) ? 1 : 0;

		nqgap += (q[k] == bGAP) ? 1 : 0;
		ntgap += (t[k] == bGAP) ? 1 : 0;
	}

	if (terminal == P_AT || terminal == P_TA) { dH += SL_AT_CLOSING_H; dS += SL_AT_CLOSING_S; }

	a->dH = dH;
	a->dS = dS;
	if (dH >= 0.0f) return 0; /* binding must be enthalpically driven */

	dS += SL_SALT*(0.5f*num_base - 1)*nc->th->log_na;
	a->dS = dS;
	const float tm = nc->hairpin ? dH/dS - NC_ZERO_C /* :2286-2289: no strand concentration */
		: dH/(NC_R*logf(nc->strand*1.0f) + dS) - NC_ZERO_C;
	a->tm = tm > 0.0f ? tm : 0.0f;
	return 1;
}

/* enumerate_dimer_alignments (nuc_cruc.cpp:973-1170) for one maximal cell */
static int enumerate_alignments(nc_t *nc, int max_cell)
{
	const float T = nc->th->T;
	const int Lq = nc->Lq, Lt = nc->Lt;
	int first_time = 1, nstack = 0, zero_count = -1;
	unsigned trace_count = 0;
	float best_dg = nc->best.dH - T*nc->best.dS;
	branch_t *stack = (branch_t *)malloc(sizeof(branch_t)*(size_t)(3*(Lq + Lt) + 8));
	aln_t *a = (aln_t *)malloc(sizeof(aln_t));

	for (;;) {
		if (!first_time && nstack == 0 && zero_count <= 0) break;
		if (MAX_DP_PATH_ENUM != 0 && MAX_DP_PATH_ENUM < trace_count) break;
		++trace_count;
		first_time = 0;

		aln_clear(a);
		if (trace_back(nc, max_cell, stack, &nstack, &zero_count, a) < 0) { free(stack); free(a); return -1; }

		/* trim frayed (non Watson-Crick) ends, back then front (:1022-1054) */
		while (ALN_N(a) > 0 && !WC[BBP[a->q[a->e - 1]][a->t[a->e - 1]]]) {
			if (!IS_VIRTUAL(a->q[a->e - 1])) --a->lm_q;
			if (!IS_VIRTUAL(a->t[a->e - 1])) ++a->lm_t;
			--a->e;
		}
		while (ALN_N(a) > 0 && !WC[BBP[a->q[a->b]][a->t[a->b]]]) {
			if (!IS_VIRTUAL(a->q[a->b])) ++a->fm_q;
			if (!IS_VIRTUAL(a->t[a->b])) --a->fm_t;
			++a->b;
		}

		if (zero_count == 0 && nstack > 0) {
			while (nstack > 0 && !branch_next(&stack[nstack - 1])) --nstack;
			zero_count = -1;
		}

		/* optional dangling-end virtual bases (:1088-1137) */
		if (nc->dangle5 && (a->fm_q != 0 || a->fm_t != Lt - 1)) {
			int qb, tb;
			if (a->fm_q == 0) qb = bE;
			else { --a->fm_q; qb = q_at(nc, a->fm_q); }
			if (a->fm_t == Lt - 1) tb = bE;
			else { ++a->fm_t; tb = t_at(nc, a->fm_t); }
			if (a->b > 0) { --a->b; a->q[a->b] = (uint8_t)qb; a->t[a->b] = (uint8_t)tb; }
		}
		if (nc->dangle3 && (a->lm_q != Lq - 1 || a->lm_t != 0)) {
			int qb, tb;
			if (a->lm_q == Lq - 1) qb = bE;
			else { ++a->lm_q; qb = q_at(nc, a->lm_q); }
			if (a->lm_t == 0) tb = bE;
			else { --a->lm_t; tb = t_at(nc, a->lm_t); }
			aln_push_back(a, qb, tb);
		}

		if (ALN_N(a) < 3) continue;

		if (evaluate_alignment(nc, a)) {
			const float local_dg = a->dH - T*a->dS;
			if (!nc->best.valid || local_dg < best_dg) {
				nc->best = *a;
				nc->best.valid = 1;
				best_dg = local_dg;
			}
		}
	}
	free(stack);
	free(a);
	return 0;
}

/* NucCruc::dinkelbach(bool) (nuc_cruc.h:763-766): per-thread switch of this library, like the one of the harness */
static __thread int g_dinkelbach;
void orc_set_dinkelbach(int on) { g_dinkelbach = on ? 1 : 0; }

/* approximate_tm_heterodimer / _homodimer (nuc_cruc.cpp:2397-2516) + tm_dimer (:2517-2540).
 * With use_dinkelbach (:2399-2440): align at 0 degC, then at the melting temperature of the previous
 * pass, while delta_G() of the pass -- dH - target_T*dS at the temperature it was aligned at -- is
 * negative and still rising; temperature() re-derives the penalty table every time (nuc_cruc.h:1232-1242).
 * The alignment of the last pass stays; the initial temperature is restored at the end. */
static int nc_run(nc_t *nc, float *dp_dg)
{
	++g_align_count;
	nc->oob = 0;
	if (!g_dinkelbach) {
		aln_clear(&nc->best);
		const int32_t max_score = align_dimer(nc);
		for (int k = 0; k < nc->n_max; ++k)
			if (enumerate_alignments(nc, nc->max_cells[k]) < 0) return -1;
		if (dp_dg) *dp_dg = -((float)max_score/10000.0f);
		return 0;
	}
	const thermo_t *keep = nc->th;
	thermo_t *th = (thermo_t *)malloc(sizeof(thermo_t));
	float q = -999999.9f, last_q, local_tm;
	int32_t max_score = 0;
	int rc = 0;
	thermo_init(th, 273.15f, keep->na);
	nc->th = th;
	do {
		aln_clear(&nc->best);
		max_score = align_dimer(nc);
		for (int k = 0; k < nc->n_max && rc == 0; ++k)
			if (enumerate_alignments(nc, nc->max_cells[k]) < 0) rc = -1;
		if (rc < 0) break;
		local_tm = nc->best.tm;
		last_q = q;
		q = nc->best.dH - th->T*nc->best.dS;
		thermo_init(th, 273.15f + local_tm, keep->na);
	} while (q < 0.0 && q > last_q);
	nc->th = keep;
	free(th);
	if (dp_dg) *dp_dg = -((float)max_score/10000.0f);
	return rc;
}

/* anchor5_query / anchor3_query (nuc_cruc_anchor.cpp:143-192, :249-298) */
static unsigned anchor5_query(const nc_t *nc)
{
	const aln_t *a = &nc->best;
	unsigned anchor = 0;
	int qi = 0, ti = a->fm_q + a->fm_t;
	if (ALN_N(a) > 0 && a->t[a->b] == bE) return 0;
	if (ALN_N(a) > 0 && a->q[a->b] == bE) --ti;
	if (ti >= nc->Lt) return 0;
	for (;;) {
		if (qi >= nc->Lq || ti < 0) return anchor;
		if (!is_complementary(nc->q[qi], nc->t[ti])) return anchor;
		++anchor; ++qi; --ti;
	}
}

static unsigned anchor3_query(const nc_t *nc)
{
	const aln_t *a = &nc->best;
	unsigned anchor = 0;
	int qi = nc->Lq - 1, ti = (a->lm_q + a->lm_t + 1) - nc->Lq;
	if (ALN_N(a) > 0 && a->t[a->e - 1] == bE) return 0;
	if (ALN_N(a) > 0 && a->q[a->e - 1] == bE) ++ti;
	if (ti >= nc->Lt || ti < 0) return 0;
	for (;;) {
		if (qi < 0 || ti >= nc->Lt) return anchor;
		if (!is_complementary(nc->q[qi], nc->t[ti])) return anchor;
		++anchor; --qi; ++ti;
	}
}

/* num_mismatch_by_query / num_gap / max_contiguous_target_degen (nuc_cruc.h:389-483) */
static unsigned count_mismatch(const nc_t *nc)
{
	const aln_t *a = &nc->best;
	unsigned mm = 0, aligned = 0;
	for (int k = a->b; k < a->e; ++k) {
		if (!IS_VIRTUAL(a->q[k])) {
			if (!IS_VIRTUAL(a->t[k]) && !is_complementary(a->q[k], a->t[k])) ++mm;
			++aligned;
		}
	}
	return mm + (unsigned)nc->Lq - aligned;
}

static unsigned count_gap(const nc_t *nc)
{
	const aln_t *a = &nc->best;
	unsigned g = 0;
	for (int k = a->b; k < a->e; ++k) g += (a->q[k] == bGAP) + (a->t[k] == bGAP);
	return g;
}

static unsigned max_target_degen(const nc_t *nc)
{
	const aln_t *a = &nc->best;
	unsigned best = 0, run = 0;
	for (int k = a->b; k < a->e; ++k) {
		if (a->t[k] >= bM && a->t[k] <= bN) { if (++run > best) best = run; }
		else run = 0;
	}
	return best;
}

/* operator<< for dimers (nuc_cruc_output.cpp:74-205) */
static void render_alignment(nc_t *nc, char *out, size_t cap)
{
	const aln_t *a = &nc->best;
	const int Lq = nc->Lq, Lt = nc->Lt;
	int prefix = a->fm_q < Lt - 1 - a->fm_t ? a->fm_q : Lt - 1 - a->fm_t;
	if (prefix < 0) prefix = 0;
	int suffix = Lq - 1 - a->lm_q < a->lm_t ? Lq - 1 - a->lm_q : a->lm_t;
	if (suffix < 0) suffix = 0;

	size_t n = 0;
#define PUT(c) do { if (n + 1 < cap) out[n++] = (c); } while (0)
#define PUTS(s) do { for (const char *_p = (s); *_p; ++_p) PUT(*_p); } while (0)
	PUTS("5' ");
	for (int i = 0; i < prefix; ++i) PUT(BASE_CHAR[q_at(nc, a->fm_q - prefix + i)]);
	for (int k = a->b; k < a->e; ++k) PUT(BASE_CHAR[a->q[k]]);
	for (int i = 0; i < suffix; ++i) PUT(BASE_CHAR[q_at(nc, a->lm_q + 1 + i)]);
	PUTS(" 3'\n   ");
	for (int i = 0; i < prefix; ++i)
		PUT(is_complementary(q_at(nc, a->fm_q - prefix + i), t_at(nc, a->fm_t + prefix - i)) ? ':' : ' ');
	for (int k = a->b; k < a->e; ++k) PUT(is_complementary(a->t[k], a->q[k]) ? '|' : ' ');
	for (int i = 0; i < suffix; ++i)
		PUT(is_complementary(q_at(nc, a->lm_q + 1 + i), t_at(nc, a->lm_t - i - 1)) ? ':' : ' ');
	PUTS("\n3' ");
	for (int i = prefix; i > 0; --i) PUT(BASE_CHAR[t_at(nc, a->fm_t + i)]);
	for (int k = a->b; k < a->e; ++k) PUT(BASE_CHAR[a->t[k]]);
	for (int i = 1; i <= suffix; ++i) PUT(BASE_CHAR[t_at(nc, a->lm_t - i)]);
	PUTS(" 5'");
	out[n] = '\0';
#undef PUT
#undef PUTS
}

static void fill_out(nc_t *nc, float dp_dg, ref_align_out *out)
{
	memset(out, 0, sizeof(*out));
	const aln_t *a = &nc->best;
	out->tm = a->tm;
	out->dH = a->dH;
	out->dS = a->dS;
	out->dG = a->dH - nc->th->T*a->dS;
	out->dp_dg = dp_dg;
	out->valid = a->valid;
	if (!a->valid) return;
	out->anchor5 = (int32_t)anchor5_query(nc);
	out->anchor3 = (int32_t)anchor3_query(nc);
	out->num_mismatch = (int32_t)count_mismatch(nc);
	out->num_gap = (int32_t)count_gap(nc);
	out->max_poly_degen = (int32_t)max_target_degen(nc);
	out->q_first = a->fm_q;
	out->q_last = a->lm_q;
	out->t_first = a->lm_t; /* alignment_range_target (nuc_cruc_anchor.cpp:386-389) */
	out->t_last = a->fm_t;
	render_alignment(nc, out->alignment, sizeof(out->alignment));
}

static int set_query(nc_t *nc, const char *oligo)
{
	const size_t L = strlen(oligo);
	if (L > MAXSEQ) return fail("set_query: Query size out of bounds");
	for (size_t i = 0; i < L; ++i) {
		const int b = ascii_to_base(oligo[i]);
		if (b < 0) return fail("char_to_nucleic_acid: Illegal base");
		nc->q[i] = (uint8_t)b;
	}
	nc->Lq = (int)L;
	return 0;
}

static __thread nc_t *g_nc;
static __thread thermo_t *g_th;

static nc_t *get_nc(float T, float na)
{
	if (!g_th) g_th = (thermo_t *)calloc(1, sizeof(thermo_t));
	if (g_th->T != T || g_th->na != na) thermo_init(g_th, T, na);
	if (!g_nc) g_nc = (nc_t *)calloc(1, sizeof(nc_t));
	g_nc->th = g_th;
	return g_nc;
}

int orc_align(const char *query, const uint8_t *target, int target_len, float T, float na,
	float strand_conc, int dangle5, int dangle3, ref_align_out *out)
{
	nc_t *nc = get_nc(T, na);
	if (set_query(nc, query) < 0) return -1;
	if (target_len > MAXSEQ) return fail("target too long");
	memcpy(nc->t, target, (size_t)target_len);
	nc->Lt = target_len;
	nc->strand = strand_conc;
	nc->dangle5 = dangle5;
	nc->dangle3 = dangle3;
	float dp_dg;
	if (nc_run(nc, &dp_dg) < 0) return -1;
	fill_out(nc, dp_dg, out);
	return nc->oob ? 1 : 0;
}

/* Oligo-only duplexes the driver attaches to every hit (tntblast_local.cpp:657-686):
 * target == NULL: approximate_tm_homodimer (nuc_cruc.cpp:2457-2516) -- the query against itself
 * (align_homodimer, nuc_cruc.h:718-723), symmetry entropy in the initiation term (:1632);
 * otherwise approximate_tm_heterodimer (:2397-2455) with `target` as the second strand, 5'->3'
 * (set_query / set_target, nuc_cruc.h:947-990).  Ct follows strand(c_a, c_b) (nuc_cruc.h:893-910). */
int orc_dimer(const char *query, const char *target, float T, float na, float conc_a, float conc_b, ref_align_out *out)
{
	nc_t *nc = get_nc(T, na);
	if (set_query(nc, target ? target : query) < 0) return -1;
	memcpy(nc->t, nc->q, (size_t)nc->Lq);
	nc->Lt = nc->Lq;
	if (set_query(nc, query) < 0) return -1;
	nc->strand = (conc_a > conc_b) ? conc_a - 0.5f*conc_b : conc_b - 0.5f*conc_a;
	nc->dangle5 = nc->dangle3 = 0;
	nc->homo = target ? 0 : 1;
	float dp_dg;
	const int rc = nc_run(nc, &dp_dg);
	nc->homo = 0;
	if (rc < 0) return -1;
	fill_out(nc, dp_dg, out);
	return nc->oob ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Hairpins (tntblast_local.cpp:657-686 attaches approximate_tm_hairpin of every oligo to a hit)
 * ---------------------------------------------------------------------------------------- */
/* find_loop_index (nuc_cruc.cpp:2620-2860): the loop with its closing pair, 5 or 6 letters, in the
 * table of special tri- / tetra-loops; letters come from "ACGTE"[base] (anything else never matches) */
static int find_loop_index(const nc_t *nc, int start, int len)
{
	char text[8] = {0};
	for (int k = 0; k < len; ++k) {
		const int b = (start + k >= 0 && start + k < nc->Lq) ? nc->q[start + k] : bGAP;
		text[k] = b <= bT ? "ACGT"[b] : (b <= bE ? 'E' : '?');
	}
	for (int i = 0; i < SL_NUM_HAIRPIN_LOOP; ++i)
		if (strcmp(SL_HAIRPIN_LOOP[i], text) == 0) return i;
	return -1;
}

/* evaluate_hairpin_alignment (nuc_cruc.cpp:2301-2394) */
static int evaluate_hairpin_alignment(nc_t *nc, aln_t *a)
{
	const int last_3 = a->fm_q, last_5 = a->fm_t;
	const unsigned loop_len = (unsigned)(last_3 - last_5 - 1);
	a->dH = 0.0f;
	a->dS = 0.0f;
	if (loop_len > 512) { nc->oob = 1; return 0; }
	a->dS += SL_HAIRPIN_S[loop_len];
	const int last_pair = BBP[q_at(nc, last_5)][q_at(nc, last_3)];
	int idx;
	switch (loop_len) {
	case 3:
		idx = find_loop_index(nc, last_5, 5);
		if (idx >= 0) { a->dH += SL_HAIRPIN_SPECIAL_H[idx]; a->dS += SL_HAIRPIN_SPECIAL_S[idx]; }
		if (last_pair == P_AT || last_pair == P_TA) a->dS += SL_BULGE_AT_CLOSING_S;
		break;
	case 4:
		idx = find_loop_index(nc, last_5, 6);
		if (idx >= 0) { a->dH += SL_HAIRPIN_SPECIAL_H[idx]; a->dS += SL_HAIRPIN_SPECIAL_S[idx]; }
		/* falls through to the terminal mismatch */
	default: {
		const int cur = BBP[q_at(nc, last_5 + 1)][q_at(nc, last_3 - 1)];
		/* param_hairpin_terminal_* are copies of the stacking tables (nuc_cruc_santa_lucia.cpp:594-595) */
		a->dH += SL_PARAM_H[SIDX(last_pair, cur)];
		a->dS += SL_PARAM_S[SIDX(last_pair, cur)];
		break;
	}
	}
	nc->hairpin = 1;
	const int ok = evaluate_alignment(nc, a);
	nc->hairpin = 0;
	return ok;
}

static void hairpin_keep_if_better(nc_t *nc, const aln_t *a, float *best_dg)
{
	const float local_dg = a->dH - nc->th->T*a->dS;
	if (!nc->best.valid || local_dg < *best_dg) {
		nc->best = *a;
		nc->best.valid = 1;
		*best_dg = local_dg;
	}
}

/* enumerate_hairpin_alignments (nuc_cruc.cpp:1172-1407) for one maximal cell */
static int enumerate_hairpin(nc_t *nc, int max_cell)
{
	const int Lq = nc->Lq;
	int first_time = 1, nstack = 0, zero_count = -1;
	unsigned trace_count = 0;
	float best_dg = nc->best.dH - nc->th->T*nc->best.dS;
	branch_t *stack = (branch_t *)malloc(sizeof(branch_t)*(size_t)(6*Lq + 8));
	aln_t *a = (aln_t *)malloc(sizeof(aln_t));
	for (;;) {
		if (!first_time && nstack == 0 && zero_count <= 0) break;
		if (MAX_DP_PATH_ENUM != 0 && MAX_DP_PATH_ENUM < trace_count) break;
		++trace_count;
		first_time = 0;
		aln_clear(a);
		if (trace_back(nc, max_cell, stack, &nstack, &zero_count, a) < 0) { free(stack); free(a); return -1; }
		while (ALN_N(a) > 0 && !WC[BBP[a->q[a->e - 1]][a->t[a->e - 1]]]) {
			if (!IS_VIRTUAL(a->q[a->e - 1])) --a->lm_q;
			if (!IS_VIRTUAL(a->t[a->e - 1])) ++a->lm_t;
			--a->e;
		}
		while (ALN_N(a) > 0 && !WC[BBP[a->q[a->b]][a->t[a->b]]]) {
			if (!IS_VIRTUAL(a->q[a->b])) ++a->fm_q;
			if (!IS_VIRTUAL(a->t[a->b])) --a->fm_t;
			++a->b;
		}
		if (zero_count == 0 && nstack > 0) {
			while (nstack > 0 && !branch_next(&stack[nstack - 1])) --nstack;
			zero_count = -1;
		}
		/* the stem as it is (:1265-1286) */
		if (ALN_N(a) >= 3 && evaluate_hairpin_alignment(nc, a)) hairpin_keep_if_better(nc, a, &best_dg);
		/* one more column at the open end: the next bases, or a dangling-end virtual base (:1307-1326) */
		if (a->lm_t != 0 || a->lm_q != Lq - 1) {
			int tb, qb;
			if (a->lm_t == 0) tb = bE;
			else { --a->lm_t; tb = q_at(nc, a->lm_t); }
			if (a->lm_q == Lq - 1) qb = bE;
			else { ++a->lm_q; qb = q_at(nc, a->lm_q); }
			aln_push_back(a, qb, tb);
		}
		const int align_size = ALN_N(a);
		if (align_size < 3) continue;
		if (evaluate_hairpin_alignment(nc, a)) hairpin_keep_if_better(nc, a, &best_dg);
		/* without the closing pair, unless it is G-C / C-G (:1360-1406) */
		if (align_size <= 3) continue;
		const int last_pair = BBP[q_at(nc, a->fm_t)][q_at(nc, a->fm_q)];
		if (last_pair == bG*7 + bC || last_pair == bC*7 + bG) continue;
		++a->fm_q;
		--a->fm_t;
		++a->b;
		if (evaluate_hairpin_alignment(nc, a)) hairpin_keep_if_better(nc, a, &best_dg);
	}
	free(stack);
	free(a);
	return 0;
}

/* approximate_tm_hairpin (nuc_cruc.cpp:2542-2618) */
int orc_hairpin(const char *query, float T, float na, ref_align_out *out)
{
	nc_t *nc = get_nc(T, na);
	if (set_query(nc, query) < 0) return -1;
	memcpy(nc->t, nc->q, (size_t)nc->Lq);
	nc->Lt = nc->Lq;
	nc->dangle5 = nc->dangle3 = 0;
	nc->homo = 0;
	const int max_stem_len = nc->Lq - 4; /* steric limit: 3 loop bases + 1 anchor (:781-790) */
	nc->oob = 0;
	int32_t max_score = 0;
	if (!g_dinkelbach) {
		nc->tri = max_stem_len > 0 ? max_stem_len : -1;
		aln_clear(&nc->best);
		max_score = align_dimer(nc);
		nc->tri = 0;
		for (int k = 0; k < nc->n_max; ++k)
			if (enumerate_hairpin(nc, nc->max_cells[k]) < 0) return -1;
	}
	else {
		/* the Dinkelbach iteration of approximate_tm_hairpin (nuc_cruc.cpp:2548-2588), like nc_run */
		const thermo_t *keep = nc->th;
		thermo_t *th = (thermo_t *)malloc(sizeof(thermo_t));
		float q = -999999.9f, last_q;
		int rc = 0;
		thermo_init(th, 273.15f, keep->na);
		nc->th = th;
		do {
			nc->tri = max_stem_len > 0 ? max_stem_len : -1;
			aln_clear(&nc->best);
			max_score = align_dimer(nc);
			nc->tri = 0;
			for (int k = 0; k < nc->n_max && rc == 0; ++k)
				if (enumerate_hairpin(nc, nc->max_cells[k]) < 0) rc = -1;
			if (rc < 0) break;
			last_q = q;
			q = nc->best.dH - th->T*nc->best.dS;
			thermo_init(th, 273.15f + nc->best.tm, keep->na);
		} while (q < 0.0 && q > last_q);
		nc->th = keep;
		free(th);
		if (rc < 0) return -1;
	}
	memset(out, 0, sizeof(*out));
	const aln_t *a = &nc->best;
	out->tm = a->tm;
	out->valid = a->valid;
	out->dH = a->dH;
	out->dS = a->dS;
	out->dG = a->dH - T*a->dS;
	out->dp_dg = -((float)max_score/10000.0f);
	if (a->valid) {
		out->q_first = a->fm_q; out->t_first = a->fm_t;
		out->q_last = a->lm_q; out->t_last = a->lm_t;
		out->num_gap = ALN_N(a);
	}
	return nc->oob ? 1 : 0;
}

/* Window extraction (bind_oligo.cpp:502-592 minus strand, :1205-1295 plus strand) */
static void load_window(nc_t *nc, const uint8_t *codes, uint32_t len, int plus_strand,
	uint32_t query_loc, uint32_t target_loc, unsigned *start_out, unsigned *stop_out)
{
	const unsigned window = (unsigned)nc->Lq;
	const unsigned target_length = window + 2*NUM_FLANK;
	int s = (int)target_loc - (int)(query_loc + NUM_FLANK);
	const unsigned start = s > 0 ? (unsigned)s : 0u;
	unsigned stop = start + target_length;
	if (stop > len) stop = len;

	int n = 0;
	uint8_t tmp[MAXSEQ];
	for (unsigned i = start; i < stop && n < MAXSEQ; ++i) {
		const uint8_t c = codes[i];
		if (c > 15) continue;
		tmp[n++] = (uint8_t)(plus_strand ? DB_TO_BASE[c] : DB_TO_COMP[c]);
	}
	if (plus_strand) memcpy(nc->t, tmp, (size_t)n);
	else for (int i = 0; i < n; ++i) nc->t[i] = tmp[n - 1 - i]; /* push_front == reversed */
	nc->Lt = n;
	*start_out = start;
	*stop_out = stop;
}

static void map_coords(const nc_t *nc, int plus_strand, unsigned start, unsigned stop, int *loc5, int *loc3)
{
	const aln_t *a = &nc->best;
	const int window = nc->Lq;
	const int q_first = a->fm_q, q_last = a->lm_q, t_first = a->lm_t, t_last = a->fm_t;
	int t5 = (int)start, t3 = (int)start;
	if (plus_strand) { /* bind_oligo.cpp:1424-1434 */
		t5 += t_first;
		t3 += t_last;
		t3 += q_first;
		t5 -= (window - 1) - q_last;
	}
	else { /* bind_oligo.cpp:721-731 */
		t5 += (int)(stop - start) - 1 - t_last;
		t3 += (int)(stop - start) - 1 - t_first;
		t5 -= q_first;
		t3 += (window - 1) - q_last;
	}
	*loc5 = t5;
	*loc3 = t3;
}

int orc_bind_window(const uint8_t *codes, uint32_t len, const char *oligo, int plus_strand,
	uint32_t query_loc, uint32_t target_loc, float T, float na, float strand_conc,
	int dangle5, int dangle3, ref_align_out *out)
{
	nc_t *nc = get_nc(T, na);
	if (set_query(nc, oligo) < 0) return -1;
	nc->strand = strand_conc;
	nc->dangle5 = dangle5;
	nc->dangle3 = dangle3;
	unsigned start, stop;
	load_window(nc, codes, len, plus_strand, query_loc, target_loc, &start, &stop);
	float dp_dg;
	if (nc_run(nc, &dp_dg) < 0) return -1;
	fill_out(nc, dp_dg, out);
	out->target_start = (int32_t)start;
	out->target_stop = (int32_t)stop;
	if (nc->best.valid) map_coords(nc, plus_strand, start, stop, &out->loc_5, &out->loc_3);
	return nc->oob ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Seeds (seq_hash.h)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	int W;
	uint32_t nwords;   /* 4^W */
	uint32_t *first;   /* [nwords+1] bucket starts */
	uint32_t *index;   /* start positions grouped by word, ascending inside a bucket */
} khash_t;

/* DNAHash::hash<SEQPTR> (seq_hash.h:524-642).  `b = code & 3` makes every target base
 * "valid" (b <= DB_MAX_ATGC is always true), so non-ACGT codes alias to code&3 and a window
 * is indexed at every position >= W-1. */
static void khash_build(khash_t *h, const uint8_t *codes, uint32_t len, int W)
{
	h->W = W;
	h->nwords = 1u << (2*W);
	const uint32_t mask = h->nwords - 1;
	h->first = (uint32_t *)calloc((size_t)h->nwords + 1, sizeof(uint32_t));
	uint32_t total = 0;
	uint32_t word = 0;
	for (uint32_t i = 0; i < len; ++i) {
		word = ((word << 2) | (codes[i] & 3u)) & 0xffffu; /* unsigned short accumulator */
		if (i + 1 >= (uint32_t)W) { h->first[(word & mask) + 1]++; total++; }
	}
	for (uint32_t w = 0; w < h->nwords; ++w) h->first[w + 1] += h->first[w];
	h->index = (uint32_t *)malloc(sizeof(uint32_t)*(total ? total : 1));
	uint32_t *fill = (uint32_t *)malloc(sizeof(uint32_t)*h->nwords);
	memcpy(fill, h->first, sizeof(uint32_t)*h->nwords);
	word = 0;
	for (uint32_t i = 0; i < len; ++i) {
		word = ((word << 2) | (codes[i] & 3u)) & 0xffffu;
		if (i + 1 >= (uint32_t)W) h->index[fill[word & mask]++] = i + 1 - (uint32_t)W;
	}
	free(fill);
}

static void khash_free(khash_t *h)
{
	free(h->first);
	free(h->index);
	h->first = h->index = NULL;
}

/* DNAHash_iterator::build_word_list<std::string> (seq_hash.h:287-374).  Non-ACGT oligo
 * letters restart the run but do not shift `word`; the list is compacted, so the index of a
 * word in the list (what offset() reports) is NOT its base offset once a letter was skipped. */
static int build_word_list(const char *oligo, int W, int complement, uint16_t *words)
{
	const int L = (int)strlen(oligo);
	if (W > L) return 0;
	const uint16_t mask = (uint16_t)((1u << (2*W)) - 1);
	uint16_t word = 0;
	int run = 0, n = 0;
	for (int k = 0; k < L; ++k) {
		const char c = complement ? oligo[L - 1 - k] : oligo[k];
		++run;
		int b = -1;
		switch (c) {
		case 'A': case 'a': b = 0; break;
		case 'C': case 'c': b = 1; break;
		case 'G': case 'g': b = 2; break;
		case 'T': case 't': b = 3; break;
		default: run = 0; break;
		}
		if (b >= 0) word = (uint16_t)((word << 2) | (complement ? 3 - b : b));
		if (run >= W) words[n++] = word & mask;
	}
	return n;
}

typedef struct { uint32_t q, t; } seed_t;

/* iteration of DNAHash::find / find_complement: by word-list index, then ascending position */
static long enumerate_seeds(const khash_t *h, const char *oligo, int complement, seed_t **out)
{
	uint16_t words[MAXSEQ];
	const int nw = build_word_list(oligo, h->W, complement, words);
	long n = 0, cap = 256;
	seed_t *s = (seed_t *)malloc(sizeof(seed_t)*cap);
	for (int k = 0; k < nw; ++k) {
		for (uint32_t p = h->first[words[k]]; p < h->first[words[k] + 1]; ++p) {
			if (n == cap) { cap *= 2; s = (seed_t *)realloc(s, sizeof(seed_t)*cap); }
			s[n].q = (uint32_t)k;
			s[n].t = h->index[p];
			++n;
		}
	}
	*out = s;
	return n;
}

/* stable sort by diagonal (q - t as int) + unique: one seed per diagonal, the first in
 * iteration order (bind_oligo.cpp:98-99, :157-158) */
static int seed_diag(const seed_t *s) { return (int)s->q - (int)s->t; }

static void seed_merge_sort(seed_t *a, seed_t *tmp, long n)
{
	if (n < 2) return;
	const long h = n/2;
	seed_merge_sort(a, tmp, h);
	seed_merge_sort(a + h, tmp, n - h);
	long i = 0, j = h, k = 0;
	while (i < h && j < n) tmp[k++] = (seed_diag(&a[j]) < seed_diag(&a[i])) ? a[j++] : a[i++];
	while (i < h) tmp[k++] = a[i++];
	while (j < n) tmp[k++] = a[j++];
	memcpy(a, tmp, sizeof(seed_t)*(size_t)n);
}

static long unique_seeds(seed_t *s, long n)
{
	if (n == 0) return 0;
	seed_t *tmp = (seed_t *)malloc(sizeof(seed_t)*(size_t)n);
	seed_merge_sort(s, tmp, n);
	free(tmp);
	long m = 1;
	for (long i = 1; i < n; ++i)
		if (seed_diag(&s[i]) != seed_diag(&s[m - 1])) s[m++] = s[i];
	return m;
}

long orc_seeds_raw(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int complement, uint32_t *q_out, uint32_t *t_out, long cap)
{
	if (word_size < 2 || word_size > 8) return fail("DNAHash: Unsupported word length");
	khash_t h;
	khash_build(&h, codes, len, word_size);
	seed_t *s;
	const long n = enumerate_seeds(&h, oligo, complement, &s);
	for (long i = 0; i < n && i < cap; ++i) { q_out[i] = s[i].q; t_out[i] = s[i].t; }
	free(s);
	khash_free(&h);
	return n;
}

long orc_seeds_unique(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int plus_strand, uint32_t *q_out, uint32_t *t_out, long cap)
{
	if (word_size < 2 || word_size > 8) return fail("DNAHash: Unsupported word length");
	khash_t h;
	khash_build(&h, codes, len, word_size);
	seed_t *s;
	long n = enumerate_seeds(&h, oligo, plus_strand, &s);
	n = unique_seeds(s, n);
	for (long i = 0; i < n && i < cap; ++i) { q_out[i] = s[i].q; t_out[i] = s[i].t; }
	free(s);
	khash_free(&h);
	return n;
}

/* ------------------------------------------------------------------------------------------
 * Assay search: bind_oligo.cpp, amplicon_search.cpp, probe_search.cpp, padlock_search.cpp
 * ---------------------------------------------------------------------------------------- */
enum { M_F = 1, M_R = 2, M_P = 4, M_PLUS = 8, M_MINUS = 16, M_VALID = 32 }; /* tntblast.h:147-154 */

typedef struct {
	int loc_5, loc_3;
	float tm, dH, dS;
	unsigned anchor_5, anchor_3, num_mm, num_gap;
	char *alignment; /* owned by the string pool */
	unsigned query_loc, target_loc;
	unsigned char mask;
} oinfo_t;

typedef struct { oinfo_t *v; long n, cap; } olist_t;

static void ol_push(olist_t *l, const oinfo_t *e)
{
	if (l->n == l->cap) { l->cap = l->cap ? 2*l->cap : 64; l->v = (oinfo_t *)realloc(l->v, sizeof(oinfo_t)*(size_t)l->cap); }
	l->v[l->n++] = *e;
}

/* string pool: alignment strings live until the end of the search call */
typedef struct spool { struct spool *next; char s[1]; } spool_t;
static __thread spool_t *g_pool;
static char *pool_str(const char *s)
{
	const size_t n = strlen(s);
	spool_t *p = (spool_t *)malloc(sizeof(spool_t) + n);
	memcpy(p->s, s, n + 1);
	p->next = g_pool;
	g_pool = p;
	return p->s;
}
static void pool_free(void)
{
	while (g_pool) { spool_t *n = g_pool->next; free(g_pool); g_pool = n; }
}
static const char EMPTY[] = "";

/* std::list::sort of libstdc++ (bits/list.tcc): bottom-up merge with 64 bins.  Restated because
 * sort_by_oligo_loc (amplicon_search.cpp:12-26) is not a strict weak ordering once bound and
 * unbound elements are mixed, so the resulting order depends on the exact merge sequence. */
typedef int (*ocmp_t)(const oinfo_t *, const oinfo_t *);

static void ol_merge(olist_t *a, olist_t *b, ocmp_t less) /* a.merge(b): result in a, b emptied */
{
	olist_t r = {0};
	r.cap = a->n + b->n + 1;
	r.v = (oinfo_t *)malloc(sizeof(oinfo_t)*(size_t)r.cap);
	long i = 0, j = 0;
	while (i < a->n && j < b->n) {
		if (less(&b->v[j], &a->v[i])) r.v[r.n++] = b->v[j++];
		else r.v[r.n++] = a->v[i++];
	}
	while (i < a->n) r.v[r.n++] = a->v[i++];
	while (j < b->n) r.v[r.n++] = b->v[j++];
	free(a->v);
	free(b->v);
	*a = r;
	b->v = NULL; b->n = b->cap = 0;
}

static void ol_swap(olist_t *a, olist_t *b) { olist_t t = *a; *a = *b; *b = t; }

static void ol_sort(olist_t *l, ocmp_t less)
{
	if (l->n < 2) return;
	olist_t carry = {0}, bins[64];
	memset(bins, 0, sizeof(bins));
	int fill = 0;
	for (long k = 0; k < l->n; ++k) {
		ol_push(&carry, &l->v[k]);
		int counter;
		for (counter = 0; counter != fill && bins[counter].n != 0; ++counter) {
			ol_merge(&bins[counter], &carry, less);
			ol_swap(&carry, &bins[counter]);
		}
		ol_swap(&carry, &bins[counter]);
		if (counter == fill) ++fill;
	}
	for (int counter = 1; counter != fill; ++counter) ol_merge(&bins[counter], &bins[counter - 1], less);
	free(l->v);
	*l = bins[fill - 1];
	for (int k = 0; k < fill - 1; ++k) free(bins[k].v);
	free(carry.v);
}

static int less_hash_match(const oinfo_t *a, const oinfo_t *b) /* bind_oligo.cpp:33-39 */
{
	return ((int)a->query_loc - (int)a->target_loc) < ((int)b->query_loc - (int)b->target_loc);
}

static int less_oligo_loc(const oinfo_t *a, const oinfo_t *b) /* amplicon_search.cpp:12-26 */
{
	if (!(a->loc_5 + a->loc_3) || !(b->loc_5 + b->loc_3)) return a->target_loc < b->target_loc;
	if (a->loc_5 == b->loc_5) return a->loc_3 < b->loc_3;
	return a->loc_5 < b->loc_5;
}

static int less_bound_match(const oinfo_t *a, const oinfo_t *b) /* bind_oligo.cpp:49-82 */
{
	if (a->loc_5 != b->loc_5) return a->loc_5 < b->loc_5;
	if (a->loc_3 != b->loc_3) return a->loc_3 < b->loc_3;
	if (a->tm == b->tm) {
		if (a->num_mm == b->num_mm) return strlen(a->alignment) > strlen(b->alignment);
		return a->num_mm > b->num_mm;
	}
	return a->tm > b->tm;
}

static int less_oinfo(const oinfo_t *a, const oinfo_t *b) /* oligo_info::operator< tntblast.h:230-242 */
{
	if (a->loc_5 != b->loc_5) return a->loc_5 < b->loc_5;
	if (a->loc_3 != b->loc_3) return a->loc_3 < b->loc_3;
	return a->tm > b->tm;
}

/* melt cache (tntblast.h:247-324): keyed by (oligo string, window start, window stop) */
typedef struct {
	const char *oligo;
	unsigned start, stop;
	float tm, dg, dH, dS;
	unsigned anchor_5, anchor_3, num_mm, num_gap, poly_degen;
	int target_5, target_3;
	char *align;
} centry_t;

typedef struct { centry_t *v; long n, cap; long *slots; long nslots; } cache_t;

static uint64_t cache_hash(const char *oligo, unsigned start, unsigned stop)
{
	uint64_t h = 1469598103934665603ULL;
	for (const char *p = oligo; *p; ++p) { h ^= (unsigned char)*p; h *= 1099511628211ULL; }
	h ^= ((uint64_t)start << 32) | stop;
	h *= 0x9E3779B97F4A7C15ULL;
	return h;
}

static void cache_rehash(cache_t *c, long nslots)
{
	free(c->slots);
	c->nslots = nslots;
	c->slots = (long *)malloc(sizeof(long)*(size_t)nslots);
	for (long i = 0; i < nslots; ++i) c->slots[i] = -1;
	for (long k = 0; k < c->n; ++k) {
		uint64_t s = cache_hash(c->v[k].oligo, c->v[k].start, c->v[k].stop)%(uint64_t)nslots;
		while (c->slots[s] >= 0) s = (s + 1)%(uint64_t)nslots;
		c->slots[s] = k;
	}
}

static centry_t *cache_find(cache_t *c, const char *oligo, unsigned start, unsigned stop)
{
	if (!c->nslots) return NULL;
	uint64_t s = cache_hash(oligo, start, stop)%(uint64_t)c->nslots;
	while (c->slots[s] >= 0) {
		centry_t *e = &c->v[c->slots[s]];
		if (e->start == start && e->stop == stop && strcmp(e->oligo, oligo) == 0) return e;
		s = (s + 1)%(uint64_t)c->nslots;
	}
	return NULL;
}

static void cache_put(cache_t *c, const centry_t *e)
{
	if (c->n == c->cap) { c->cap = c->cap ? 2*c->cap : 256; c->v = (centry_t *)realloc(c->v, sizeof(centry_t)*(size_t)c->cap); }
	c->v[c->n++] = *e;
	if (c->n*2 > c->nslots) cache_rehash(c, c->nslots ? c->nslots*4 : 1024);
	else {
		uint64_t s = cache_hash(e->oligo, e->start, e->stop)%(uint64_t)c->nslots;
		while (c->slots[s] >= 0) s = (s + 1)%(uint64_t)c->nslots;
		c->slots[s] = c->n - 1;
	}
}

static void cache_free(cache_t *c) { free(c->v); free(c->slots); memset(c, 0, sizeof(*c)); }

typedef struct {
	float min_tm, max_tm, min_dg, max_dg;
	unsigned clamp_5, clamp_3, max_gap, max_mismatch, max_poly_degen;
} bind_limits_t;

/* Evaluate one seed: window, NucCruc, filters in the reference's order, coordinates, text.
 * Returns 1 when the oligo binds (fields of `e` filled in), 0 otherwise.
 * (bind_oligo.cpp:502-803 / :1205-1506, including the partial-result cache semantics) */
static int bind_seed(nc_t *nc, cache_t *cache, const uint8_t *codes, uint32_t len,
	const char *oligo, int plus_strand, const bind_limits_t *lim, oinfo_t *e)
{
	const unsigned window = (unsigned)nc->Lq;
	int s = (int)e->target_loc - (int)(e->query_loc + NUM_FLANK);
	const unsigned start = s > 0 ? (unsigned)s : 0u;
	unsigned stop = start + window + 2*NUM_FLANK;
	if (stop > len) stop = len;

	centry_t *hit = cache ? cache_find(cache, oligo, start, stop) : NULL;
	centry_t ce;
	if (!hit) {
		memset(&ce, 0, sizeof(ce));
		ce.oligo = oligo;
		ce.start = start;
		ce.stop = stop;
		ce.align = (char *)EMPTY;
		unsigned s2, e2;
		load_window(nc, codes, len, plus_strand, e->query_loc, e->target_loc, &s2, &e2);
		nc_run(nc, NULL);
		const aln_t *a = &nc->best;
		int ok = 1;
		ce.tm = a->tm;
		if (ce.tm < lim->min_tm || ce.tm > lim->max_tm) ok = 0;
		if (ok) {
			ce.dg = a->dH - nc->th->T*a->dS;
			if (ce.dg < lim->min_dg || ce.dg > lim->max_dg) ok = 0;
		}
		if (ok) { ce.anchor_5 = anchor5_query(nc); if (ce.anchor_5 < lim->clamp_5) ok = 0; }
		if (ok) { ce.anchor_3 = anchor3_query(nc); if (ce.anchor_3 < lim->clamp_3) ok = 0; }
		if (ok) { ce.num_mm = count_mismatch(nc); if (ce.num_mm > lim->max_mismatch) ok = 0; }
		if (ok) { ce.num_gap = count_gap(nc); if (ce.num_gap > lim->max_gap) ok = 0; }
		if (ok) { ce.poly_degen = max_target_degen(nc); if (ce.poly_degen > lim->max_poly_degen) ok = 0; }
		if (ok) {
			map_coords(nc, plus_strand, start, stop, &ce.target_5, &ce.target_3);
			char buf[4*MAXSEQ];
			render_alignment(nc, buf, sizeof(buf));
			ce.align = pool_str(buf);
			ce.dH = a->dH;
			ce.dS = a->dS;
		}
		if (cache) cache_put(cache, &ce);
		if (!ok) return 0;
		hit = &ce;
	}
	else {
		/* cache hit: the *current* limits are re-applied to the stored (possibly partial) record */
		if (hit->tm < lim->min_tm || hit->tm > lim->max_tm) return 0;
		if (hit->dg < lim->min_dg || hit->dg > lim->max_dg) return 0;
		if (hit->anchor_5 < lim->clamp_5 || hit->anchor_3 < lim->clamp_3) return 0;
		if (hit->num_mm > lim->max_mismatch) return 0;
		if (hit->num_gap > lim->max_gap) return 0;
		if (hit->poly_degen > lim->max_poly_degen) return 0;
	}
	e->loc_5 = hit->target_5;
	e->loc_3 = hit->target_3;
	e->tm = hit->tm;
	e->dH = hit->dH;
	e->dS = hit->dS;
	e->anchor_5 = hit->anchor_5;
	e->anchor_3 = hit->anchor_3;
	e->num_mm = hit->num_mm;
	e->num_gap = hit->num_gap;
	e->alignment = hit->align;
	return 1;
}

typedef struct {
	const uint8_t *codes;
	uint32_t len;
	khash_t hash;
	nc_t *nc;
	cache_t plus_cache, minus_cache;
} search_ctx_t;

/* match_oligo_to_{minus,plus}_strand (bind_oligo.cpp:84-122).  list::merge with
 * oligo_info::operator< on all-unbound elements degenerates to an append. */
static void match_oligo(search_ctx_t *cx, olist_t *list, const char *oligo, int plus, unsigned char mask)
{
	seed_t *s;
	long n = enumerate_seeds(&cx->hash, oligo, plus, &s);
	n = unique_seeds(s, n);
	for (long i = 0; i < n; ++i) {
		oinfo_t e;
		memset(&e, 0, sizeof(e));
		e.query_loc = s[i].q;
		e.target_loc = s[i].t;
		e.mask = (unsigned char)(mask | (plus ? M_PLUS : M_MINUS));
		e.tm = e.dH = e.dS = -1.0f;
		e.alignment = (char *)EMPTY;
		ol_push(list, &e);
	}
	free(s);
}

/* mask-variant bind (bind_oligo.cpp:456-827, :1159-1530) */
static void bind_masked(search_ctx_t *cx, olist_t *list, unsigned char oligo_mask, const char *oligo,
	int plus, float strand, const bind_limits_t *lim)
{
	nc_t *nc = cx->nc;
	set_query(nc, oligo);
	nc->strand = strand;
	const unsigned char want = (unsigned char)(oligo_mask | (plus ? M_PLUS : M_MINUS));
	olist_t keep = {0}, cur = {0};
	for (long k = 0; k < list->n; ++k) {
		oinfo_t e = list->v[k];
		if ((e.mask & want) != want) { ol_push(&keep, &e); continue; }
		if (bind_seed(nc, plus ? &cx->plus_cache : &cx->minus_cache, cx->codes, cx->len, oligo, plus, lim, &e))
			ol_push(&cur, &e);
	}
	/* curr_oligo was built with push_front: reverse before the stable sort */
	for (long i = 0, j = cur.n - 1; i < j; ++i, --j) { oinfo_t t = cur.v[i]; cur.v[i] = cur.v[j]; cur.v[j] = t; }
	ol_sort(&cur, less_bound_match);
	for (long k = 0; k < cur.n; ++k) {
		if (k == 0) { ol_push(&keep, &cur.v[k]); continue; }
		const oinfo_t *back = &keep.v[keep.n - 1];
		if (back->loc_5 != cur.v[k].loc_5 || back->loc_3 != cur.v[k].loc_3) ol_push(&keep, &cur.v[k]);
	}
	free(cur.v);
	free(list->v);
	*list = keep;
}

/* hash-variant bind (bind_oligo.cpp:124-454, :829-1157) */
static void bind_hashed(search_ctx_t *cx, olist_t *out, const char *oligo, int plus, float strand,
	const bind_limits_t *lim, int use_cache)
{
	nc_t *nc = cx->nc;
	set_query(nc, oligo);
	nc->strand = strand;
	seed_t *s;
	long n = enumerate_seeds(&cx->hash, oligo, plus, &s);
	n = unique_seeds(s, n);
	olist_t hits = {0};
	for (long i = 0; i < n; ++i) {
		oinfo_t e;
		memset(&e, 0, sizeof(e));
		e.query_loc = s[i].q;
		e.target_loc = s[i].t;
		e.alignment = (char *)EMPTY;
		cache_t *c = use_cache ? (plus ? &cx->plus_cache : &cx->minus_cache) : NULL;
		if (bind_seed(nc, c, cx->codes, cx->len, oligo, plus, lim, &e)) {
			e.query_loc = e.target_loc = 0; /* oligo_info bound-constructor, tntblast.h:156-170 */
			e.mask = 0;
			ol_push(&hits, &e);
		}
	}
	free(s);
	out->n = 0;
	ol_sort(&hits, less_oinfo);
	for (long k = 0; k < hits.n; ++k) {
		if (k == 0 || out->v[out->n - 1].loc_5 != hits.v[k].loc_5 || out->v[out->n - 1].loc_3 != hits.v[k].loc_3)
			ol_push(out, &hits.v[k]);
	}
	free(hits.v);
}

/* cull_oligo_match (amplicon_search.cpp:679-765) */
static void cull(olist_t *l, unsigned max_amplicon_len, int has_probe, int single_primer_pcr,
	unsigned *n_minus, unsigned *n_plus)
{
	const unsigned threshold = max_amplicon_len + 50;
	ol_sort(l, less_oligo_loc);
	for (long i = 0; i < l->n; ++i) l->v[i].mask &= (unsigned char)~M_VALID;
	for (long f = 0; f < l->n; ++f) {
		if (l->v[f].mask & (M_PLUS | M_P)) continue;
		for (long r = f + 1; r < l->n; ++r) {
			if ((unsigned)(l->v[r].target_loc - l->v[f].target_loc) > threshold) break; /* unsigned wrap kept */
			if (l->v[r].mask & (M_MINUS | M_P)) continue;
			if (!single_primer_pcr && ((l->v[f].mask & (M_R | M_F)) == (l->v[r].mask & (M_R | M_F)))) continue;
			if (has_probe) {
				for (long p = f + 1; p < r; ++p)
					if (l->v[p].mask & M_P) {
						l->v[p].mask |= M_VALID;
						l->v[f].mask |= M_VALID;
						l->v[r].mask |= M_VALID;
					}
			}
			else { l->v[f].mask |= M_VALID; l->v[r].mask |= M_VALID; }
		}
	}
	/* The reference counts strands on the element *after* each kept one (:748-753, reads
	 * end() for the last).  The counts only choose the binding order; the sentinel is read as 0. */
	unsigned cm = 0, cp = 0;
	long m = 0;
	for (long i = 0; i < l->n; ++i) {
		if (l->v[i].mask & M_VALID) {
			const unsigned char next = (i + 1 < l->n) ? l->v[i + 1].mask : 0;
			cm += (next & M_MINUS) ? 1 : 0;
			cp += (next & M_PLUS) ? 1 : 0;
			l->v[m++] = l->v[i];
		}
	}
	l->n = m;
	if (n_minus) *n_minus = cm;
	if (n_plus) *n_plus = cp;
}

static __thread ref_hit *g_hits;
static __thread long g_nhits, g_caphits;

static ref_hit *new_hit(void)
{
	if (g_nhits == g_caphits) {
		g_caphits = g_caphits ? 2*g_caphits : 64;
		g_hits = (ref_hit *)realloc(g_hits, sizeof(ref_hit)*(size_t)g_caphits);
	}
	ref_hit *h = &g_hits[g_nhits++];
	memset(h, 0, sizeof(*h));
	/* hybrid_sig::init() (hybrid_sig.h:52-107) */
	h->forward_tm = h->reverse_tm = h->probe_tm = -1.0f;
	h->forward_dH = h->reverse_dH = h->probe_dH = 100.0f;
	h->forward_mm = h->forward_gap = h->reverse_mm = h->reverse_gap = h->probe_mm = h->probe_gap = -1;
	h->forward_clamp = h->reverse_clamp = -1;
	return h;
}

static void set_str(char *dst, size_t cap, const char *s)
{
	strncpy(dst, s, cap - 1);
	dst[cap - 1] = '\0';
}

static const char DB_ASCII[] = "ACGTIMRSVWYHKDBN-";      /* hash_base_to_ascii, seq.h:58-101 */
static const char DB_ASCII_COMP[] = "TGCAIKYSBWRDMHVN-"; /* hash_base_to_ascii_complement, seq.h:103-146 */

static void set_amplicon(ref_hit *h, const char *amp, size_t n)
{
	h->amplicon_len = (int32_t)n;
	uint64_t hash = 1469598103934665603ULL;
	for (size_t k = 0; k < n; ++k) { hash ^= (unsigned char)amp[k]; hash *= 1099511628211ULL; }
	h->amplicon_fnv = hash;
	const size_t m = n < sizeof(h->amplicon_head) - 1 ? n : sizeof(h->amplicon_head) - 1;
	memcpy(h->amplicon_head, amp, m);
	h->amplicon_head[m] = '\0';
}

/* amplicon text, forward orientation (amplicon_search.cpp:508-523) */
static void amplicon_plus(const search_ctx_t *cx, ref_hit *h, int start, int stop, int first_i)
{
	const int n = stop - start + 1;
	char *amp = (char *)malloc((size_t)n + 1);
	memset(amp, '-', (size_t)n);
	long p = start > 0 ? start : 0;
	for (int i = first_i; i < n; ++i, ++p) {
		if (p >= (long)cx->len) break;
		amp[i] = DB_ASCII[cx->codes[p] > 16 ? 16 : cx->codes[p]];
	}
	set_amplicon(h, amp, (size_t)n);
	free(amp);
}

/* amplicon text, complemented and reversed (amplicon_search.cpp:524-537) */
static void amplicon_minus(const search_ctx_t *cx, ref_hit *h, int start, int stop, int first_i)
{
	const int n = stop - start + 1;
	char *amp = (char *)malloc((size_t)n + 1);
	memset(amp, '-', (size_t)n);
	long p = stop < (int)cx->len - 1 ? stop : (long)cx->len - 1;
	for (int i = first_i; i < n; ++i, --p) {
		if (p < 0) break;
		amp[i] = DB_ASCII_COMP[cx->codes[p] > 16 ? 16 : cx->codes[p]];
	}
	set_amplicon(h, amp, (size_t)n);
	free(amp);
}

typedef struct {
	const char *F, *R, *P;
	int fdeg, rdeg, pdeg;
} assay_t;

/* amplicon() (amplicon_search.cpp:58-677) */
static void search_pcr(search_ctx_t *cx, const assay_t *as, const ref_options *o)
{
	const int has_probe = as->P && as->P[0];
	const int apply_mmc = o->min_max_primer_clamp >= 0;
	const unsigned mmc = apply_mmc ? (unsigned)o->min_max_primer_clamp : 0;
	const float fs = o->forward_primer_strand/as->fdeg;
	const float rs = o->reverse_primer_strand/as->rdeg;
	const float ps = o->probe_strand/as->pdeg;
	/* strand(c, 0) => Ct = c - 0.5*0 (nuc_cruc.h:890-910) */
	const float f_ct = fs - 0.5f*0.0f, r_ct = rs - 0.5f*0.0f, p_ct = ps - 0.5f*0.0f;

	olist_t ml = {0};
	match_oligo(cx, &ml, as->F, 0, M_F);
	match_oligo(cx, &ml, as->R, 0, M_R);
	const long n_minus = ml.n;
	if (n_minus == 0) { free(ml.v); return; }
	match_oligo(cx, &ml, as->F, 1, M_F);
	match_oligo(cx, &ml, as->R, 1, M_R);
	const long n_plus = ml.n;
	if (n_plus == n_minus) { free(ml.v); return; }
	if (has_probe) {
		match_oligo(cx, &ml, as->P, 0, M_P);
		match_oligo(cx, &ml, as->P, 1, M_P);
		if (ml.n == n_plus) { free(ml.v); return; }
	}

	unsigned cm, cp;
	cull(&ml, o->max_len, has_probe, o->single_primer_pcr, &cm, &cp);

	bind_limits_t pl = {o->min_primer_tm, o->max_primer_tm, o->min_primer_dg, o->max_primer_dg,
		0, o->primer_clamp, o->max_gap, o->max_mismatch, o->max_poly_degen};
	bind_limits_t bl = {o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
		o->probe_clamp_5, o->probe_clamp_3, o->max_gap, o->max_mismatch, o->max_poly_degen};

	const int first_plus = !(cm < cp); /* :131 vs :218 */
	for (int stage = 0; stage < 4; ++stage) {
		const int plus = (stage < 2) ? first_plus : !first_plus;
		const int is_r = stage & 1;
		bind_masked(cx, &ml, is_r ? M_R : M_F, is_r ? as->R : as->F, plus, is_r ? r_ct : f_ct, &pl);
		if (stage < 3) {
			cull(&ml, o->max_len, has_probe, o->single_primer_pcr, NULL, NULL);
			/* early exits at :153, :177, :240, :264, :288 -- not after the third bind on the
			 * minus-first path (:199), where the code simply carries on */
			if (ml.n == 0 && !(stage == 2 && !first_plus)) { free(ml.v); return; }
		}
	}

	if (has_probe) {
		cull(&ml, o->max_len, has_probe, o->single_primer_pcr, NULL, NULL);
		if (ml.n == 0) { free(ml.v); return; }
		bind_masked(cx, &ml, M_P, as->P, 0, p_ct, &bl);
		bind_masked(cx, &ml, M_P, as->P, 1, p_ct, &bl);
	}

	ol_sort(&ml, less_oligo_loc);

	for (long f = 0; f < ml.n; ++f) {
		const oinfo_t *F = &ml.v[f];
		if (F->mask & (M_PLUS | M_P)) continue;
		for (long r = f + 1; r < ml.n; ++r) {
			const oinfo_t *R = &ml.v[r];
			if (R->mask & (M_MINUS | M_P)) continue;
			if (!o->single_primer_pcr && ((F->mask & (M_R | M_F)) == (R->mask & (M_R | M_F)))) continue;
			if (F->loc_3 >= R->loc_5) continue;
			if ((R->loc_3 - F->loc_5 + 1) > (int)o->max_len) continue;
			if (apply_mmc && ((F->anchor_3 > R->anchor_3 ? F->anchor_3 : R->anchor_3) <= mmc)) continue;

			const int amp_start = F->loc_5, amp_stop = R->loc_3;
			const int plus_primer = (F->mask & M_F) != 0;

			for (long p = has_probe ? f + 1 : r; p <= r; ++p) {
				const oinfo_t *Pp = NULL;
				if (has_probe) {
					if (p == r) break;
					Pp = &ml.v[p];
					if (!(Pp->mask & M_P)) continue;
					if (!(Pp->loc_5 >= amp_start && Pp->loc_3 <= amp_stop)) continue;
					if ((Pp->mask & (M_PLUS | M_MINUS)) == (F->mask & (M_PLUS | M_MINUS))) {
						if (Pp->loc_5 <= F->loc_3) continue;
					}
					else if (Pp->loc_3 >= R->loc_5) continue;
				}
				ref_hit *h = new_hit();
				const char *fo = as->F, *ro = as->R;
				if ((F->mask & M_R) && (R->mask & M_R)) fo = as->R;
				if ((F->mask & M_F) && (R->mask & M_F)) ro = as->F;
				set_str(h->forward_oligo, sizeof(h->forward_oligo), fo);
				set_str(h->reverse_oligo, sizeof(h->reverse_oligo), ro);
				h->primer_strand = plus_primer ? 0 : 1;
				h->amp_first = amp_start;
				h->amp_last = amp_stop;
				const oinfo_t *fo_i = F, *ro_i = R;
				if ((F->mask & M_R) && (R->mask & M_F)) { fo_i = R; ro_i = F; }
				h->forward_tm = fo_i->tm; h->forward_dH = fo_i->dH; h->forward_dS = fo_i->dS;
				h->reverse_tm = ro_i->tm; h->reverse_dH = ro_i->dH; h->reverse_dS = ro_i->dS;
				h->forward_mm = (int8_t)fo_i->num_mm; h->reverse_mm = (int8_t)ro_i->num_mm;
				h->forward_gap = (int8_t)fo_i->num_gap; h->reverse_gap = (int8_t)ro_i->num_gap;
				h->forward_clamp = (int8_t)fo_i->anchor_3;
				h->reverse_clamp = (int8_t)ro_i->anchor_3;
				set_str(h->forward_align, sizeof(h->forward_align), fo_i->alignment);
				set_str(h->reverse_align, sizeof(h->reverse_align), ro_i->alignment);
				if (plus_primer) amplicon_plus(cx, h, amp_start, amp_stop, amp_start < 0 ? -amp_start : 0);
				else {
					const int skip = amp_stop - (int)cx->len + 1;
					amplicon_minus(cx, h, amp_start, amp_stop, skip > 0 ? skip : 0);
				}
				if (Pp) {
					h->probe_first = Pp->loc_5;
					h->probe_last = Pp->loc_3;
					h->probe_tm = Pp->tm; h->probe_dH = Pp->dH; h->probe_dS = Pp->dS;
					h->probe_mm = (int8_t)Pp->num_mm; h->probe_gap = (int8_t)Pp->num_gap;
					h->probe_strand = (Pp->mask & M_PLUS) ? 0 : 1;
					set_str(h->probe_align, sizeof(h->probe_align), Pp->alignment);
				}
				if (!has_probe) break;
			}
		}
	}
	free(ml.v);
}

/* hybrid() (probe_search.cpp:67-230) */
static void search_probe(search_ctx_t *cx, const assay_t *as, const ref_options *o)
{
	const float ct = o->probe_strand/as->pdeg; /* strand(c): Ct = c (nuc_cruc.h:879-886) */
	bind_limits_t bl = {o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
		o->probe_clamp_5, o->probe_clamp_3, o->max_gap, o->max_mismatch, o->max_poly_degen};
	for (int pass = 0; pass < 2; ++pass) {
		const int plus = pass;
		if (!(o->target_strand & (plus ? 1 : 2))) continue;
		olist_t b = {0};
		bind_hashed(cx, &b, as->P, plus, ct, &bl, 0);
		for (long k = 0; k < b.n; ++k) {
			const oinfo_t *e = &b.v[k];
			ref_hit *h = new_hit();
			h->probe_tm = e->tm; h->probe_dH = e->dH; h->probe_dS = e->dS;
			h->probe_mm = (int8_t)e->num_mm; h->probe_gap = (int8_t)e->num_gap;
			h->probe_first = e->loc_5;
			h->probe_last = e->loc_3;
			h->probe_strand = plus ? 0 : 1;
			set_str(h->probe_align, sizeof(h->probe_align), e->alignment);
			if (plus) amplicon_plus(cx, h, e->loc_5, e->loc_3, 0);
			else amplicon_minus(cx, h, e->loc_5, e->loc_3, 0);
		}
		free(b.v);
	}
}

/* padlock() (padlock_search.cpp:62-361) */
static void search_padlock(search_ctx_t *cx, const assay_t *as, const ref_options *o, int max_len)
{
	const float f_ct = o->forward_primer_strand/as->fdeg - 0.5f*0.0f;
	const float r_ct = o->reverse_primer_strand/as->rdeg - 0.5f*0.0f;
	bind_limits_t up = {o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
		o->probe_clamp_5, 0, o->max_gap, o->max_mismatch, o->max_poly_degen};
	bind_limits_t dn = {o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
		0, o->probe_clamp_3, o->max_gap, o->max_mismatch, o->max_poly_degen};
	for (int pass = 0; pass < 2; ++pass) {
		const int plus = pass;
		olist_t U = {0}, D = {0};
		if (o->target_strand & (plus ? 1 : 2)) {
			bind_hashed(cx, &U, as->R, plus, r_ct, &up, 1);
			bind_hashed(cx, &D, as->F, plus, f_ct, &dn, 1);
		}
		for (long u = 0; u < U.n; ++u)
			for (long d = 0; d < D.n; ++d) {
				const oinfo_t *ue = &U.v[u], *de = &D.v[d];
				const int gap = plus ? de->loc_5 - ue->loc_3 - 1 : ue->loc_5 - de->loc_3 - 1;
				if (gap < 0 || gap > max_len) continue;
				const int start = plus ? ue->loc_5 : de->loc_5;
				const int stop = plus ? de->loc_3 : ue->loc_3;
				ref_hit *h = new_hit();
				set_str(h->forward_oligo, sizeof(h->forward_oligo), as->F);
				set_str(h->reverse_oligo, sizeof(h->reverse_oligo), as->R);
				h->primer_strand = plus ? 0 : 1;
				h->amp_first = start;
				h->amp_last = stop;
				h->forward_tm = de->tm; h->forward_dH = de->dH; h->forward_dS = de->dS;
				h->reverse_tm = ue->tm; h->reverse_dH = ue->dH; h->reverse_dS = ue->dS;
				h->forward_mm = (int8_t)de->num_mm; h->reverse_mm = (int8_t)ue->num_mm;
				h->forward_gap = (int8_t)de->num_gap; h->reverse_gap = (int8_t)ue->num_gap;
				h->forward_clamp = (int8_t)de->anchor_3;
				h->reverse_clamp = (int8_t)ue->anchor_5;
				set_str(h->forward_align, sizeof(h->forward_align), de->alignment);
				set_str(h->reverse_align, sizeof(h->reverse_align), ue->alignment);
				if (!plus) {
					/* minus-strand ligation site: text copied forward from max(0,start),
					 * first index max(0, 1 - start) (padlock_search.cpp:206-218) */
					const int fi = 1 - start > 0 ? 1 - start : 0;
					amplicon_plus(cx, h, start, stop, fi);
				}
				else {
					/* plus-strand site: complemented, first index max(0, stop - len - 1)
					 * (padlock_search.cpp:341-352) */
					const int fi = stop - (int)cx->len - 1 > 0 ? stop - (int)cx->len - 1 : 0;
					amplicon_minus(cx, h, start, stop, fi);
				}
			}
		free(U.v);
		free(D.v);
	}
}

long orc_search(const uint8_t *codes, uint32_t len, const char *forward, const char *reverse,
	const char *probe, int forward_degen, int reverse_degen, int probe_degen, const ref_options *o)
{
	g_nhits = 0;
	g_align_count = 0;
	pool_free();
	if (o->word_size < 2 || o->word_size > 8) return fail("DNAHash: Unsupported word length");
	if (len < (uint32_t)o->word_size) return 0; /* tntblast_local.cpp:513-529 */

	search_ctx_t cx;
	memset(&cx, 0, sizeof(cx));
	cx.codes = codes;
	cx.len = len;
	khash_build(&cx.hash, codes, len, o->word_size);
	cx.nc = get_nc(o->target_T, o->salt);
	cx.nc->dangle5 = o->dangle5;
	cx.nc->dangle3 = o->dangle3;

	assay_t as = {forward, reverse, probe, forward_degen, reverse_degen, probe_degen};
	const int has_primers = forward && reverse && forward[0] && reverse[0];
	const int has_probe = probe && probe[0];

	if (has_primers) {
		switch (o->assay_format) {
		case 0: search_pcr(&cx, &as, o); break;
		case 2: search_padlock(&cx, &as, o, 0); break;
		case 3: search_padlock(&cx, &as, o, (int)o->max_len); break;
		default: khash_free(&cx.hash); return fail("orc_search: unsupported assay format");
		}
	}
	else if (has_probe) search_probe(&cx, &as, o);

	khash_free(&cx.hash);
	cache_free(&cx.plus_cache);
	cache_free(&cx.minus_cache);
	return g_nhits;
}

int orc_get_hits(ref_hit *out, long cap)
{
	const long n = cap < g_nhits ? cap : g_nhits;
	if (n > 0) memcpy(out, g_hits, sizeof(ref_hit)*(size_t)n);
	return (int)n;
}
