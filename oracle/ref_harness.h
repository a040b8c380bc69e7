/* TEST INFRASTRUCTURE ONLY.
 * Plain-C records shared by the reference harness (oracle/_ref/libtntref.so, built from
 * /root/reference by oracle/Makefile) and the C restatement (oracle/libtntoracle.so).
 * Both libraries fill the same structs so tests can memcmp/compare them field by field. */
#ifndef TNT_REF_HARNESS_H
#define TNT_REF_HARNESS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REF_NUM_BASE_PAIR 49
#define REF_TABLE_SIZE (REF_NUM_BASE_PAIR*REF_NUM_BASE_PAIR)
#define REF_MAX_LOOP 512

typedef struct {
	int32_t delta_g[REF_TABLE_SIZE];          /* nuc_cruc.h:606, filled by update_dp_param */
	float param_H[REF_TABLE_SIZE];            /* nuc_cruc.h:608 */
	float param_S[REF_TABLE_SIZE];
	float loop_terminal_H[REF_TABLE_SIZE];    /* nuc_cruc.h:617 */
	float loop_terminal_S[REF_TABLE_SIZE];
	float loop_S[REF_MAX_LOOP + 1];           /* nuc_cruc.h:614 */
	float bulge_S[REF_MAX_LOOP + 1];          /* nuc_cruc.h:623 */
	float supp[12];                           /* nuc_cruc.h:640 (NUM_SUPP_PARAM) */
	float supp_salt[4];                       /* nuc_cruc.h:643 */
	float init_H, init_S;
	float AT_closing_H, AT_closing_S;
	float symmetry_S;
	float SALT;
	float asymmetric_loop_dS;
	float bulge_AT_closing_S;
	uint8_t watson_and_crick[REF_NUM_BASE_PAIR];
} ref_tables;

typedef struct {
	float tm, dH, dS, dG, dp_dg;
	int32_t valid;
	int32_t anchor5, anchor3;
	int32_t num_mismatch, num_gap, max_poly_degen;
	int32_t q_first, q_last, t_first, t_last;   /* alignment_range() */
	int32_t target_start, target_stop;          /* window in fragment coordinates (bind only) */
	int32_t loc_5, loc_3;                       /* mapped target coordinates (bind only) */
	char alignment[512];                        /* operator<< text */
} ref_align_out;

typedef struct {
	int32_t assay_format;                        /* hybrid_sig.h:19 enum: 0 PCR, 1 PROBE, 2 PADLOCK, 3 MIPS */
	int32_t word_size;
	float target_T, salt;
	int32_t dangle5, dangle3;
	float forward_primer_strand, reverse_primer_strand, probe_strand;
	float min_primer_tm, max_primer_tm, min_primer_dg, max_primer_dg;
	float min_probe_tm, max_probe_tm, min_probe_dg, max_probe_dg;
	uint32_t primer_clamp;
	int32_t min_max_primer_clamp;
	uint32_t probe_clamp_5, probe_clamp_3;
	uint32_t max_gap, max_mismatch, max_poly_degen;
	uint32_t max_len;
	int32_t single_primer_pcr;
	int32_t target_strand;                       /* seq.h:36-40 bit mask: 1 plus, 2 minus */
} ref_options;

typedef struct {
	int32_t primer_strand, probe_strand;         /* hybrid_sig::PLUS=0 / MINUS=1 */
	int32_t amp_first, amp_last, probe_first, probe_last;
	float forward_tm, forward_dH, forward_dS;
	float reverse_tm, reverse_dH, reverse_dS;
	float probe_tm, probe_dH, probe_dS;
	int32_t forward_mm, forward_gap, reverse_mm, reverse_gap, probe_mm, probe_gap;
	int32_t forward_clamp, reverse_clamp;
	int32_t amplicon_len;
	uint64_t amplicon_fnv;                       /* FNV-1a of the amplicon text */
	char forward_oligo[128], reverse_oligo[128];
	char forward_align[512], reverse_align[512], probe_align[512];
	char amplicon_head[256];
} ref_hit;

const char *ref_last_error(void);
/* melt.dinkelbach(on) (nuc_cruc.h:763-766, tntblast_local.cpp:367) for every later call of this thread (default off) */
void ref_set_dinkelbach(int on);
int ref_dump_tables(float T, float na, ref_tables *out);
long ref_seeds_raw(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int complement, uint32_t *q_out, uint32_t *t_out, long cap);
long ref_seeds_unique(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int plus_strand, uint32_t *q_out, uint32_t *t_out, long cap);
int ref_align(const char *query, const uint8_t *target, int target_len, float T, float na,
	float strand_conc, int dangle5, int dangle3, ref_align_out *out);
int ref_bind_window(const uint8_t *codes, uint32_t len, const char *oligo, int plus_strand,
	uint32_t query_loc, uint32_t target_loc, float T, float na, float strand_conc,
	int dangle5, int dangle3, ref_align_out *out);
long ref_search(const uint8_t *codes, uint32_t len, const char *forward, const char *reverse,
	const char *probe, int forward_degen, int reverse_degen, int probe_degen,
	const ref_options *o);
int ref_get_hits(ref_hit *out, long cap);
/* homodimer (target == NULL) / heterodimer of two oligos, as tntblast_local.cpp:657-686 computes them */
int ref_dimer(const char *query, const char *target, float T, float na, float conc_a, float conc_b, ref_align_out *out);

/* FASTA reader of the reference (sequence_data, sequence_data_fastx.cpp) on a file */
long ref_fasta_open(const char *path);                      /* number of records, -1 on error */
long ref_fasta_approx_len(long index);                      /* sequence_data::approx_seq_len */
long ref_fasta_read(long index, uint32_t start, uint32_t stop, int whole, uint8_t *out, long cap,
	char *defline, long defline_cap);                       /* read_bio_seq; bases, -1 on throw */
void ref_fasta_close(void);
void ref_seq_len_increment(uint32_t len, uint32_t max_len, uint32_t *delta, uint32_t *pieces);

/* approximate_tm_hairpin of one oligo (nuc_cruc.cpp:2542-2618), as tntblast_local.cpp:657-686 calls it */
int ref_hairpin(const char *query, float T, float na, ref_align_out *out);
/* hairpin parameter tables of the reference (data): loop entropies by loop length, and the special
 * tri- / tetra-loops -- 5 or 6 letters each (found by probing find_loop_index with every 5- and
 * 6-mer), with their dH / dS bonuses.  Returns the number of special loops (<= cap). */
#define REF_MAX_HAIRPIN 512
int ref_hairpin_tables(float *hairpin_S /* [REF_MAX_HAIRPIN + 1] */, char (*loops)[8], float *special_H, float *special_S, int cap);

/* Post-processing of a result list by the reference's own select_best_match / uniquify_results
 * (tntblast_util.cpp:1482-1755) and hybrid_sig::operator< sort (tntblast_local.cpp:918-930).  `hits`
 * is ONE list in the order the driver would hold it (all records of one assay id); the indices of
 * the surviving records come back in output order.  Returns their number, -1 on a throw. */
typedef struct {
	int32_t id, degen_id, seq_id;
	int32_t has_primers, has_probe;
	int32_t amp_first, amp_last, probe_first, probe_last;
	float forward_tm, reverse_tm, probe_tm;
	int32_t forward_len, reverse_len;            /* lengths of the oligos in the forward / reverse slot */
	const char *forward_align, *reverse_align, *probe_align;
} ref_post_hit;
long ref_finalize(const ref_post_hit *hits, long n, int best_match, int uniquify, int32_t *order_out);

#ifdef __cplusplus
}
#endif
#endif
