/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the tntblast v2.77 search hot path.
 *
 * Exports the same entry points as oracle/_ref/libtntref.so (see ref_harness.h) with the
 * prefix orc_ instead of ref_, filling the same records, so tests can run the reference, the
 * restatement and the CUDA engine on identical inputs.  Never imported by the product. */
#ifndef TNT_ORACLE_H
#define TNT_ORACLE_H

#include "ref_harness.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *orc_last_error(void);
/* NucCruc::dinkelbach(bool) for every later call of this thread (default off) */
void orc_set_dinkelbach(int on);
int orc_dump_tables(float T, float na, ref_tables *out);
long orc_seeds_raw(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int complement, uint32_t *q_out, uint32_t *t_out, long cap);
long orc_seeds_unique(const uint8_t *codes, uint32_t len, int word_size, const char *oligo,
	int plus_strand, uint32_t *q_out, uint32_t *t_out, long cap);
int orc_align(const char *query, const uint8_t *target, int target_len, float T, float na,
	float strand_conc, int dangle5, int dangle3, ref_align_out *out);
int orc_bind_window(const uint8_t *codes, uint32_t len, const char *oligo, int plus_strand,
	uint32_t query_loc, uint32_t target_loc, float T, float na, float strand_conc,
	int dangle5, int dangle3, ref_align_out *out);
long orc_search(const uint8_t *codes, uint32_t len, const char *forward, const char *reverse,
	const char *probe, int forward_degen, int reverse_degen, int probe_degen,
	const ref_options *o);
int orc_get_hits(ref_hit *out, long cap);

/* number of NucCruc alignments (approximate_tm_heterodimer equivalents, cache misses only)
 * executed by the last orc_search call -- the unit of the "alignments/s" metric */
long orc_last_alignment_count(void);
int orc_dimer(const char *query, const char *target, float T, float na, float conc_a, float conc_b, ref_align_out *out);
int orc_hairpin(const char *query, float T, float na, ref_align_out *out);

/* FASTA reader + fragment rule (tnt_oracle_fasta.c) */
long orc_fasta_index(const char *text, uint64_t n, uint64_t *pos, long cap);
long orc_fasta_read(const char *text, uint64_t rec_begin, uint64_t rec_end, uint32_t start, uint32_t stop,
	uint8_t *out, long cap, uint64_t *def_begin, uint32_t *def_len);
void orc_seq_len_increment(uint32_t len, uint32_t max_len, uint32_t *delta, uint32_t *pieces);
long orc_fragments(uint32_t len, uint32_t max_len, uint32_t *start_out, uint32_t *stop_out, long cap);

#ifdef __cplusplus
}
#endif
#endif
