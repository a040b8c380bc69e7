/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's FASTA reader and of the way
 * its driver cuts records into fragments.  Never imported, linked or executed by the product.
 *
 * Follows (tntblast v2.77):
 *   sequence_data::load_fasta               sequence_data_fastx.cpp:13-79
 *   sequence_data::read_bio_seq_fasta_slow  sequence_data_fastx.cpp:190-382
 *   ascii_to_hash_base                      seq.h:148-189
 *   seq_len_increment                       sequence_data.cpp:739-754
 *   work queue of the search driver         tntblast_local.cpp:282-289,448-468,510-511
 *
 * Parity: pinned against oracle/_ref (the compiled reference's own sequence_data class, called
 * through ref_fasta_* in ref_harness.cpp) by tests/test_oracle_golden.py and by the committed
 * vectors tests/golden/fasta.json (generated from _ref by tests/golden/make_golden.py). */
#include <ctype.h>
#include <stdint.h>
#include <string.h>

#include "tnt_oracle.h"

/* load_fasta :33-58: file offsets of the deflines; a '>' opens a record unless an earlier '>'
 * of the same line already did; only '\n' ends the line.  Returns the number of records;
 * pos[] receives at most cap offsets. */
long orc_fasta_index(const char *text, uint64_t n, uint64_t *pos, long cap)
{
	long count = 0;
	int read_fasta = 0;
	for (uint64_t i = 0; i < n; ++i) {
		if (!read_fasta && text[i] == '>') {
			read_fasta = 1;
			if (count < cap) pos[count] = i;
			++count;
		}
		else if (text[i] == '\n') read_fasta = 0;
	}
	return count;
}

/* seq.h:148-189 */
static uint8_t orc_ascii_to_hash_base(char c)
{
	switch (toupper((unsigned char)c)) {
		case 'A': return 0;
		case 'C': return 1;
		case 'G': return 2;
		case 'T': case 'U': return 3;
		case 'I': return 4;
		case 'M': return 5;
		case 'R': return 6;
		case 'S': return 7;
		case 'V': return 8;
		case 'W': return 9;
		case 'Y': return 10;
		case 'H': return 11;
		case 'K': return 12;
		case 'D': return 13;
		case 'B': return 14;
		case 'N': return 15;
		case '-': return 16;
	}
	return 17;
}

/* read_bio_seq_fasta_slow(m_seq, index, start, stop) on the record text[rec_begin, rec_end):
 * bases start..stop (inclusive, counted over the sequence characters) go to out[] (at most
 * cap), the defline range to def_begin / def_len.  Returns the number of bases, or -1 where
 * the reference throws "Truncated fasta file detected!" (:259-275). */
long orc_fasta_read(const char *text, uint64_t rec_begin, uint64_t rec_end, uint32_t start, uint32_t stop,
	uint8_t *out, long cap, uint64_t *def_begin, uint32_t *def_len)
{
	uint64_t p = rec_begin + 1;                                /* :252 skip the '>' */
	while (p < rec_end && isspace((unsigned char)text[p])) ++p; /* :254-257 */
	if (p == rec_end) return -1;
	const uint64_t d0 = p;
	while (p < rec_end && text[p] != '\n' && text[p] != '\r') ++p; /* :267-269 */
	if (p == rec_end) return -1;
	if (def_begin) *def_begin = d0;
	if (def_len) *def_len = (uint32_t)(p - d0);
	/* :285-301: the size estimate bounds stop */
	const uint64_t seq_size = 4 + (rec_end - rec_begin) - (p - rec_begin);
	uint32_t last = stop;
	if ((int32_t)stop < 0 || (uint64_t)stop >= seq_size) last = (uint32_t)(seq_size - 1);
	long nb = 0;
	uint32_t index = 0;
	for (; p < rec_end; ++p) {                                 /* :327-376 */
		if (index > last) break;
		const char c = text[p];
		if (!isspace((unsigned char)c) && c != '*' && c != '-' && c != '\r' && (index++ >= start)) {
			if (nb < cap) out[nb] = orc_ascii_to_hash_base(c);
			++nb;
		}
	}
	return nb;
}

/* seq_len_increment (sequence_data.cpp:739-754): (piece length, number of pieces) */
void orc_seq_len_increment(uint32_t len, uint32_t max_len, uint32_t *delta, uint32_t *pieces)
{
	if (len <= max_len) { *delta = len - 1; *pieces = 1; return; }
	uint32_t n = 2;
	while ((uint64_t)len > (uint64_t)n*max_len) ++n;
	*delta = len/n + ((len % n) ? 1u : 0u);
	*pieces = n;
}

/* The (start, stop) pairs the driver hands out for one record of approximate length `len`
 * (tntblast_local.cpp:282-289 first piece, :448-468 the following ones).  Returns the count. */
long orc_fragments(uint32_t len, uint32_t max_len, uint32_t *start_out, uint32_t *stop_out, long cap)
{
	uint32_t delta, pieces;
	orc_seq_len_increment(len, max_len, &delta, &pieces);
	const uint32_t max_stop = len - 1;
	uint32_t start = 0, stop = delta;
	long n = 0;
	while (1) {
		if (n < cap) { start_out[n] = start; stop_out[n] = stop; }
		++n;
		if (stop == max_stop) break;
		start = stop + 1;
		stop = (stop + delta < max_stop) ? stop + delta : max_stop;
	}
	return n;
}
