// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// C-ABI harness around the *unmodified* reference objects (tntblast v2.77) that
// oracle/Makefile compiles from the sources where they lie under /root/reference.
// The result (oracle/_ref/libtntref.so) is the ground truth that
//   * pins the plain-C restatement in oracle/tnt_oracle.c, and
//   * is compared with the CUDA engine in tests/ (-m gpu) on identical inputs.
//
// Nothing in here re-implements reference logic except the window loader in
// ref_bind_window(), which replays bind_oligo.cpp:502-592 / :1205-1295 through
// NucCruc's own push_front_target/push_back_target so that a single
// (oligo, seed) candidate can be examined in isolation.  The end-to-end entry
// point ref_search() calls the reference's own amplicon()/hybrid()/padlock().
//
// `#define private public` is used solely to read NucCruc's parameter tables
// (delta_g, param_H, ...) for the table-export test; no private method is called.

#include <cstdint>
#include <cstring>
#include <cmath>
#include <sstream>
#include <string>
#include <vector>
#include <list>
#include <unordered_map>

#define private public
#include "nuc_cruc.h"
#undef private

#include "tntblast.h"
#include "seq_hash.h"
#include "hybrid_sig.h"
#include "compress.h"
#include "sequence_data.h"

#include "ref_harness.h"

static thread_local std::string g_err;

extern "C" const char *ref_last_error() { return g_err.c_str(); }

#define GUARD_BEGIN try {
#define GUARD_END                                   \
	}                                               \
	catch (const char *e) { g_err = e; return -1; } \
	catch (const std::string &e) { g_err = e; return -1; } \
	catch (std::exception &e) { g_err = e.what(); return -1; } \
	catch (...) { g_err = "unknown exception"; return -1; }

// ---------------------------------------------------------------------------
// 1. Parameter tables for a given (T, [Na+])
// ---------------------------------------------------------------------------
extern "C" int ref_dump_tables(float T, float na, ref_tables *out)
{
	GUARD_BEGIN
	NucCruc melt(NucCruc::SANTA_LUCIA, T);
	melt.Salt(na);

	memcpy(out->delta_g, melt.delta_g, sizeof(out->delta_g));
	memcpy(out->param_H, melt.param_H, sizeof(out->param_H));
	memcpy(out->param_S, melt.param_S, sizeof(out->param_S));
	memcpy(out->loop_terminal_H, melt.param_loop_terminal_H, sizeof(out->loop_terminal_H));
	memcpy(out->loop_terminal_S, melt.param_loop_terminal_S, sizeof(out->loop_terminal_S));
	memcpy(out->loop_S, melt.param_loop_S, sizeof(out->loop_S));
	memcpy(out->bulge_S, melt.param_bulge_S, sizeof(out->bulge_S));
	memcpy(out->supp, melt.param_supp, sizeof(out->supp));
	memcpy(out->supp_salt, melt.param_supp_salt, sizeof(out->supp_salt));
	out->init_H = melt.param_init_H;
	out->init_S = melt.param_init_S;
	out->AT_closing_H = melt.param_AT_closing_H;
	out->AT_closing_S = melt.param_AT_closing_S;
	out->symmetry_S = melt.param_symmetry_S;
	out->SALT = melt.param_SALT;
	out->asymmetric_loop_dS = melt.param_asymmetric_loop_dS;
	out->bulge_AT_closing_S = melt.param_bulge_AT_closing_S;
	for (int i = 0; i < 49; ++i) out->watson_and_crick[i] = melt.watson_and_crick[i] ? 1 : 0;
	return 0;
	GUARD_END
}

// ---------------------------------------------------------------------------
// 2. Seeds: DNAHash over a SEQPTR fragment, queried with an ASCII oligo
// ---------------------------------------------------------------------------
static std::vector<unsigned char> make_seqptr(const uint8_t *codes, uint32_t len)
{
	std::vector<unsigned char> buf(sizeof(unsigned int) + len);
	unsigned int n = len;
	memcpy(buf.data(), &n, sizeof(n));
	if (len) memcpy(buf.data() + sizeof(unsigned int), codes, len);
	return buf;
}

// Raw iteration order of DNAHash::find / find_complement (seq_hash.h:749-779, :244-274):
// pairs (offset(), *iter).  Returns the number of seeds (may exceed cap; only cap are stored).
extern "C" long ref_seeds_raw(const uint8_t *codes, uint32_t len, int word_size,
	const char *oligo, int complement, uint32_t *q_out, uint32_t *t_out, long cap)
{
	GUARD_BEGIN
	std::vector<unsigned char> buf = make_seqptr(codes, len);
	SEQPTR seq = buf.data();
	DNAHash dbase((unsigned char)word_size);
	dbase.hash(seq, SEQ_SIZE(seq), 0, SEQ_SIZE(seq));

	const std::string o(oligo);
	long n = 0;
	DNAHash::iterator it = complement ? dbase.find_complement(o) : dbase.find(o);
	for (; it != dbase.end(); ++it) {
		if (n < cap) {
			q_out[n] = (uint32_t)it.offset();
			t_out[n] = (uint32_t)(*it);
		}
		++n;
	}
	return n;
	GUARD_END
}

// match_oligo_to_{minus,plus}_strand (bind_oligo.cpp:84-122): one seed per diagonal.
extern "C" long ref_seeds_unique(const uint8_t *codes, uint32_t len, int word_size,
	const char *oligo, int plus_strand, uint32_t *q_out, uint32_t *t_out, long cap)
{
	GUARD_BEGIN
	std::vector<unsigned char> buf = make_seqptr(codes, len);
	SEQPTR seq = buf.data();
	DNAHash dbase((unsigned char)word_size);
	dbase.hash(seq, SEQ_SIZE(seq), 0, SEQ_SIZE(seq));

	std::list<oligo_info> info;
	if (plus_strand) match_oligo_to_plus_strand(info, dbase, std::string(oligo), oligo_info::F);
	else match_oligo_to_minus_strand(info, dbase, std::string(oligo), oligo_info::F);

	long n = 0;
	for (std::list<oligo_info>::const_iterator i = info.begin(); i != info.end(); ++i, ++n) {
		if (n < cap) {
			q_out[n] = i->query_loc;
			t_out[n] = i->target_loc;
		}
	}
	return n;
	GUARD_END
}

// ---------------------------------------------------------------------------
// 3. One (oligo, window) evaluation == one "alignment" of the headline metric
// ---------------------------------------------------------------------------
// One NucCruc per thread, rebuilt when (T, [Na+]) changes (the 15 MB dp_matrix makes
// per-call construction far too slow for million-window differential runs).
static thread_local bool g_dinkelbach = false;
extern "C" void ref_set_dinkelbach(int on) { g_dinkelbach = on != 0; }

static NucCruc *get_melt(float T, float na)
{
	static thread_local NucCruc *melt = NULL;
	static thread_local float cur_T = -1.0f, cur_na = -1.0f;
	if (melt && (cur_T != T || cur_na != na)) { delete melt; melt = NULL; }
	if (!melt) {
		melt = new NucCruc(NucCruc::SANTA_LUCIA, T);
		melt->Salt(na);
		cur_T = T; cur_na = na;
	}
	melt->dinkelbach(g_dinkelbach);
	return melt;
}

static void fill_align_out(NucCruc &melt, float tm, ref_align_out *out)
{
	memset(out, 0, sizeof(*out));
	out->tm = tm;
	out->dH = melt.delta_H();
	out->dS = melt.delta_S();
	out->dG = melt.delta_G();
	out->valid = melt.curr_align.valid ? 1 : 0;
	out->dp_dg = melt.curr_align.dp_dg;

	if (!melt.curr_align.valid) {
		// Everything else is undefined in the reference when no alignment exists.
		return;
	}

	out->anchor5 = melt.anchor5_query();
	out->anchor3 = melt.anchor3_query();
	out->num_mismatch = melt.num_mismatch();
	out->num_gap = melt.num_gap();
	out->max_poly_degen = melt.max_contiguous_target_degen();

	std::pair<unsigned int, unsigned int> qr, tr;
	melt.alignment_range(qr, tr);
	out->q_first = (int)qr.first;
	out->q_last = (int)qr.second;
	out->t_first = (int)tr.first;
	out->t_last = (int)tr.second;

	std::stringstream ss;
	ss << melt;
	const std::string s = ss.str();
	strncpy(out->alignment, s.c_str(), sizeof(out->alignment) - 1);
}

// Query = ASCII oligo; target = NucCruc base codes (BASE::nucleic_acid values) 5'->3'.
extern "C" int ref_align(const char *query, const uint8_t *target, int target_len,
	float T, float na, float strand_conc, int dangle5, int dangle3, ref_align_out *out)
{
	GUARD_BEGIN
	NucCruc *melt = get_melt(T, na);
	melt->dangle(dangle5 != 0, dangle3 != 0);
	melt->strand(strand_conc);
	melt->set_query(std::string(query));
	melt->clear_target();
	for (int i = 0; i < target_len; ++i) melt->push_back_target((BASE::nucleic_acid)target[i]);

	const float tm = melt->approximate_tm_heterodimer();
	fill_align_out(*melt, tm, out);
	return 0;
	GUARD_END
}

// The oligo-only duplex temperatures of tntblast_local.cpp:657-686.
extern "C" int ref_dimer(const char *query, const char *target, float T, float na, float conc_a, float conc_b, ref_align_out *out)
{
	GUARD_BEGIN
	NucCruc *melt = get_melt(T, na);
	melt->dangle(false, false);
	float tm;
	if (target) {
		melt->set_query(std::string(query));
		melt->set_target(std::string(target));
		melt->strand(conc_a, conc_b);
		tm = melt->approximate_tm_heterodimer();
	}
	else {
		melt->set_duplex(std::string(query));
		melt->strand(conc_a, conc_b);
		tm = melt->approximate_tm_homodimer();
	}
	fill_align_out(*melt, tm, out);
	return 0;
	GUARD_END
}

// Replay of one candidate exactly as bind_oligo_to_{minus,plus}_strand would see it:
// fragment codes (seq.h DB_* values), seed (query_loc, target_loc), strand.
// Fills the window bounds and the mapped target coordinates as well.
extern "C" int ref_bind_window(const uint8_t *codes, uint32_t len, const char *oligo,
	int plus_strand, uint32_t query_loc, uint32_t target_loc,
	float T, float na, float strand_conc, int dangle5, int dangle3, ref_align_out *out)
{
	GUARD_BEGIN
	NucCruc *melt = get_melt(T, na);
	melt->dangle(dangle5 != 0, dangle3 != 0);
	melt->strand(strand_conc);
	const std::string q(oligo);
	melt->set_query(q);

	const unsigned int window = melt->size_query();
	const unsigned int target_length = window + 2*NUM_FLANK_BASE;
	unsigned int target_start = std::max(int(target_loc) - int(query_loc + NUM_FLANK_BASE), 0);
	unsigned int target_stop = std::min(target_start + target_length, (unsigned int)len);

	melt->clear_target();

	// Complement table == the switch at bind_oligo.cpp:524-591 (exercised through the
	// reference's own char_to_complement_nucleic_acid for the letter of each code).
	for (unsigned int i = target_start; i < target_stop; ++i) {
		const unsigned char c = codes[i];
		if (c > DB_N) continue; // GAP / UNKNOWN are silently dropped (bind_oligo.cpp:574-591)
		const char letter = hash_base_to_ascii(c);
		if (plus_strand) melt->push_back_target(BASE::char_to_nucleic_acid(letter));
		else melt->push_front_target(BASE::char_to_complement_nucleic_acid(letter));
	}

	const float tm = melt->approximate_tm_heterodimer();
	fill_align_out(*melt, tm, out);
	out->target_start = (int)target_start;
	out->target_stop = (int)target_stop;

	if (melt->curr_align.valid) {
		int t5 = target_start, t3 = target_start;
		if (plus_strand) { // bind_oligo.cpp:1424-1434
			t5 += out->t_first;
			t3 += out->t_last;
			t3 += out->q_first;
			t5 -= (int)(window - 1) - out->q_last;
		}
		else { // bind_oligo.cpp:721-731
			t5 += target_stop - target_start - 1 - out->t_last;
			t3 += target_stop - target_start - 1 - out->t_first;
			t5 -= out->q_first;
			t3 += (int)(window - 1) - out->q_last;
		}
		out->loc_5 = t5;
		out->loc_3 = t3;
	}
	return 0;
	GUARD_END
}

// ---------------------------------------------------------------------------
// 4. End to end: the reference's amplicon()/hybrid()/padlock() on one fragment
// ---------------------------------------------------------------------------
static thread_local std::vector<ref_hit> g_hits;

static void copy_str(char *dst, size_t cap, const std::string &s)
{
	strncpy(dst, s.c_str(), cap - 1);
	dst[cap - 1] = '\0';
}

extern "C" long ref_search(const uint8_t *codes, uint32_t len, const char *forward,
	const char *reverse, const char *probe, int forward_degen, int reverse_degen,
	int probe_degen, const ref_options *o)
{
	GUARD_BEGIN
	g_hits.clear();

	std::vector<unsigned char> buf = make_seqptr(codes, len);
	std::pair<std::string, SEQPTR> bio_seq("target", buf.data());

	DNAHash dbase((unsigned char)o->word_size);
	if (len < dbase.min_sequence_size()) return 0; // tntblast_local.cpp:513-529
	dbase.hash(bio_seq.second, SEQ_SIZE(bio_seq.second), 0, SEQ_SIZE(bio_seq.second));

	NucCruc melt(NucCruc::SANTA_LUCIA, o->target_T);
	melt.Salt(o->salt);
	melt.dangle(o->dangle5 != 0, o->dangle3 != 0);
	melt.dinkelbach(g_dinkelbach);

	std::unordered_map<BindCacheKey, BindCacheValue> plus_cache, minus_cache;
	std::unordered_map<std::string, size_t> str_table;
	std::vector<std::string> oligo_table;

	const bool has_primers = forward && reverse && forward[0] && reverse[0];
	const bool has_probe = probe && probe[0];

	size_t fi = INVALID_INDEX, ri = INVALID_INDEX, pi = INVALID_INDEX;
	oligo_table.push_back("assay");
	str_table["assay"] = 0;
	if (has_primers) {
		fi = str_to_index(std::string(forward), str_table);
		if (fi == oligo_table.size()) oligo_table.push_back(forward);
		ri = str_to_index(std::string(reverse), str_table);
		if (ri == oligo_table.size()) oligo_table.push_back(reverse);
	}
	if (has_probe) {
		pi = str_to_index(std::string(probe), str_table);
		if (pi == oligo_table.size()) oligo_table.push_back(probe);
	}

	hybrid_sig sig;
	if (has_primers && has_probe) sig = hybrid_sig(0, fi, ri, pi, 0);
	else if (has_primers) sig = hybrid_sig(0, fi, ri, 0);
	else sig = hybrid_sig(0, pi, 0);
	sig.forward_degen = forward_degen;
	sig.reverse_degen = reverse_degen;
	sig.probe_degen = probe_degen;

	std::list<hybrid_sig> res;

	if (sig.has_primers()) {
		switch (o->assay_format) {
		case ASSAY_PCR:
			res = amplicon(dbase, bio_seq, sig, melt, plus_cache, minus_cache,
				o->forward_primer_strand, o->reverse_primer_strand, o->probe_strand,
				o->min_primer_tm, o->max_primer_tm, o->min_primer_dg, o->max_primer_dg,
				o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
				o->primer_clamp, o->min_max_primer_clamp, o->probe_clamp_5, o->probe_clamp_3,
				o->max_gap, o->max_mismatch, o->max_poly_degen, o->max_len,
				o->single_primer_pcr != 0, 0 /*NO_MASK*/, oligo_table, str_table);
			break;
		case ASSAY_PADLOCK:
			res = padlock(dbase, bio_seq, sig, melt, plus_cache, minus_cache,
				o->forward_primer_strand, o->reverse_primer_strand,
				o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
				o->probe_clamp_5, o->probe_clamp_3, o->max_gap, o->max_mismatch,
				o->max_poly_degen, o->target_strand, 0, oligo_table, str_table);
			break;
		case ASSAY_MIPS:
			res = padlock(dbase, bio_seq, sig, melt, plus_cache, minus_cache,
				o->forward_primer_strand, o->reverse_primer_strand,
				o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
				o->probe_clamp_5, o->probe_clamp_3, o->max_gap, o->max_mismatch,
				o->max_poly_degen, o->target_strand, (int)o->max_len, oligo_table, str_table);
			break;
		default:
			THROW("ref_search: unsupported assay format");
		}
	}
	else if (sig.has_probe()) {
		res = hybrid(dbase, bio_seq, sig, melt, o->probe_strand,
			o->min_probe_tm, o->max_probe_tm, o->min_probe_dg, o->max_probe_dg,
			o->probe_clamp_5, o->probe_clamp_3, o->max_gap, o->max_mismatch,
			o->max_poly_degen, o->target_strand, oligo_table, str_table);
	}

	const std::vector<std::string> keys = ordered_keys(str_table);

	for (std::list<hybrid_sig>::const_iterator i = res.begin(); i != res.end(); ++i) {
		ref_hit h;
		memset(&h, 0, sizeof(h));
		h.primer_strand = i->primer_strand;
		h.probe_strand = i->probe_strand;
		h.amp_first = i->amplicon_range.first;
		h.amp_last = i->amplicon_range.second;
		h.probe_first = i->probe_range.first;
		h.probe_last = i->probe_range.second;
		h.forward_tm = i->forward_tm; h.forward_dH = i->forward_dH; h.forward_dS = i->forward_dS;
		h.reverse_tm = i->reverse_tm; h.reverse_dH = i->reverse_dH; h.reverse_dS = i->reverse_dS;
		h.probe_tm = i->probe_tm; h.probe_dH = i->probe_dH; h.probe_dS = i->probe_dS;
		h.forward_mm = i->forward_mm; h.forward_gap = i->forward_gap;
		h.reverse_mm = i->reverse_mm; h.reverse_gap = i->reverse_gap;
		h.probe_mm = i->probe_mm; h.probe_gap = i->probe_gap;
		h.forward_clamp = i->forward_primer_clamp;
		h.reverse_clamp = i->reverse_primer_clamp;
		if (i->forward_oligo_str_index != INVALID_INDEX)
			copy_str(h.forward_oligo, sizeof(h.forward_oligo), keys[i->forward_oligo_str_index]);
		if (i->reverse_oligo_str_index != INVALID_INDEX)
			copy_str(h.reverse_oligo, sizeof(h.reverse_oligo), keys[i->reverse_oligo_str_index]);
		if (i->forward_align_str_index != INVALID_INDEX)
			copy_str(h.forward_align, sizeof(h.forward_align), inflate_dna_seq(keys[i->forward_align_str_index]));
		if (i->reverse_align_str_index != INVALID_INDEX)
			copy_str(h.reverse_align, sizeof(h.reverse_align), inflate_dna_seq(keys[i->reverse_align_str_index]));
		if (i->probe_align_str_index != INVALID_INDEX)
			copy_str(h.probe_align, sizeof(h.probe_align), inflate_dna_seq(keys[i->probe_align_str_index]));
		if (i->amplicon_str_index != INVALID_INDEX) {
			const std::string amp = inflate_dna_seq(keys[i->amplicon_str_index]);
			h.amplicon_len = (int)amp.size();
			// FNV-1a over the amplicon text: lets tests compare long amplicons cheaply
			uint64_t hash = 1469598103934665603ULL;
			for (size_t k = 0; k < amp.size(); ++k) { hash ^= (unsigned char)amp[k]; hash *= 1099511628211ULL; }
			h.amplicon_fnv = hash;
			copy_str(h.amplicon_head, sizeof(h.amplicon_head), amp);
		}
		g_hits.push_back(h);
	}
	return (long)g_hits.size();
	GUARD_END
}

extern "C" int ref_get_hits(ref_hit *out, long cap)
{
	const long n = std::min<long>(cap, (long)g_hits.size());
	if (n > 0) memcpy(out, g_hits.data(), n*sizeof(ref_hit));
	return (int)n;
}

// ---------------------------------------------------------------------------
// FASTA reader: the reference's own sequence_data class on a file
// ---------------------------------------------------------------------------
static sequence_data *g_fasta = nullptr;

extern "C" void ref_fasta_close()
{
	if (g_fasta) { g_fasta->close(); delete g_fasta; g_fasta = nullptr; }
}

extern "C" long ref_fasta_open(const char *path)
{
	GUARD_BEGIN
	ref_fasta_close();
	g_fasta = new sequence_data;
	g_fasta->open(path, std::vector<std::string>(), std::vector<std::string>());
	return (long)g_fasta->size();
	GUARD_END
}

extern "C" long ref_fasta_approx_len(long index)
{
	return g_fasta ? (long)g_fasta->approx_seq_len((size_t)index) : -1;
}

extern "C" long ref_fasta_read(long index, uint32_t start, uint32_t stop, int whole, uint8_t *out, long cap,
	char *defline, long defline_cap)
{
	GUARD_BEGIN
	if (!g_fasta) throw "no file open";
	std::pair<std::string, SEQPTR> seq;
	seq.second = NULL;
	const unsigned int n = whole ? g_fasta->read_bio_seq(seq, (unsigned int)index)
		: g_fasta->read_bio_seq(seq, (unsigned int)index, start, stop);
	for (unsigned int i = 0; i < n && (long)i < cap; ++i) out[i] = SEQ_START(seq.second)[i];
	if (defline && defline_cap > 0) {
		strncpy(defline, seq.first.c_str(), (size_t)defline_cap - 1);
		defline[defline_cap - 1] = '\0';
	}
	delete [] seq.second;
	return (long)n;
	GUARD_END
}

extern "C" void ref_seq_len_increment(uint32_t len, uint32_t max_len, uint32_t *delta, uint32_t *pieces)
{
	const std::pair<unsigned int, unsigned int> r = seq_len_increment(len, max_len);
	*delta = r.first;
	*pieces = r.second;
}

// ---------------------------------------------------------------------------
// Post-processing: the reference's select_best_match / uniquify_results / sort on one list
// ---------------------------------------------------------------------------
extern "C" long ref_finalize(const ref_post_hit *hits, long n, int best_match, int uniquify, int32_t *order_out)
{
	GUARD_BEGIN
	std::unordered_map<std::string, size_t> str_table;
	std::list<hybrid_sig> l;
	for (long i = 0; i < n; ++i) {
		const ref_post_hit &h = hits[i];
		hybrid_sig s;
		s.my_id(h.id);
		s.my_degen_id(h.degen_id);
		s.seq_id(h.seq_id);
		s.name_str_index = str_to_index("assay", str_table);
		if (h.has_primers) {
			// only the lengths of the oligo strings matter to uniquify_results (:1601-1602)
			s.forward_oligo_str_index = str_to_index(std::string((size_t)h.forward_len, 'A'), str_table);
			s.reverse_oligo_str_index = str_to_index(std::string((size_t)h.reverse_len, 'C'), str_table);
			s.amplicon_range = std::make_pair(h.amp_first, h.amp_last);
			s.forward_tm = h.forward_tm;
			s.reverse_tm = h.reverse_tm;
			s.forward_align_str_index = str_to_index(deflate_dna_seq(h.forward_align), str_table);
			s.reverse_align_str_index = str_to_index(deflate_dna_seq(h.reverse_align), str_table);
		}
		if (h.has_probe) {
			s.probe_oligo_str_index = str_to_index(std::string("GGGG"), str_table);
			s.probe_range = std::make_pair(h.probe_first, h.probe_last);
			s.probe_tm = h.probe_tm;
			s.probe_align_str_index = str_to_index(deflate_dna_seq(h.probe_align), str_table);
		}
		s.forward_degen = (int)i; // carries the input index through the list operations
		l.push_back(s);
	}
	const std::vector<std::string> index_table = ordered_keys(str_table);
	if (best_match) select_best_match(l);
	if (uniquify) uniquify_results(l, index_table);
	l.sort();
	long k = 0;
	for (std::list<hybrid_sig>::const_iterator i = l.begin(); i != l.end(); ++i) order_out[k++] = i->forward_degen;
	return k;
	GUARD_END
}

// ---------------------------------------------------------------------------
// Hairpins: the reference's approximate_tm_hairpin and its parameter tables
// ---------------------------------------------------------------------------
extern "C" int ref_hairpin(const char *query, float T, float na, ref_align_out *out)
{
	GUARD_BEGIN
	NucCruc *melt = get_melt(T, na);
	melt->dangle(false, false);
	melt->set_duplex(query); // tntblast_local.cpp:659
	const float tm = melt->approximate_tm_hairpin();
	memset(out, 0, sizeof(*out));
	out->tm = tm;
	out->valid = melt->curr_align.valid ? 1 : 0;
	out->dH = melt->curr_align.dH;
	out->dS = melt->curr_align.dS;
	out->dG = melt->curr_align.dH - T*melt->curr_align.dS;
	out->dp_dg = melt->curr_align.dp_dg;
	if (melt->curr_align.valid) {
		out->q_first = melt->curr_align.first_match.first;
		out->t_first = melt->curr_align.first_match.second;
		out->q_last = melt->curr_align.last_match.first;
		out->t_last = melt->curr_align.last_match.second;
		out->num_gap = (int)melt->curr_align.query_align.size(); // columns of the stem (incl. an attached end column)
	}
	return 0;
	GUARD_END
}

extern "C" int ref_hairpin_tables(float *hairpin_S, char (*loops)[8], float *special_H, float *special_S, int cap)
{
	GUARD_BEGIN
	NucCruc *melt = get_melt(310.15f, 0.05f);
	for (int i = 0; i <= REF_MAX_HAIRPIN; ++i) hairpin_S[i] = melt->param_hairpin_S[i];
	int n = 0;
	const char letters[] = "ACGT";
	for (int len = 5; len <= 6; ++len) {
		const int total = 1 << (2*len);
		for (int code = 0; code < total; ++code) {
			char text[8] = {0};
			for (int k = 0; k < len; ++k) text[k] = letters[(code >> (2*(len - 1 - k))) & 3];
			CircleBuffer<BASE::nucleic_acid, MAX_SEQUENCE_LENGTH> q;
			for (int k = 0; k < len; ++k) q.push_back(BASE::char_to_nucleic_acid(text[k]));
			const int idx = melt->find_loop_index(q, 0, (unsigned int)len);
			if (idx < 0) continue;
			if (n == cap) THROW("ref_hairpin_tables: more special loops than expected");
			strcpy(loops[n], text);
			special_H[n] = melt->param_hairpin_special_H[idx];
			special_S[n] = melt->param_hairpin_special_S[idx];
			++n;
		}
	}
	return n;
	GUARD_END
}
