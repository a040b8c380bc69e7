// tntb200 shim: the reference's search call surface, served by the B200 engine through the C ABI.
//
// TEST / INTEGRATION INFRASTRUCTURE.  Compiles only where the reference headers are present
// (-I/root/reference); it is the concrete form of the binding INTEGRATION.md describes.
//
// What it is: definitions of
//     amplicon()   tntblast.h:409-433   (reference body: amplicon_search.cpp:58-677)
//     padlock()    tntblast.h:437-451   (padlock_search.cpp:62-361)
//     hybrid()     tntblast.h:465-477   (probe_search.cpp:67-230)
// with the reference's own signatures.  Linked INSTEAD of amplicon_search.o, padlock_search.o,
// probe_search.o and bind_oligo.o, every other object of the reference stays as it is: the
// unmodified driver (tntblast_local.cpp:554-626) calls these functions once per (fragment, assay)
// and gets its std::list<hybrid_sig> back, filled exactly like the reference fills it
// (amplicon_search.cpp:447-555,565-671; padlock_search.cpp:155-222,287-355; probe_search.cpp:103-151,
// 179-227).  Hairpin / dimer temperatures, the truncation filter, uniquify, sorting and printing
// remain the driver's (tntblast_local.cpp:635-1280).
//
// Two ways a call is answered:
//  * from the batch table: tntb200_prefetch() (called by tests/shim/tntblast_gpu_main.cpp before
//    local_main) parses the same command line with the reference's own Options class, walks the
//    database with the reference's sequence_data reader and the driver's fragment rule
//    (tntblast_local.cpp:282-289,448-468,510), registers every fragment with
//    tnt_engine_add_target and searches all assays against all fragments in a few
//    tnt_engine_search calls.  Hits are filed under (fragment checksum, assay); a later
//    amplicon()/padlock()/hybrid() call recognises its fragment by the same checksum.
//  * directly: any call the table cannot answer (no prefetch, other thresholds, another assay)
//    uploads its fragment and searches its single assay on the spot (one engine, serialised).
// Nothing falls back to the reference's CPU search: what the engine refuses is thrown as
// `const char *` like every reference error (throw.h:16-17).

#include "tntblast.h"
#include "options.h"
#include "hybrid_sig.h"
#include "compress.h"
#include "throw.h"

#include <getopt.h>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <atomic>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tntb200.h"

using namespace std;

namespace {

struct StoredHit {
	tnt_hit h;
	string forward_align, reverse_align, probe_align;
	string sequence;     // tnt_engine_hit_sequences
};

struct FragEntry {
	uint32_t len = 0;
	// hits of this fragment, grouped by assay index (ascending), in the engine's (== reference's) order
	vector<pair<int, vector<StoredHit> > > by_assay;
};

struct AssayKey { string F, R, P; int fd, rd, pd; };

struct ShimState {
	mutex mu;                      // the direct path uses one engine for all driver threads
	int word_size = 7;
	tnt_engine *eng = nullptr;     // created on first use
	tnt_engine_params prm{};
	bool have_prm = false;
	// batch table
	bool prefetched = false;
	tnt_search_options pre_opt{};
	vector<AssayKey> pre_assays;
	unordered_map<uint64_t, FragEntry> table;
	// what the engine currently holds on the direct path
	uint64_t direct_frag = 0;
	bool direct_has_frag = false;
	// counters (reported by tntb200_report)
	atomic<size_t> calls_table{0};
	size_t calls_direct = 0, fragments = 0, batches = 0;
	unsigned long long bases = 0, alignments = 0, hits = 0;
	double search_ms = 0.0;
	string last_error; // storage for rethrown messages
};

ShimState g;

[[noreturn]] void rethrow_engine_error()
{
	// same convention as the reference: throw a C string (caught in local_main, tntblast_local.cpp:1377-1391)
	g.last_error = tnt_last_error();
	throw g.last_error.c_str();
}

inline void check(int rc) { if (rc < 0) rethrow_engine_error(); }

// 64-bit checksum of the fragment bytes: the identity of a fragment between the batch pass and
// the per-call lookups (the call surface carries no target index)
uint64_t fragment_checksum(SEQPTR seq)
{
	const uint32_t n = SEQ_SIZE(seq);
	const unsigned char *p = SEQ_START(seq);
	uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
	uint32_t i = 0;
	for (; i + 8 <= n; i += 8) {
		uint64_t w;
		memcpy(&w, p + i, 8);
		h = (h ^ w)*0xFF51AFD7ED558CCDull;
		h ^= h >> 29;
	}
	uint64_t tail = 0;
	for (uint32_t k = 0; i + k < n; ++k) tail |= (uint64_t)p[i + k] << (8*k);
	h = (h ^ tail)*0xC4CEB9FE1A85EC53ull;
	h ^= h >> 32;
	return h;
}

void params_from(const Options &opt)
{
	tnt_engine_params p{};
	p.target_T = opt.target_t;
	p.salt = opt.salt;
	p.dangle5 = opt.allow_dangle_5;
	p.dangle3 = opt.allow_dangle_3;
	p.dinkelbach = opt.use_dinkelbach;
	p.word_size = opt.hash_word_size;
	g.prm = p;
	g.have_prm = true;
}

void ensure_engine(const NucCruc *melt, int word_size)
{
	if (g.eng) return;
	if (!g.have_prm) {
		// no command line seen (the shim linked under a stock main): the per-thread NucCruc / DNAHash the
		// driver built from its options (tntblast_local.cpp:345,363-367) carry all but the dangling-end
		// switches, which NucCruc does not expose
		tnt_engine_params p{};
		p.target_T = melt->temperature();
		p.salt = melt->salt();
		p.dangle5 = getenv("TNTB200_DANGLE5") ? 1 : 0;
		p.dangle3 = getenv("TNTB200_DANGLE3") ? 1 : 0;
		p.dinkelbach = melt->dinkelbach();
		p.word_size = word_size;
		g.prm = p;
		g.have_prm = true;
	}
	g.prm.device = 0;
	if (const char *d = getenv("TNTB200_DEVICE")) g.prm.device = atoi(d);
	check(tnt_engine_create(&g.prm, &g.eng));
}

vector<StoredHit> collect_hits(size_t &n_out, vector<uint32_t> &targets, vector<int> &assays)
{
	const tnt_hit *hits = nullptr;
	size_t n = 0;
	const char *arena = nullptr;
	check(tnt_engine_get_hits(g.eng, &hits, &n, &arena, nullptr));
	const char *text = nullptr;
	const uint64_t *off = nullptr;
	size_t nseq = 0;
	check(tnt_engine_hit_sequences(g.eng, &text, &off, &nseq));
	if (nseq != n) throw "tntb200 shim: hit / sequence count mismatch";
	vector<StoredHit> out(n);
	targets.resize(n);
	assays.resize(n);
	for (size_t i = 0; i < n; ++i) {
		StoredHit &s = out[i];
		s.h = hits[i];
		s.forward_align = arena + hits[i].forward.align_off;
		s.reverse_align = arena + hits[i].reverse.align_off;
		s.probe_align = arena + hits[i].probe.align_off;
		s.sequence.assign(text + off[i], (size_t)(off[i + 1] - off[i] - 1));
		targets[i] = hits[i].target_id;
		assays[i] = hits[i].assay_index;
	}
	n_out = n;
	return out;
}

// ---- hybrid_sig construction: field by field what the reference assigns -----------------------
void fill_primers(hybrid_sig &tmp, const hybrid_sig &sig, const StoredHit &s, bool padlock_format,
	unordered_map<string, size_t> &str_table)
{
	const tnt_hit &h = s.h;
	if (!padlock_format) {
		// amplicon_search.cpp:453-463 (a single primer can make both ends)
		if (h.forward.oligo == TNT_OLIGO_R) tmp.forward_oligo_str_index = sig.reverse_oligo_str_index;
		if (h.reverse.oligo == TNT_OLIGO_F) tmp.reverse_oligo_str_index = sig.forward_oligo_str_index;
	}
	tmp.primer_strand = (int8_t)h.primer_strand;
	tmp.amplicon_range = make_pair(h.amp_first, h.amp_last);
	tmp.forward_tm = h.forward.tm;   tmp.reverse_tm = h.reverse.tm;
	tmp.forward_dH = h.forward.dH;   tmp.reverse_dH = h.reverse.dH;
	tmp.forward_dS = h.forward.dS;   tmp.reverse_dS = h.reverse.dS;
	tmp.forward_mm = (int8_t)h.forward.num_mm;   tmp.reverse_mm = (int8_t)h.reverse.num_mm;
	tmp.forward_gap = (int8_t)h.forward.num_gap; tmp.reverse_gap = (int8_t)h.reverse.num_gap;
	tmp.forward_primer_clamp = (int8_t)h.forward_clamp;
	tmp.reverse_primer_clamp = (int8_t)h.reverse_clamp;
	tmp.forward_align_str_index = str_to_index(deflate_dna_seq(s.forward_align), str_table);
	tmp.reverse_align_str_index = str_to_index(deflate_dna_seq(s.reverse_align), str_table);
}

void fill_probe(hybrid_sig &tmp, const StoredHit &s, unordered_map<string, size_t> &str_table)
{
	const tnt_hit &h = s.h;
	tmp.probe_range = make_pair(h.probe_first, h.probe_last);
	tmp.probe_tm = h.probe.tm;
	tmp.probe_dH = h.probe.dH;
	tmp.probe_dS = h.probe.dS;
	tmp.probe_mm = (int8_t)h.probe.num_mm;
	tmp.probe_gap = (int8_t)h.probe.num_gap;
	tmp.probe_strand = (int8_t)h.probe_strand;
	tmp.probe_align_str_index = str_to_index(deflate_dna_seq(s.probe_align), str_table);
}

// the shared body of the three entry points
struct CallArgs {
	const DNAHash *hash;
	const pair<string, SEQPTR> *seq;
	const hybrid_sig *sig;
	NucCruc *melt;
	tnt_search_options o;
	int mask_options;
	float min_primer_tm, min_probe_tm;   // mask_binding_sites arguments
	const vector<string> *oligo_table;
	unordered_map<string, size_t> *str_table;
};

AssayKey key_of(const hybrid_sig &sig, const vector<string> &oligo_table)
{
	AssayKey k;
	k.F = index_to_str(sig.forward_oligo_str_index, oligo_table);
	k.R = index_to_str(sig.reverse_oligo_str_index, oligo_table);
	k.P = index_to_str(sig.probe_oligo_str_index, oligo_table);
	k.fd = sig.forward_degen; k.rd = sig.reverse_degen; k.pd = sig.probe_degen;
	return k;
}

bool same_assay(const AssayKey &a, const AssayKey &b)
{
	return a.F == b.F && a.R == b.R && a.P == b.P && a.fd == b.fd && a.rd == b.rd && a.pd == b.pd;
}

// Does the batch pass (options `pre`) answer a call made with options `call`?  A probe-only assay in a
// PCR run goes through hybrid() (tntblast_local.cpp:612-625): only the probe-side fields matter there.
bool same_search(const tnt_search_options &pre, const tnt_search_options &call)
{
	if (call.assay_format == TNT_ASSAY_PROBE && pre.assay_format == TNT_ASSAY_PCR) {
		return pre.probe_strand == call.probe_strand && pre.min_probe_tm == call.min_probe_tm && pre.max_probe_tm == call.max_probe_tm &&
			pre.min_probe_dg == call.min_probe_dg && pre.max_probe_dg == call.max_probe_dg &&
			pre.probe_clamp_5 == call.probe_clamp_5 && pre.probe_clamp_3 == call.probe_clamp_3 &&
			pre.max_gap == call.max_gap && pre.max_mismatch == call.max_mismatch && pre.max_poly_degen == call.max_poly_degen &&
			pre.target_strand == call.target_strand;
	}
	if (call.assay_format == TNT_ASSAY_PCR && pre.assay_format == TNT_ASSAY_PCR) {
		// amplicon() has no target-strand argument (primer assays search both strands)
		tnt_search_options a = pre, b = call;
		a.target_strand = b.target_strand = 0;
		return memcmp(&a, &b, sizeof(a)) == 0;
	}
	return memcmp(&pre, &call, sizeof(pre)) == 0;
}

tnt_assay c_assay(const AssayKey &k, int id)
{
	tnt_assay a{};
	a.id = id;
	a.forward = k.F.empty() ? nullptr : k.F.c_str();
	a.reverse = k.R.empty() ? nullptr : k.R.c_str();
	a.probe = k.P.empty() ? nullptr : k.P.c_str();
	a.forward_degen = k.fd; a.reverse_degen = k.rd; a.probe_degen = k.pd;
	return a;
}

list<hybrid_sig> build_list(const CallArgs &c, const vector<StoredHit> &hits)
{
	list<hybrid_sig> out;
	const bool padlock_format = c.o.assay_format == TNT_ASSAY_PADLOCK || c.o.assay_format == TNT_ASSAY_MIPS;
	for (const StoredHit &s : hits) {
		hybrid_sig tmp = *c.sig;                                  // id, name, oligo indices, degeneracies
		const bool primers = s.h.forward.oligo != TNT_OLIGO_NONE;
		if (primers) {
			fill_primers(tmp, *c.sig, s, padlock_format, *c.str_table);
			tmp.amplicon_def_str_index = str_to_index(c.seq->first, *c.str_table);
			string amp = s.sequence;
			if (!padlock_format) {
				// amplicon_search.cpp:539-542 / :663-666 (the probe fields are still unset at this point)
				mask_binding_sites(amp, tmp, c.mask_options, c.min_primer_tm, c.min_probe_tm, *c.melt,
					c.o.forward_primer_strand, c.o.reverse_primer_strand, c.o.probe_strand, *c.oligo_table);
			}
			tmp.amplicon_str_index = str_to_index(deflate_dna_seq(amp), *c.str_table);
			if (s.h.probe.oligo != TNT_OLIGO_NONE) fill_probe(tmp, s, *c.str_table);
		}
		else {
			fill_probe(tmp, s, *c.str_table);
			tmp.amplicon_def_str_index = str_to_index(c.seq->first, *c.str_table);
			tmp.amplicon_str_index = str_to_index(deflate_dna_seq(s.sequence), *c.str_table);
		}
		out.push_back(tmp);
	}
	return out;
}

list<hybrid_sig> serve(const CallArgs &c)
{
	const uint64_t sum = fragment_checksum(c.seq->second);
	const AssayKey key = key_of(*c.sig, *c.oligo_table);
	const int degen_id = c.sig->my_degen_id();

	if (g.prefetched && same_search(g.pre_opt, c.o) && degen_id >= 0 &&
		(size_t)degen_id < g.pre_assays.size() && same_assay(g.pre_assays[(size_t)degen_id], key)) {
		const auto it = g.table.find(sum);
		if (it != g.table.end() && it->second.len == SEQ_SIZE(c.seq->second)) {
			++g.calls_table;
			for (const auto &pa : it->second.by_assay)
				if (pa.first == degen_id) return build_list(c, pa.second);
			return list<hybrid_sig>();
		}
	}

	// direct path: this fragment, this assay, now
	lock_guard<mutex> lk(g.mu);
	++g.calls_direct;
	ensure_engine(c.melt, (int)c.hash->word_size());
	if (!g.direct_has_frag || g.direct_frag != sum) {
		check(tnt_engine_clear_targets(g.eng));
		uint32_t id = 0;
		check(tnt_engine_add_target(g.eng, SEQ_START(c.seq->second), SEQ_SIZE(c.seq->second), &id));
		g.direct_frag = sum;
		g.direct_has_frag = true;
	}
	const tnt_assay a = c_assay(key, c.sig->my_id());
	check(tnt_engine_set_assays(g.eng, &a, 1));
	check(tnt_engine_search(g.eng, &c.o));
	size_t n = 0;
	vector<uint32_t> t;
	vector<int> as;
	const vector<StoredHit> hits = collect_hits(n, t, as);
	return build_list(c, hits);
}

tnt_search_options options_of(int format, float fps, float rps, float ps,
	float min_primer_tm, float max_primer_tm, float min_primer_dg, float max_primer_dg,
	float min_probe_tm, float max_probe_tm, float min_probe_dg, float max_probe_dg,
	unsigned primer_clamp, int min_max_primer_clamp, unsigned probe_clamp_5, unsigned probe_clamp_3,
	unsigned max_gap, unsigned max_mismatch, unsigned max_poly_degen, unsigned max_len, bool single_primer_pcr,
	int target_strand)
{
	tnt_search_options o;
	memset(&o, 0, sizeof(o)); // compared with memcmp
	o.assay_format = format;
	o.forward_primer_strand = fps; o.reverse_primer_strand = rps; o.probe_strand = ps;
	o.min_primer_tm = min_primer_tm; o.max_primer_tm = max_primer_tm;
	o.min_primer_dg = min_primer_dg; o.max_primer_dg = max_primer_dg;
	o.min_probe_tm = min_probe_tm; o.max_probe_tm = max_probe_tm;
	o.min_probe_dg = min_probe_dg; o.max_probe_dg = max_probe_dg;
	o.primer_clamp = primer_clamp; o.min_max_primer_clamp = min_max_primer_clamp;
	o.probe_clamp_5 = probe_clamp_5; o.probe_clamp_3 = probe_clamp_3;
	o.max_gap = max_gap; o.max_mismatch = max_mismatch; o.max_poly_degen = max_poly_degen;
	o.max_len = max_len; o.single_primer_pcr = single_primer_pcr ? 1 : 0;
	o.target_strand = target_strand;
	return o;
}

// the options the driver will pass for every call (tntblast_local.cpp:566-624), from its Options
tnt_search_options options_of(const Options &opt)
{
	const float fps = opt.asymmetric_strand_ratio*opt.primer_strand; // tntblast_local.cpp:232-234
	switch (opt.assay_format) {
	case ASSAY_PCR:
		return options_of(TNT_ASSAY_PCR, fps, opt.primer_strand, opt.probe_strand,
			opt.min_primer_tm, opt.max_primer_tm, opt.min_primer_dg, opt.max_primer_dg,
			opt.min_probe_tm, opt.max_probe_tm, opt.min_probe_dg, opt.max_probe_dg,
			opt.primer_clamp, opt.min_max_primer_clamp, opt.probe_clamp_5, opt.probe_clamp_3,
			opt.max_gap, opt.max_mismatch, opt.max_poly_degen, opt.max_len, opt.single_primer_pcr, opt.target_strand);
	case ASSAY_PADLOCK:
	case ASSAY_MIPS: {
		const int max_len = opt.assay_format == ASSAY_MIPS ? opt.max_len : 0; // tntblast_local.cpp:593,607
		return options_of(max_len > 0 ? TNT_ASSAY_MIPS : TNT_ASSAY_PADLOCK, fps, opt.primer_strand, 0.0f,
			0.0f, 0.0f, 0.0f, 0.0f, opt.min_probe_tm, opt.max_probe_tm, opt.min_probe_dg, opt.max_probe_dg,
			0, -1, opt.probe_clamp_5, opt.probe_clamp_3, opt.max_gap, opt.max_mismatch, opt.max_poly_degen,
			(unsigned)max_len, false, opt.target_strand);
	}
	default:
		return options_of(TNT_ASSAY_PROBE, 0.0f, 0.0f, opt.probe_strand,
			0.0f, 0.0f, 0.0f, 0.0f, opt.min_probe_tm, opt.max_probe_tm, opt.min_probe_dg, opt.max_probe_dg,
			0, -1, opt.probe_clamp_5, opt.probe_clamp_3, opt.max_gap, opt.max_mismatch, opt.max_poly_degen,
			0, false, opt.target_strand);
	}
}

struct CoutSilencer {
	streambuf *old_out, *old_err;
	stringstream sink;
	CoutSilencer() : old_out(cout.rdbuf(sink.rdbuf())), old_err(cerr.rdbuf(sink.rdbuf())) {}
	~CoutSilencer() { cout.rdbuf(old_out); cerr.rdbuf(old_err); }
};

} // namespace

// ---------------------------------------------------------------------------------------------
// The reference's call surface
// ---------------------------------------------------------------------------------------------
list<hybrid_sig> amplicon(DNAHash &m_hash, const pair<string, SEQPTR> &m_seq,
	const hybrid_sig &m_sig, NucCruc &m_melt,
	unordered_map<BindCacheKey, BindCacheValue> &, unordered_map<BindCacheKey, BindCacheValue> &,
	const float &m_forward_primer_strand, const float &m_reverse_primer_strand, const float &m_probe_strand,
	const float &m_min_primer_tm, const float &m_max_primer_tm,
	const float &m_min_primer_dg, const float &m_max_primer_dg,
	const float &m_min_probe_tm, const float &m_max_probe_tm,
	const float &m_min_probe_dg, const float &m_max_probe_dg,
	const unsigned int &m_primer_clamp, const int &m_min_primer_clamp,
	const unsigned int &m_probe_clamp_5, const unsigned int &m_probe_clamp_3,
	const unsigned int &m_max_gap, const unsigned int &m_max_mismatch, const unsigned int &m_max_poly_degen,
	const unsigned int &m_max_amplicon_len, const bool &m_single_primer_pcr, const int &m_mask_options,
	const vector<string> &m_oligo_table, unordered_map<string, size_t> &m_str_table)
{
	CallArgs c;
	c.hash = &m_hash; c.seq = &m_seq; c.sig = &m_sig; c.melt = &m_melt;
	c.o = options_of(TNT_ASSAY_PCR, m_forward_primer_strand, m_reverse_primer_strand, m_probe_strand,
		m_min_primer_tm, m_max_primer_tm, m_min_primer_dg, m_max_primer_dg,
		m_min_probe_tm, m_max_probe_tm, m_min_probe_dg, m_max_probe_dg,
		m_primer_clamp, m_min_primer_clamp, m_probe_clamp_5, m_probe_clamp_3,
		m_max_gap, m_max_mismatch, m_max_poly_degen, m_max_amplicon_len, m_single_primer_pcr, Seq_strand_both);
	c.mask_options = m_mask_options;
	c.min_primer_tm = m_min_primer_tm; c.min_probe_tm = m_min_probe_tm;
	c.oligo_table = &m_oligo_table; c.str_table = &m_str_table;
	return serve(c);
}

list<hybrid_sig> padlock(DNAHash &m_hash, const pair<string, SEQPTR> &m_seq,
	const hybrid_sig &m_sig, NucCruc &m_melt,
	unordered_map<BindCacheKey, BindCacheValue> &, unordered_map<BindCacheKey, BindCacheValue> &,
	const float &m_forward_primer_strand, const float &m_reverse_primer_strand,
	const float &m_min_primer_tm, const float &m_max_primer_tm,
	const float &m_min_primer_dg, const float &m_max_primer_dg,
	const unsigned int &m_probe_clamp_5, const unsigned int &m_probe_clamp_3,
	const unsigned int &m_max_gap, const unsigned int &m_max_mismatch, const unsigned int &m_max_poly_degen,
	const int &m_target_strand, const int &m_max_len,
	const vector<string> &m_oligo_table, unordered_map<string, size_t> &m_str_table)
{
	CallArgs c;
	c.hash = &m_hash; c.seq = &m_seq; c.sig = &m_sig; c.melt = &m_melt;
	// the driver calls padlock() with max_len 0 for PADLOCK and opt.max_len for MIPS
	// (tntblast_local.cpp:584-609); the two formats differ in nothing else
	c.o = options_of(m_max_len > 0 ? TNT_ASSAY_MIPS : TNT_ASSAY_PADLOCK, m_forward_primer_strand, m_reverse_primer_strand, 0.0f,
		0.0f, 0.0f, 0.0f, 0.0f, m_min_primer_tm, m_max_primer_tm, m_min_primer_dg, m_max_primer_dg,
		0, -1, m_probe_clamp_5, m_probe_clamp_3, m_max_gap, m_max_mismatch, m_max_poly_degen,
		(unsigned)m_max_len, false, m_target_strand);
	c.mask_options = 0;
	c.min_primer_tm = c.min_probe_tm = 0.0f;
	c.oligo_table = &m_oligo_table; c.str_table = &m_str_table;
	return serve(c);
}

list<hybrid_sig> hybrid(DNAHash &m_hash, const pair<string, SEQPTR> &m_seq,
	const hybrid_sig &m_sig, NucCruc &m_melt, const float &m_probe_strand,
	const float &m_min_probe_tm, const float &m_max_probe_tm,
	const float &m_min_probe_dg, const float &m_max_probe_dg,
	const unsigned int &m_probe_clamp_5, const unsigned int &m_probe_clamp_3,
	const unsigned int &m_max_gap, const unsigned int &m_max_mismatch, const unsigned int &m_max_poly_degen,
	const int &m_target_strand,
	const vector<string> &m_oligo_table, unordered_map<string, size_t> &m_str_table)
{
	CallArgs c;
	c.hash = &m_hash; c.seq = &m_seq; c.sig = &m_sig; c.melt = &m_melt;
	c.o = options_of(TNT_ASSAY_PROBE, 0.0f, 0.0f, m_probe_strand,
		0.0f, 0.0f, 0.0f, 0.0f, m_min_probe_tm, m_max_probe_tm, m_min_probe_dg, m_max_probe_dg,
		0, -1, m_probe_clamp_5, m_probe_clamp_3, m_max_gap, m_max_mismatch, m_max_poly_degen,
		0, false, m_target_strand);
	c.mask_options = 0;
	c.min_primer_tm = c.min_probe_tm = 0.0f;
	c.oligo_table = &m_oligo_table; c.str_table = &m_str_table;
	return serve(c);
}

// ---------------------------------------------------------------------------------------------
// Batch pass
// ---------------------------------------------------------------------------------------------
// Everything up to the search loop exactly as local_main prepares it (tntblast_local.cpp:39-175),
// through the reference's own functions; then all fragments x all assays on the GPU.  Returns false
// (and leaves the table empty) when the command line is not a search the batch pass understands --
// the driver then reports its own errors and every call takes the direct path.
bool tntb200_prefetch(int argc, char *argv[])
{
	Options opt;
	unordered_map<string, size_t> str_table;
	vector<string> index_table;
	try {
		CoutSilencer quiet; // the driver prints these messages itself, once
		try { opt.parse(argc, argv); }
		catch (...) { optind = 0; return false; }
		optind = 0; // glibc: re-initialise getopt for the driver's own parse
		if (opt.print_usage) return false;
		if (opt.input_filename != "")
			read_input_file(opt.input_filename, opt.sig_list, opt.ignore_probe, (opt.assay_format == ASSAY_PROBE), str_table);
		index_table = ordered_keys(str_table);
		if (opt.multiplex) {
			opt.sig_list = multiplex_expansion(opt.sig_list, opt.assay_format, index_table, str_table);
			index_table = ordered_keys(str_table);
		}
		opt.sig_list = expand_degenerate_signatures(opt.sig_list, opt.degen_rescale_ct, index_table, str_table);
		index_table = ordered_keys(str_table);
		opt.validate_search_threshold();
		if (opt.sig_list.empty()) return false;
	}
	catch (...) { optind = 0; return false; }
	params_from(opt); // the direct path needs them as well
	if (getenv("TNTB200_NO_PREFETCH")) return false;
	if (opt.assay_format != ASSAY_PCR && opt.assay_format != ASSAY_PROBE && opt.assay_format != ASSAY_PADLOCK &&
		opt.assay_format != ASSAY_MIPS) return false;

	const unsigned int max_product_length = opt.max_product_length(index_table) + 2; // tntblast_local.cpp:174
	sequence_data seq_file;
	{
		CoutSilencer quiet;
		try {
			seq_file.open(opt.dbase_filename != "" ? opt.dbase_filename : opt.local_dbase_filename, opt.blast_include, opt.blast_exclude);
		}
		catch (...) { return false; }
	}
	const size_t num_seq = seq_file.size();
	if (num_seq == 0) return false;

	lock_guard<mutex> lk(g.mu);
	ensure_engine(nullptr, opt.hash_word_size);

	// assays in sig_list order: index == my_degen_id() (expand_degenerate_signatures renumbers them)
	g.pre_assays.clear();
	vector<tnt_assay> c_assays;
	for (size_t i = 0; i < opt.sig_list.size(); ++i) {
		if (opt.sig_list[i].my_degen_id() != (int)i) return false;
		g.pre_assays.push_back(key_of(opt.sig_list[i], index_table));
	}
	for (size_t i = 0; i < g.pre_assays.size(); ++i) c_assays.push_back(c_assay(g.pre_assays[i], opt.sig_list[i].my_id()));
	check(tnt_engine_set_assays(g.eng, c_assays.data(), (int32_t)c_assays.size()));
	g.pre_opt = options_of(opt);

	// bases registered per engine search: bounded by what one GPU keeps resident comfortably
	unsigned long long batch_limit = 8000000000ull;
	if (const char *v = getenv("TNTB200_BATCH_BASES")) batch_limit = strtoull(v, nullptr, 10);

	vector<uint64_t> batch_sum;      // checksum of every fragment of the open batch, by target id
	unsigned long long batch_bases = 0;
	auto run_batch = [&]() {
		if (batch_sum.empty()) return;
		check(tnt_engine_search(g.eng, &g.pre_opt));
		tnt_stats st;
		check(tnt_engine_get_stats(g.eng, &st));
		g.alignments += st.alignments;
		g.search_ms += st.total_ms;
		size_t n = 0;
		vector<uint32_t> targets;
		vector<int> assays;
		vector<StoredHit> hits = collect_hits(n, targets, assays);
		g.hits += n;
		for (size_t i = 0; i < n; ++i) {
			// hits arrive ordered by (fragment, assay); a fragment is registered once (identical
			// fragments share their entry)
			FragEntry &fe = g.table[batch_sum[targets[i]]];
			if (fe.by_assay.empty() || fe.by_assay.back().first != assays[i])
				fe.by_assay.push_back(make_pair(assays[i], vector<StoredHit>()));
			fe.by_assay.back().second.push_back(std::move(hits[i]));
		}
		check(tnt_engine_clear_targets(g.eng));
		batch_sum.clear();
		batch_bases = 0;
		++g.batches;
	};

	// the driver's fragment walk (tntblast_local.cpp:282-289, 448-468) and read (:510)
	pair<string, SEQPTR> bio_seq = make_pair(string(), SEQPTR(NULL));
	const DNAHash probe_hash(opt.hash_word_size);
	for (unsigned int target = 0; target < num_seq; ++target) {
		const unsigned int len = seq_file.approx_seq_len(target);
		const unsigned int max_stop = len - 1;
		const unsigned int delta = seq_len_increment(len, opt.fragment_target_threshold).first;
		unsigned int start = 0, stop = delta;
		while (true) {
			if (bio_seq.second != NULL) { delete [] bio_seq.second; bio_seq.second = NULL; }
			const unsigned int target_len = seq_file.read_bio_seq(bio_seq, target, start, stop + max_product_length);
			if (target_len >= probe_hash.min_sequence_size()) { // shorter sequences are skipped (:513-527)
				const uint64_t sum = fragment_checksum(bio_seq.second);
				FragEntry &fe = g.table[sum];
				const bool duplicate = fe.len == SEQ_SIZE(bio_seq.second) && fe.len != 0;
				fe.len = SEQ_SIZE(bio_seq.second);
				if (!duplicate) {
					uint32_t id = 0;
					check(tnt_engine_add_target(g.eng, SEQ_START(bio_seq.second), SEQ_SIZE(bio_seq.second), &id));
					if (id != batch_sum.size()) throw "tntb200 shim: unexpected fragment id";
					batch_sum.push_back(sum);
					batch_bases += target_len;
					g.bases += target_len;
					++g.fragments;
					if (batch_bases >= batch_limit) run_batch();
				}
			}
			if (stop == max_stop) break;
			start = stop + 1;
			stop = min(stop + delta, max_stop);
		}
	}
	if (bio_seq.second != NULL) { delete [] bio_seq.second; bio_seq.second = NULL; }
	run_batch();
	g.prefetched = true;
	return true;
}

// one line on stderr: how the calls of this run were answered (the tests require calls_direct == 0
// for the batch runs, so a silent detour cannot pass)
void tntb200_report()
{
	fprintf(stderr, "[tntb200] prefetched=%d fragments=%zu bases=%llu batches=%zu alignments=%llu hits=%llu device_ms=%.3f "
		"calls_from_table=%zu calls_direct=%zu\n", (int)g.prefetched, g.fragments, g.bases, g.batches, g.alignments, g.hits,
		g.search_ms, (size_t)g.calls_table, g.calls_direct);
	if (g.eng) { tnt_engine_destroy(g.eng); g.eng = nullptr; }
}
