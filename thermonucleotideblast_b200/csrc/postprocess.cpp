// Post-processing of hit lists on the host: what the reference driver does with
// the std::list<hybrid_sig> that amplicon()/padlock()/hybrid() return, restated over flat tnt_hit
// records so that hits of several engines (one per GPU) can be gathered into one result:
//
//   tntblast_local.cpp:635-654   hits that touch a cut edge of their fragment are dropped
//                                (start_overlap / stop_overlap, hybrid_sig.h:396-418), coordinates
//                                become record coordinates (offset_ranges :382-394), seq_id = record
//   tntblast_local.cpp:701-706   the lists of one assay id are spliced in front of each other
//   tntblast_local.cpp:918-930   select_best_match (tntblast_util.cpp:1482-1547) when asked for,
//                                uniquify_results (:1555-1755) when any record was cut, sort
//                                (hybrid_sig::operator<, hybrid_sig.h:328-355)
//
// The list operations are kept literal (std::list, its stable merge sort, the order in which
// uniquify_results visits and replaces its `valid` entries): which of two overlapping matches
// survives depends on them.  Only the alignment strings are not inflated / deflated: the engine
// hands them out as text.

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tntb200.h"

namespace tnt {

namespace {

struct Match {
	tnt_hit h;                 // record coordinates, target_id = record index
	int id, degen_id, seq;     // hybrid_sig::my_id(), my_degen_id(), seq_id()
	bool primers, probe;
	int fwd_len, rev_len;      // lengths of the oligos in the forward / reverse slot
	const char *fa, *ra, *pa;  // alignment text
	uint32_t block, index;

	float min_primer_tm() const { return std::max(0.0f, std::min(h.forward.tm, h.reverse.tm)); } // hybrid_sig.h:367-371
	float max_primer_tm() const { return std::max(h.forward.tm, h.reverse.tm); }
	std::pair<int, int> amplicon() const { return std::make_pair(h.amp_first, h.amp_last); }
	std::pair<int, int> probe_range() const { return std::make_pair(h.probe_first, h.probe_last); }
};

// hybrid_sig::operator< (hybrid_sig.h:328-355)
bool less_output(const Match &a, const Match &b)
{
	if (a.id == b.id) {
		if (a.min_primer_tm() == b.min_primer_tm()) {
			if (a.h.probe.tm == b.h.probe.tm) {
				if (a.max_primer_tm() == b.max_primer_tm()) return a.seq < b.seq;
				return a.max_primer_tm() > b.max_primer_tm();
			}
			return a.h.probe.tm > b.h.probe.tm;
		}
		return a.min_primer_tm() > b.min_primer_tm();
	}
	return a.id < b.id;
}

bool less_match(const Match &a, const Match &b) // sort_by_match, hybrid_sig.h:452-468
{
	if (a.id == b.id) return a.seq < b.seq;
	return a.id < b.id;
}

bool less_loc(const Match &a, const Match &b) // sort_by_loc, hybrid_sig.h:470-500
{
	if (a.id == b.id) {
		if (a.seq == b.seq) return a.primers ? a.amplicon() < b.amplicon() : a.probe_range() < b.probe_range();
		return a.seq < b.seq;
	}
	return a.id < b.id;
}

// top_strand (tntblast_util.cpp:1757-1776): the text between "5' " and " 3'"
std::string top_strand(const char *align)
{
	const char *s = std::strstr(align, "5' ");
	if (!s) throw std::runtime_error(":top_strand: Unable to parse alignment");
	s += 3;
	const char *e = std::strstr(align, " 3'");
	if (!e) throw std::runtime_error(":top_strand: Unable to parse alignment");
	return e >= s ? std::string(s, (size_t)(e - s)) : std::string();
}

// select_best_match (tntblast_util.cpp:1482-1547)
void select_best_match(std::list<Match> &l)
{
	if (l.empty()) return;
	l.sort(less_match);
	auto best = l.begin();
	auto cur = best;
	++cur;
	while (cur != l.end()) {
		if (cur->id == best->id && cur->seq == best->seq) {
			bool delete_cur = false;
			if (cur->primers) {
				if (cur->min_primer_tm() == best->min_primer_tm()) {
					if (cur->h.probe.tm < best->h.probe.tm) delete_cur = true;
					if (cur->max_primer_tm() < best->max_primer_tm()) delete_cur = true;
				}
				else if (cur->min_primer_tm() < best->min_primer_tm()) delete_cur = true;
			}
			else if (cur->h.probe.tm < best->h.probe.tm) delete_cur = true;
			if (delete_cur) cur = l.erase(cur);
			else {
				l.erase(best);
				best = cur;
				++cur;
			}
		}
		else {
			best = cur;
			++cur;
		}
	}
}

// uniquify_results (tntblast_util.cpp:1555-1755)
void uniquify(std::list<Match> &l)
{
	if (l.size() < 2) return;
	l.sort(less_loc);
	typedef std::list<Match>::iterator I;
	I start = l.begin(), stop = start;
	std::vector<I> reaper;
	enum State { NO_MATCH, A_CONTAINS_B, B_CONTAINS_A };
	for (;;) {
		if (stop != l.end() && start->id == stop->id && start->degen_id == stop->degen_id && start->seq == stop->seq) { ++stop; continue; }
		std::list<I> valid;
		for (I it = start; it != stop; ++it) {
			if (valid.empty()) { valid.push_back(it); continue; }
			const int fl = it->fwd_len/2, rl = it->rev_len/2;
			const std::string fa = it->primers ? top_strand(it->fa) : std::string();
			const std::string ra = it->primers ? top_strand(it->ra) : std::string();
			const std::string pa = it->primers ? std::string() : top_strand(it->pa);
			State status = NO_MATCH;
			for (auto v = valid.begin(); v != valid.end(); ++v) {
				State same = NO_MATCH;
				const Match &B = **v;
				if (it->primers) {
					const bool primers_overlap = std::abs(it->h.amp_first - B.h.amp_first) < fl && std::abs(it->h.amp_last - B.h.amp_last) < rl;
					if (primers_overlap) {
						const std::string vf = top_strand(B.fa), vr = top_strand(B.ra);
						if (it->h.amp_first <= B.h.amp_first && it->h.amp_last >= B.h.amp_last &&
							fa.find(vf) != std::string::npos && ra.find(vr) != std::string::npos) same = A_CONTAINS_B;
						else if (B.h.amp_first <= it->h.amp_first && B.h.amp_last >= it->h.amp_last &&
							vf.find(fa) != std::string::npos && vr.find(ra) != std::string::npos) same = B_CONTAINS_A;
						if (it->probe && B.probe && it->probe_range() != B.probe_range()) same = NO_MATCH;
					}
				}
				else {
					const std::string vp = top_strand(B.pa);
					if (it->h.probe_first <= B.h.probe_first && it->h.probe_last >= B.h.probe_last && pa.find(vp) != std::string::npos) same = A_CONTAINS_B;
					else if (B.h.probe_first <= it->h.probe_first && B.h.probe_last >= it->h.probe_last && vp.find(pa) != std::string::npos) same = B_CONTAINS_A;
				}
				if (same == NO_MATCH) continue;
				if (same == A_CONTAINS_B) { *v = it; status = A_CONTAINS_B; }
				else { status = B_CONTAINS_A; break; }
			}
			if (status == NO_MATCH) valid.push_front(it); // valid.splice(valid.begin(), valid_update)
		}
		for (I it = start; it != stop; ++it)
			if (std::find(valid.begin(), valid.end(), it) == valid.end()) reaper.push_back(it);
		start = stop;
		if (stop == l.end()) break;
	}
	while (!reaper.empty()) { l.erase(reaper.back()); reaper.pop_back(); }
}

thread_local std::string g_pp_error;

} // namespace

} // namespace tnt

using namespace tnt;

extern "C" {

const char *tnt_postprocess_error(void) { return g_pp_error.c_str(); }

void tnt_free(void *p) { std::free(p); }

int tnt_finalize_hits(const tnt_hit_block *blocks, size_t n_blocks, const tnt_assay *assays, int32_t n_assays,
	int32_t best_match, int32_t uniquify_mode, tnt_final_hit **out, size_t *n_out)
{
	try {
		if ((!blocks && n_blocks) || !out || !n_out || (!assays && n_assays)) throw std::runtime_error("null argument");
		*out = nullptr;
		*n_out = 0;
		// one list per assay id, like search_results[sig.my_id()] (tntblast_local.cpp:255, :701-706)
		std::map<int, std::list<Match> > by_id;
		bool any_cut = false;
		for (size_t b = 0; b < n_blocks; ++b) {
			const tnt_hit_block &blk = blocks[b];
			if ((blk.n_hits && (!blk.hits || !blk.arena)) || (blk.n_fragments && !blk.fragments)) throw std::runtime_error("null argument");
			for (size_t f = 0; f < blk.n_fragments; ++f)
				any_cut = any_cut || blk.fragments[f].start != 0 || blk.fragments[f].stop != blk.fragments[f].max_stop;
			size_t i = 0;
			while (i < blk.n_hits) {
				// hits of one (fragment, assay) call, spliced in front of the assay's list as a block
				size_t j = i;
				while (j < blk.n_hits && blk.hits[j].target_id == blk.hits[i].target_id && blk.hits[j].assay_index == blk.hits[i].assay_index) ++j;
				const tnt_hit &h0 = blk.hits[i];
				if (h0.target_id >= blk.n_fragments) throw std::runtime_error("hit refers to an unknown fragment");
				if (h0.assay_index < 0 || h0.assay_index >= n_assays) throw std::runtime_error("hit refers to an unknown assay");
				const tnt_fragment &fr = blk.fragments[h0.target_id];
				const tnt_assay &as = assays[h0.assay_index];
				const size_t len[3] = {as.forward ? std::strlen(as.forward) : 0, as.reverse ? std::strlen(as.reverse) : 0, as.probe ? std::strlen(as.probe) : 0};
				std::list<Match> local;
				for (size_t k = i; k < j; ++k) {
					Match m;
					m.h = blk.hits[k];
					m.primers = m.h.forward.oligo != TNT_OLIGO_NONE;
					m.probe = m.h.probe.oligo != TNT_OLIGO_NONE;
					const int first = m.primers ? m.h.amp_first : m.h.probe_first, last = m.primers ? m.h.amp_last : m.h.probe_last;
					if (fr.start != 0 && first <= 0) continue;                               // tntblast_local.cpp:638-642
					if (fr.stop != fr.max_stop && last >= (int)fr.len - 1) continue;          // :644-648
					if (m.primers) { m.h.amp_first += (int)fr.start; m.h.amp_last += (int)fr.start; }
					if (m.probe) { m.h.probe_first += (int)fr.start; m.h.probe_last += (int)fr.start; }
					m.h.target_id = fr.record;
					m.id = as.id;
					m.degen_id = m.h.assay_index;
					m.seq = (int)fr.record;
					m.fwd_len = m.primers ? (int)len[m.h.forward.oligo == TNT_OLIGO_R ? 1 : 0] : 0;
					m.rev_len = m.primers ? (int)len[m.h.reverse.oligo == TNT_OLIGO_F ? 0 : 1] : 0;
					m.fa = blk.arena + m.h.forward.align_off;
					m.ra = blk.arena + m.h.reverse.align_off;
					m.pa = blk.arena + m.h.probe.align_off;
					m.block = (uint32_t)b;
					m.index = (uint32_t)k;
					local.push_back(m);
				}
				std::list<Match> &dst = by_id[as.id];
				dst.splice(dst.begin(), local);
				i = j;
			}
		}
		const bool do_uniquify = uniquify_mode < 0 ? any_cut : uniquify_mode != 0;
		// the lists of different assay ids are independent: a few host threads when there is much to do
		{
			std::vector<std::list<Match> *> lists;
			size_t total = 0;
			for (auto &kv : by_id) { lists.push_back(&kv.second); total += kv.second.size(); }
			const unsigned nthreads = total < 50000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
			std::vector<std::string> errors(nthreads);
			auto work = [&](unsigned t) {
				try {
					for (size_t i = t; i < lists.size(); i += nthreads) {
						std::list<Match> &l = *lists[i];
						if (l.empty()) continue;
						if (best_match) select_best_match(l);
						if (do_uniquify) uniquify(l);
						l.sort(less_output);
					}
				}
				catch (const std::exception &ex) { errors[t] = ex.what(); }
			};
			if (nthreads == 1) work(0);
			else {
				std::vector<std::thread> pool;
				for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
				for (std::thread &t : pool) t.join();
			}
			for (const std::string &err : errors) if (!err.empty()) throw std::runtime_error(err);
		}
		std::vector<tnt_final_hit> result;
		for (auto &kv : by_id) {
			std::list<Match> &l = kv.second;
			for (const Match &m : l) {
				tnt_final_hit f;
				f.block = m.block;
				f.index = m.index;
				f.hit = m.h;
				result.push_back(f);
			}
		}
		if (!result.empty()) {
			*out = (tnt_final_hit *)std::malloc(result.size()*sizeof(tnt_final_hit));
			if (!*out) throw std::runtime_error("out of memory");
			std::memcpy(*out, result.data(), result.size()*sizeof(tnt_final_hit));
		}
		*n_out = result.size();
		return 0;
	}
	catch (const std::exception &ex) { g_pp_error = ex.what(); return -1; }
	catch (...) { g_pp_error = "unknown error"; return -1; }
}

} // extern "C"
