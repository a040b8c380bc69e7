// Stage C on the host: bound oligos -> hits.  O(#bound sites) work; the candidates themselves
// never leave the GPU.

#include "assemble.h"

#include <algorithm>
#include <cstring>
#include <list>
#include <map>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>

#include "thermo.h"

namespace tnt {

namespace {

const char kBaseChar[] = "ACGTI$-MRSVWYHKDBN"; // nuc_cruc_output.cpp:11

// Text form of an alignment, three lines, including the unaligned 5'/3' overhangs with ':' for
// complementary-but-unaligned positions (nuc_cruc_output.cpp:87-204).
std::string render_alignment_impl(const BoundRec &r, const OligoStrand &os)
{
	const int Lq = os.len, Lt = r.Lt;
	auto q = [&](int i) -> int { return (i >= 0 && i < Lq) ? os.seq[i] : (int)bGAP; };
	auto t = [&](int i) -> int { return (i >= 0 && i < Lt) ? r.win[i] : (int)bGAP; };

	const int prefix = std::max(0, std::min<int>(r.fm_q, Lt - 1 - r.fm_t));
	const int suffix = std::max(0, std::min<int>(Lq - 1 - r.lm_q, r.lm_t));

	std::string s;
	s.reserve(3*(r.ncols + prefix + suffix + 8));
	s += "5' ";
	for (int i = 0; i < prefix; ++i) s += kBaseChar[q(r.fm_q - prefix + i)];
	for (int c = 0; c < r.ncols; ++c) s += kBaseChar[r.cols_q[c]];
	for (int i = 0; i < suffix; ++i) s += kBaseChar[q(r.lm_q + 1 + i)];
	s += " 3'\n   ";
	for (int i = 0; i < prefix; ++i) s += complementary(q(r.fm_q - prefix + i), t(r.fm_t + prefix - i)) ? ':' : ' ';
	for (int c = 0; c < r.ncols; ++c) s += complementary(r.cols_t[c], r.cols_q[c]) ? '|' : ' ';
	for (int i = 0; i < suffix; ++i) s += complementary(q(r.lm_q + 1 + i), t(r.lm_t - i - 1)) ? ':' : ' ';
	s += "\n3' ";
	for (int i = prefix; i > 0; --i) s += kBaseChar[t(r.fm_t + i)];
	for (int c = 0; c < r.ncols; ++c) s += kBaseChar[r.cols_t[c]];
	for (int i = 1; i <= suffix; ++i) s += kBaseChar[t(r.lm_t - i)];
	s += " 5'";
	return s;
}

tnt_bound_oligo empty_slot()
{
	// hybrid_sig::init() (hybrid_sig.h:52-107)
	tnt_bound_oligo b{};
	b.oligo = TNT_OLIGO_NONE;
	b.tm = -1.0f;
	b.dH = 100.0f;
	b.dS = 0.0f;
	b.num_mm = b.num_gap = -1;
	return b;
}

tnt_bound_oligo slot_of(const BoundSite &s, int oligo)
{
	tnt_bound_oligo b{};
	b.oligo = oligo;
	b.loc_5 = s.loc5;
	b.loc_3 = s.loc3;
	b.tm = s.tm;
	b.dH = s.dH;
	b.dS = s.dS;
	b.num_mm = (int8_t)s.num_mm;   // hybrid_sig stores int8_t (hybrid_sig.h:165-176)
	b.num_gap = (int8_t)s.num_gap;
	b.anchor_5 = s.anchor5;
	b.anchor_3 = s.anchor3;
	b.align_off = 0;
	return b;
}

tnt_hit blank_hit(int assay_index, int assay_id, uint32_t target)
{
	tnt_hit h{};
	h.assay_index = assay_index;
	h.assay_id = assay_id;
	h.target_id = target;
	h.primer_strand = TNT_PLUS;
	h.probe_strand = TNT_PLUS;
	h.forward = h.reverse = h.probe = empty_slot();
	h.forward_clamp = h.reverse_clamp = -1;
	return h;
}

// Duplicate windows can report the same target range; one site survives per (loc_5, loc_3).
// mask-variant order (bind_oligo.cpp:49-82): highest Tm, then most mismatches, then the longest
// alignment text; hash-variant (oligo_info::operator<, tntblast.h:230-242): highest Tm.
void unique_sites(std::vector<const BoundSite *> &v, bool mask_variant)
{
	if (mask_variant) {
		// the reference collects with push_front over a list ordered by seed position
		std::stable_sort(v.begin(), v.end(), [](const BoundSite *a, const BoundSite *b) { return a->target_loc > b->target_loc; });
		std::stable_sort(v.begin(), v.end(), [](const BoundSite *a, const BoundSite *b) {
			if (a->loc5 != b->loc5) return a->loc5 < b->loc5;
			if (a->loc3 != b->loc3) return a->loc3 < b->loc3;
			if (a->tm == b->tm) {
				if (a->num_mm == b->num_mm) return a->align_len > b->align_len;
				return a->num_mm > b->num_mm;
			}
			return a->tm > b->tm;
		});
	}
	else {
		// seeds are visited in diagonal order (bind_oligo.cpp:157-161)
		std::stable_sort(v.begin(), v.end(), [](const BoundSite *a, const BoundSite *b) {
			return ((int)a->query_loc - (int)a->target_loc) < ((int)b->query_loc - (int)b->target_loc);
		});
		std::stable_sort(v.begin(), v.end(), [](const BoundSite *a, const BoundSite *b) {
			if (a->loc5 != b->loc5) return a->loc5 < b->loc5;
			if (a->loc3 != b->loc3) return a->loc3 < b->loc3;
			return a->tm > b->tm;
		});
	}
	size_t m = 0;
	for (size_t i = 0; i < v.size(); ++i)
		if (m == 0 || v[m - 1]->loc5 != v[i]->loc5 || v[m - 1]->loc3 != v[i]->loc3) v[m++] = v[i];
	v.resize(m);
}

void join_pcr_sorted(const std::vector<const BoundSite *> &all, int assay_index, int assay_id, bool has_probe,
	const AssembleOptions &opt, const BoundSite *base, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs);

void join_pcr(const std::vector<const BoundSite *> &group, int assay_index, int assay_id, bool has_probe,
	const AssembleOptions &opt, const BoundSite *base, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	// Cheap exits before any sorting: an amplicon needs a minus-strand primer site f and a
	// plus-strand primer site r with f.loc_3 < r.loc_5 and r.loc_3 - f.loc_5 + 1 <= max_len
	// (amplicon_search.cpp:383-390), and a TaqMan assay a probe site as well (:399-406).
	{
		bool any_probe = false, any_pair = false;
		for (const BoundSite *s : group) any_probe |= (s->role == TNT_OLIGO_P);
		if (has_probe && !any_probe) return;
		for (const BoundSite *f : group) {
			if (f->plus || f->role == TNT_OLIGO_P) continue;
			for (const BoundSite *r : group) {
				if (!r->plus || r->role == TNT_OLIGO_P) continue;
				if (f->loc3 < r->loc5 && (r->loc3 - f->loc5 + 1) <= (int)opt.max_len) { any_pair = true; break; }
			}
			if (any_pair) break;
		}
		if (!any_pair) return;
	}
	// per (role, strand) uniqueness, then one list ordered by (loc_5, loc_3)
	std::vector<const BoundSite *> all;
	for (int plus = 0; plus < 2; ++plus)
		for (int role = 0; role < 3; ++role) {
			std::vector<const BoundSite *> cat;
			for (const BoundSite *s : group) if (s->plus == plus && s->role == role) cat.push_back(s);
			unique_sites(cat, true);
			all.insert(all.end(), cat.begin(), cat.end());
		}
	std::stable_sort(all.begin(), all.end(), [](const BoundSite *a, const BoundSite *b) {
		if (a->loc5 == b->loc5) return a->loc3 < b->loc3; // sort_by_oligo_loc, amplicon_search.cpp:12-26
		return a->loc5 < b->loc5;
	});

	join_pcr_sorted(all, assay_index, assay_id, has_probe, opt, base, hits, refs);
}

// The final F x R (x P) loops of amplicon() over the list in its final order (amplicon_search.cpp:355-674)
void join_pcr_sorted(const std::vector<const BoundSite *> &all, int assay_index, int assay_id, bool has_probe,
	const AssembleOptions &opt, const BoundSite *base, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	const bool apply_mmc = opt.min_max_primer_clamp >= 0;
	const unsigned mmc = apply_mmc ? (unsigned)opt.min_max_primer_clamp : 0u;

	for (size_t fi = 0; fi < all.size(); ++fi) {
		const BoundSite &f = *all[fi];
		if (f.plus || f.role == TNT_OLIGO_P) continue;
		for (size_t ri = fi + 1; ri < all.size(); ++ri) {
			const BoundSite &r = *all[ri];
			if (!r.plus || r.role == TNT_OLIGO_P) continue;
			if (!opt.single_primer_pcr && f.role == r.role) continue;
			if (f.loc3 >= r.loc5) continue;
			if ((r.loc3 - f.loc5 + 1) > (int)opt.max_len) continue;
			if (apply_mmc && (unsigned)std::max(f.anchor3, r.anchor3) <= mmc) continue;

			auto emit = [&](const BoundSite *p) {
				tnt_hit h;
				HitSites hs;
				make_pcr_hit(f, r, p, assay_index, assay_id, base, h, hs);
				hits.push_back(h);
				refs.push_back(hs);
			};

			if (!has_probe) { emit(nullptr); continue; }
			for (size_t pi = fi + 1; pi < ri; ++pi) {
				const BoundSite &p = *all[pi];
				if (p.role != TNT_OLIGO_P) continue;
				if (!(p.loc5 >= f.loc5 && p.loc3 <= r.loc3)) continue;
				if (p.plus == f.plus) { if (p.loc5 <= f.loc3) continue; } // same strand as the forward primer
				else if (p.loc3 >= r.loc5) continue;                      // same strand as the reverse primer
				emit(&p);
			}
		}
	}
}

} // namespace

void make_pcr_hit(const BoundSite &f, const BoundSite &r, const BoundSite *p, int assay_index, int assay_id,
	const BoundSite *base, tnt_hit &h, HitSites &hs)
{
	// forward primer first in the record whatever bound upstream (amplicon_search.cpp:478-483)
	const bool swap_out = (f.role == TNT_OLIGO_R && r.role == TNT_OLIGO_F);
	const BoundSite &fo = swap_out ? r : f, &ro = swap_out ? f : r;
	const int f_oligo = (f.role == TNT_OLIGO_R && r.role == TNT_OLIGO_R) ? TNT_OLIGO_R : TNT_OLIGO_F;
	const int r_oligo = (f.role == TNT_OLIGO_F && r.role == TNT_OLIGO_F) ? TNT_OLIGO_F : TNT_OLIGO_R;
	h = blank_hit(assay_index, assay_id, f.target);
	h.primer_strand = (f.role == TNT_OLIGO_F) ? TNT_PLUS : TNT_MINUS;
	h.amp_first = f.loc5;
	h.amp_last = r.loc3;
	h.forward = slot_of(fo, f_oligo);
	h.reverse = slot_of(ro, r_oligo);
	hs = HitSites{(int)(&fo - base), (int)(&ro - base), -1};
	h.forward_clamp = (int8_t)fo.anchor3;
	h.reverse_clamp = (int8_t)ro.anchor3;
	if (p) {
		h.probe = slot_of(*p, TNT_OLIGO_P);
		h.probe_first = p->loc5;
		h.probe_last = p->loc3;
		h.probe_strand = p->plus ? TNT_PLUS : TNT_MINUS;
		hs.probe = (int)(p - base);
	}
}

namespace {

void join_padlock(const std::vector<const BoundSite *> &group, int assay_index, int assay_id,
	const AssembleOptions &opt, const BoundSite *base, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	const int max_len = opt.assay_format == TNT_ASSAY_MIPS ? (int)opt.max_len : 0;
	for (int plus = 0; plus < 2; ++plus) {
		std::vector<const BoundSite *> up, down;
		for (const BoundSite *s : group) {
			if (s->plus != plus) continue;
			(s->role == TNT_OLIGO_R ? up : down).push_back(s);
		}
		unique_sites(up, false);
		unique_sites(down, false);
		for (const BoundSite *u : up)
			for (const BoundSite *d : down) {
				const int gap = plus ? d->loc5 - u->loc3 - 1 : u->loc5 - d->loc3 - 1;
				if (gap < 0 || gap > max_len) continue;
				tnt_hit h = blank_hit(assay_index, assay_id, u->target);
				h.primer_strand = plus ? TNT_PLUS : TNT_MINUS;
				h.amp_first = plus ? u->loc5 : d->loc5;
				h.amp_last = plus ? d->loc3 : u->loc3;
				if (h.amp_first > h.amp_last) throw std::runtime_error(":padlock: start > stop");
				h.forward = slot_of(*d, TNT_OLIGO_F);
				h.reverse = slot_of(*u, TNT_OLIGO_R);
				h.forward_clamp = (int8_t)d->anchor3;
				h.reverse_clamp = (int8_t)u->anchor5;
				hits.push_back(h);
				refs.push_back(HitSites{(int)(d - base), (int)(u - base), -1});
			}
	}
}

void join_probe(const std::vector<const BoundSite *> &group, int assay_index, int assay_id,
	const BoundSite *base, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	for (int plus = 0; plus < 2; ++plus) { // minus strand first (probe_search.cpp:92-151)
		std::vector<const BoundSite *> b;
		for (const BoundSite *s : group) if (s->plus == plus) b.push_back(s);
		unique_sites(b, false);
		for (const BoundSite *s : b) {
			if (s->loc5 > s->loc3) throw std::runtime_error(":hybrid: probe_start > probe_stop");
			tnt_hit h = blank_hit(assay_index, assay_id, s->target);
			h.probe = slot_of(*s, TNT_OLIGO_P);
			h.probe_first = s->loc5;
			h.probe_last = s->loc3;
			h.probe_strand = plus ? TNT_PLUS : TNT_MINUS;
			hits.push_back(h);
			refs.push_back(HitSites{-1, -1, (int)(s - base)});
		}
	}
}

} // namespace

std::string render_alignment(const BoundRec &r, const OligoStrand &os)
{
	return r.valid ? render_alignment_impl(r, os) : std::string();
}

BoundSite make_site(const BoundHead &h, uint32_t index, const OligoStrand &os)
{
	BoundSite s;
	s.assay = os.assay;
	s.role = os.role;
	s.plus = os.plus;
	s.os_index = h.os;
	s.index = index;
	s.target = h.target;
	s.loc5 = h.loc5;
	s.loc3 = h.loc3;
	s.tm = h.tm; s.dH = h.dH; s.dS = h.dS;
	s.anchor5 = h.anchor5; s.anchor3 = h.anchor3;
	s.num_mm = h.num_mm; s.num_gap = h.num_gap;
	s.flags = h.flags;
	s.query_loc = h.k;
	s.target_loc = h.t;
	s.align_len = h.align_len;
	return s;
}

void assemble_hits(const std::vector<BoundSite> &sites, const AssembleOptions &opt, const std::vector<int> &assay_ids,
	const std::vector<int> &assay_has_primers, const std::vector<int> &assay_has_probe,
	std::vector<tnt_hit> &hits, std::vector<HitSites> &refs, bool primers_joined_on_device)
{
	// (fragment, assay) pairs in ascending order, like the reference's nested loops; ties keep the
	// input order (one packed 64-bit key per site: group rank, then index)
	std::vector<uint32_t> order(sites.size());
	{
		int max_assay = 0;
		for (const BoundSite &s : sites) max_assay = std::max(max_assay, s.assay);
		const uint64_t na = (uint64_t)max_assay + 1;
		bool packed = sites.size() < ((uint64_t)1 << 24);
		for (const BoundSite &s : sites) packed = packed && ((uint64_t)s.target*na + (uint64_t)s.assay) < ((uint64_t)1 << 40);
		if (packed) {
			std::vector<uint64_t> keys(sites.size());
			for (size_t i = 0; i < sites.size(); ++i)
				keys[i] = (((uint64_t)sites[i].target*na + (uint64_t)sites[i].assay) << 24) | (uint64_t)i;
			std::sort(keys.begin(), keys.end());
			for (size_t i = 0; i < keys.size(); ++i) order[i] = (uint32_t)(keys[i] & 0xffffffu);
		}
		else {
			for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
			std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
				if (sites[a].target != sites[b].target) return sites[a].target < sites[b].target;
				return sites[a].assay < sites[b].assay;
			});
		}
	}
	const BoundSite *base = sites.data();
	// group boundaries in `order`
	std::vector<size_t> starts;
	for (size_t i = 0; i < order.size();) {
		starts.push_back(i);
		size_t j = i;
		while (j < order.size() && sites[order[j]].target == sites[order[i]].target && sites[order[j]].assay == sites[order[i]].assay) ++j;
		i = j;
	}
	starts.push_back(order.size());
	const size_t ngroups = starts.size() - 1;

	auto join_range = [&](size_t g0, size_t g1, std::vector<tnt_hit> &out_hits, std::vector<HitSites> &out_refs) {
		std::vector<const BoundSite *> group;
		for (size_t g = g0; g < g1; ++g) {
			group.clear();
			for (size_t k = starts[g]; k < starts[g + 1]; ++k) group.push_back(&sites[order[k]]);
			const int ai = group[0]->assay;
			const int id = assay_ids[(size_t)ai];
			if (assay_has_primers[(size_t)ai]) { // tntblast_local.cpp:559-611
				if (opt.assay_format == TNT_ASSAY_PADLOCK || opt.assay_format == TNT_ASSAY_MIPS)
					join_padlock(group, ai, id, opt, base, out_hits, out_refs);
				else if (!primers_joined_on_device) join_pcr(group, ai, id, assay_has_probe[(size_t)ai] != 0, opt, base, out_hits, out_refs);
			}
			else join_probe(group, ai, id, base, out_hits, out_refs); // :612-625
		}
	};

	// The groups are independent; many sites -> a few host threads, results concatenated in group order
	const unsigned nthreads = order.size() < 20000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
	if (nthreads == 1) { join_range(0, ngroups, hits, refs); return; }
	std::vector<std::vector<tnt_hit>> part_hits(nthreads);
	std::vector<std::vector<HitSites>> part_refs(nthreads);
	std::vector<std::string> errors(nthreads);
	std::vector<std::thread> pool;
	// equal numbers of sites per thread
	std::vector<size_t> cut(nthreads + 1, ngroups);
	cut[0] = 0;
	for (unsigned t = 1; t < nthreads; ++t) {
		const size_t want = order.size()*t/nthreads;
		cut[t] = (size_t)(std::lower_bound(starts.begin(), starts.end() - 1, want) - starts.begin());
	}
	for (unsigned t = 0; t < nthreads; ++t)
		pool.emplace_back([&, t]() {
			try { join_range(cut[t], cut[t + 1], part_hits[t], part_refs[t]); }
			catch (const std::exception &ex) { errors[t] = ex.what(); }
		});
	for (std::thread &t : pool) t.join();
	for (const std::string &err : errors) if (!err.empty()) throw std::runtime_error(err);
	for (unsigned t = 0; t < nthreads; ++t) {
		hits.insert(hits.end(), part_hits[t].begin(), part_hits[t].end());
		refs.insert(refs.end(), part_refs[t].begin(), part_refs[t].end());
	}
}

// ------------------------------------------------------------------------------------------
// Replay of amplicon()'s staged bind / cull sequence (amplicon_search.cpp:92-355)
// ------------------------------------------------------------------------------------------
namespace {

enum : uint8_t { M_F = 1, M_R = 2, M_P = 4, M_PLUS = 8, M_MINUS = 16, M_VALID = 32 }; // oligo_info, tntblast.h:147-154

struct El {                       // oligo_info (tntblast.h:145-243) without the alignment text
	int loc5 = 0, loc3 = 0;
	float tm = -1.0f;
	int num_mm = 0;
	uint32_t align_len = 0;
	uint32_t q = 0, t = 0;
	uint8_t mask = 0;
	int site = -1;
};

bool less_oligo_loc(const El &a, const El &b) // sort_by_oligo_loc, amplicon_search.cpp:12-26
{
	if (!(a.loc5 + a.loc3) || !(b.loc5 + b.loc3)) return a.t < b.t;
	if (a.loc5 == b.loc5) return a.loc3 < b.loc3;
	return a.loc5 < b.loc5;
}

bool less_bound_match(const El &a, const El &b) // sort_by_bound_match, bind_oligo.cpp:49-82
{
	if (a.loc5 != b.loc5) return a.loc5 < b.loc5;
	if (a.loc3 != b.loc3) return a.loc3 < b.loc3;
	if (a.tm == b.tm) {
		if (a.num_mm == b.num_mm) return a.align_len > b.align_len;
		return a.num_mm > b.num_mm;
	}
	return a.tm > b.tm;
}

// std::list::sort of libstdc++ (bits/list.tcc: bottom-up merge, 64 bins, merges in place with
// "take from the second run only if strictly less") on an array.  sort_by_oligo_loc is no strict
// weak ordering once bound and unbound elements mix, so the outcome depends on this exact merge
// sequence; the reference was built with this very library.
typedef bool (*ElLess)(const El &, const El &);

void merge_runs(std::vector<El> &a, std::vector<El> &b, ElLess less) // a.merge(b): result in a, b emptied
{
	std::vector<El> r;
	r.reserve(a.size() + b.size());
	size_t i = 0, j = 0;
	while (i < a.size() && j < b.size()) {
		if (less(b[j], a[i])) r.push_back(b[j++]);
		else r.push_back(a[i++]);
	}
	for (; i < a.size(); ++i) r.push_back(a[i]);
	for (; j < b.size(); ++j) r.push_back(b[j]);
	a.swap(r);
	b.clear();
}

void list_sort(std::vector<El> &l, ElLess less)
{
	if (l.size() < 2) return;
	std::vector<El> carry, bins[64];
	int fill = 0;
	for (const El &e : l) {
		carry.assign(1, e);
		int counter;
		for (counter = 0; counter != fill && !bins[counter].empty(); ++counter) {
			merge_runs(bins[counter], carry, less);
			carry.swap(bins[counter]);
		}
		carry.swap(bins[counter]);
		if (counter == fill) ++fill;
	}
	for (int counter = 1; counter != fill; ++counter) merge_runs(bins[counter], bins[counter - 1], less);
	l.swap(bins[fill - 1]);
}

bool is_bound(const El &e) { return (e.loc5 + e.loc3) != 0; }

// The list the culls sort, as an array.  `tail` = first element of the run appended by the last bind
// step (sorted by loc_5 / loc_3); everything in front of it is still in the order of the last sort.
struct MatchList {
	std::vector<El> v;
	size_t tail = 0;
	bool canonical = false;   // the head run is sorted under a comparator that was a strict weak ordering

	// Is sort_by_oligo_loc a strict weak ordering on the present elements?  It compares bound pairs by
	// (loc_5, loc_3) and everything else by seed position: it is one iff the bound elements are ordered
	// strictly alike by both keys.
	bool ordering_is_consistent() const
	{
		std::vector<const El *> b;
		for (const El &e : v) if (is_bound(e)) b.push_back(&e);
		for (size_t i = 0; i < b.size(); ++i)
			for (size_t j = i + 1; j < b.size(); ++j) {
				const bool loc_ij = less_oligo_loc(*b[i], *b[j]), loc_ji = less_oligo_loc(*b[j], *b[i]);
				const bool t_ij = b[i]->t < b[j]->t, t_ji = b[j]->t < b[i]->t;
				if (loc_ij != t_ij || loc_ji != t_ji) return false;
			}
		return true;
	}

	void sort()
	{
		if (ordering_is_consistent()) {
			// a stable sort has one possible outcome: take the cheap route to it
			if (canonical && tail <= v.size()) std::inplace_merge(v.begin(), v.begin() + (ptrdiff_t)tail, v.end(), less_oligo_loc);
			else std::stable_sort(v.begin(), v.end(), less_oligo_loc);
			canonical = true;
		}
		else {
			list_sort(v, less_oligo_loc);
			canonical = false;
		}
		tail = v.size();
	}
};

// cull_oligo_match (amplicon_search.cpp:679-765), unsigned wrap of the seed distance included.  The
// strand counts are taken from the element *after* each kept one, as the reference does (:748-753;
// its read of end() counts as neither strand).
void cull(MatchList &ml, unsigned max_amplicon_len, bool has_probe, bool single_primer_pcr, unsigned *n_minus, unsigned *n_plus)
{
	const unsigned threshold = max_amplicon_len + 50;
	ml.sort();
	std::vector<El> &l = ml.v;
	for (El &e : l) e.mask &= (uint8_t)~M_VALID;
	const size_t n = l.size();
	for (size_t f = 0; f < n; ++f) {
		if (l[f].mask & (M_PLUS | M_P)) continue;
		for (size_t r = f + 1; r < n; ++r) {
			if ((unsigned)(l[r].t - l[f].t) > threshold) break;
			if (l[r].mask & (M_MINUS | M_P)) continue;
			if (!single_primer_pcr && ((l[f].mask & (M_R | M_F)) == (l[r].mask & (M_R | M_F)))) continue;
			if (has_probe) {
				for (size_t p = f + 1; p < r; ++p)
					if (l[p].mask & M_P) { l[p].mask |= M_VALID; l[f].mask |= M_VALID; l[r].mask |= M_VALID; }
			}
			else { l[f].mask |= M_VALID; l[r].mask |= M_VALID; }
		}
	}
	unsigned cm = 0, cp = 0;
	size_t m = 0;
	for (size_t i = 0; i < n; ++i) {
		if (!(l[i].mask & M_VALID)) continue;
		const uint8_t next = i + 1 < n ? l[i + 1].mask : (uint8_t)0;
		cm += (next & M_MINUS) ? 1u : 0u;
		cp += (next & M_PLUS) ? 1u : 0u;
		if (m != i) l[m] = l[i];
		++m;
	}
	l.resize(m);
	ml.tail = m;
	if (n_minus) *n_minus = cm;
	if (n_plus) *n_plus = cp;
}

// bind_oligo_to_{minus,plus}_strand, mask variant (bind_oligo.cpp:456-827, :1159-1530): the elements
// of one (oligo, strand) leave the list; those whose window passes every filter come back as bound
// sites, one per (loc_5, loc_3), behind everything else.
void bind_masked(MatchList &ml, uint8_t want, const std::vector<BoundSite> &sites)
{
	std::vector<El> &l = ml.v;
	std::vector<El> cur;
	size_t m = 0, head = 0; // head: survivors of the sorted run in front of `tail`
	const bool pending_run = ml.tail < l.size(); // a run appended by an earlier bind step is not sorted in yet
	for (size_t i = 0; i < l.size(); ++i) {
		if ((l[i].mask & want) != want) { if (m != i) l[m] = l[i]; ++m; if (i < ml.tail) ++head; continue; }
		El e = l[i];
		if (e.site < 0) continue;
		const BoundSite &b = sites[(size_t)e.site];
		e.loc5 = b.loc5; e.loc3 = b.loc3;
		e.tm = b.tm;
		e.num_mm = b.num_mm;
		e.align_len = b.align_len;
		cur.push_back(e);
	}
	// two bind steps in a row (the probe strands) leave two runs behind the sorted head: the next sort
	// then starts from scratch
	if (pending_run) ml.canonical = false;
	l.resize(m);
	ml.tail = ml.canonical ? head : m;
	if (cur.empty()) return;
	std::reverse(cur.begin(), cur.end());            // curr_oligo was filled with push_front
	std::stable_sort(cur.begin(), cur.end(), less_bound_match); // a strict weak ordering: any stable sort
	for (size_t k = 0; k < cur.size(); ++k)
		if (k == 0 || l.back().loc5 != cur[k].loc5 || l.back().loc3 != cur[k].loc3) l.push_back(cur[k]);
}

} // namespace

void replay_pcr_group(std::vector<ReplaySeed> seeds, const std::vector<BoundSite> &sites, const AssembleOptions &opt,
	bool has_probe, int assay_index, int assay_id, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	// match_oligo_to_*_strand (bind_oligo.cpp:84-122): per (oligo, strand) one seed per diagonal,
	// ordered by q - t; the merge of all-unbound lists is an append
	std::stable_sort(seeds.begin(), seeds.end(), [](const ReplaySeed &a, const ReplaySeed &b) {
		if (a.cat != b.cat) return a.cat < b.cat;
		return ((int)a.q - (int)a.t) < ((int)b.q - (int)b.t);
	});
	static const uint8_t kMask[6] = {M_F | M_MINUS, M_R | M_MINUS, M_F | M_PLUS, M_R | M_PLUS, M_P | M_MINUS, M_P | M_PLUS};
	MatchList ml;
	ml.v.reserve(seeds.size());
	size_t k = 0;
	auto append = [&](int cat) {
		for (; k < seeds.size() && seeds[k].cat == cat; ++k) {
			El e;
			e.q = seeds[k].q; e.t = seeds[k].t;
			e.mask = kMask[cat];
			e.site = seeds[k].site;
			ml.v.push_back(e);
		}
	};
	append(0); append(1);
	const size_t n_minus = ml.v.size();
	if (n_minus == 0) return;                       // amplicon_search.cpp:103-105
	append(2); append(3);
	const size_t n_plus = ml.v.size();
	if (n_plus == n_minus) return;                  // :113-115
	if (has_probe) {
		append(4); append(5);
		if (ml.v.size() == n_plus) return;          // :123-125
	}

	unsigned cm = 0, cp = 0;
	cull(ml, opt.max_len, has_probe, opt.single_primer_pcr, &cm, &cp);
	const bool first_plus = !(cm < cp);             // :131 vs :218
	for (int stage = 0; stage < 4; ++stage) {
		const bool plus = (stage < 2) ? first_plus : !first_plus;
		const bool is_r = stage & 1;
		bind_masked(ml, (uint8_t)((is_r ? M_R : M_F) | (plus ? M_PLUS : M_MINUS)), sites);
		if (stage < 3) {
			cull(ml, opt.max_len, has_probe, opt.single_primer_pcr, nullptr, nullptr);
			// early exits at :153, :177, :240, :264, :288 -- none after the third bind of the minus-first path (:199)
			if (ml.v.empty() && !(stage == 2 && !first_plus)) return;
		}
	}
	if (has_probe) {
		cull(ml, opt.max_len, has_probe, opt.single_primer_pcr, nullptr, nullptr);
		if (ml.v.empty()) return;
		bind_masked(ml, (uint8_t)(M_P | M_MINUS), sites);
		bind_masked(ml, (uint8_t)(M_P | M_PLUS), sites);
	}
	ml.sort();                                      // :353
	std::vector<const BoundSite *> all;
	all.reserve(ml.v.size());
	for (const El &e : ml.v) if (e.site >= 0) all.push_back(&sites[(size_t)e.site]);
	join_pcr_sorted(all, assay_index, assay_id, has_probe, opt, sites.data(), hits, refs);
}

// ------------------------------------------------------------------------------------------
// The same replay written with std::list, operation by operation as the reference has it
// (list::sort, erase while iterating, push_front / sort / push_back in the bind step).  Slower;
// kept as the yardstick the array version above is checked against (tnt_debug_replay_selftest).
// ------------------------------------------------------------------------------------------
namespace {

// cull_oligo_match (amplicon_search.cpp:679-765), unsigned wrap of the seed distance included.  The
// strand counts are taken from the element *after* each kept one, as the reference does (:748-753;
// its read of end() counts as neither strand).
void cull_list(std::list<El> &l, unsigned max_amplicon_len, bool has_probe, bool single_primer_pcr, unsigned *n_minus, unsigned *n_plus)
{
	const unsigned threshold = max_amplicon_len + 50;
	l.sort(less_oligo_loc);
	for (El &e : l) e.mask &= (uint8_t)~M_VALID;
	for (auto f = l.begin(); f != l.end(); ++f) {
		if (f->mask & (M_PLUS | M_P)) continue;
		auto r = f;
		for (++r; r != l.end(); ++r) {
			if ((unsigned)(r->t - f->t) > threshold) break;
			if (r->mask & (M_MINUS | M_P)) continue;
			if (!single_primer_pcr && ((f->mask & (M_R | M_F)) == (r->mask & (M_R | M_F)))) continue;
			if (has_probe) {
				auto p = f;
				for (++p; p != r; ++p)
					if (p->mask & M_P) { p->mask |= M_VALID; f->mask |= M_VALID; r->mask |= M_VALID; }
			}
			else { f->mask |= M_VALID; r->mask |= M_VALID; }
		}
	}
	unsigned cm = 0, cp = 0;
	for (auto i = l.begin(); i != l.end();) {
		if (i->mask & M_VALID) {
			++i;
			const uint8_t next = i != l.end() ? i->mask : (uint8_t)0;
			cm += (next & M_MINUS) ? 1u : 0u;
			cp += (next & M_PLUS) ? 1u : 0u;
			continue;
		}
		i = l.erase(i);
	}
	if (n_minus) *n_minus = cm;
	if (n_plus) *n_plus = cp;
}

// bind_oligo_to_{minus,plus}_strand, mask variant (bind_oligo.cpp:456-827, :1159-1530): the elements
// of one (oligo, strand) leave the list; those whose window passes every filter come back as bound
// sites, one per (loc_5, loc_3), behind everything else.
void bind_masked_list(std::list<El> &l, uint8_t want, const std::vector<BoundSite> &sites)
{
	std::list<El> cur;
	for (auto it = l.begin(); it != l.end();) {
		if ((it->mask & want) != want) { ++it; continue; }
		El e = *it;
		it = l.erase(it);
		if (e.site < 0) continue;
		const BoundSite &b = sites[(size_t)e.site];
		e.loc5 = b.loc5; e.loc3 = b.loc3;
		e.tm = b.tm;
		e.num_mm = b.num_mm;
		e.align_len = b.align_len;
		cur.push_front(e);
	}
	if (cur.empty()) return;
	cur.sort(less_bound_match);
	l.push_back(cur.front());
	cur.pop_front();
	while (!cur.empty()) {
		if (l.back().loc5 != cur.front().loc5 || l.back().loc3 != cur.front().loc3) l.push_back(cur.front());
		cur.pop_front();
	}
}

} // namespace

void replay_pcr_group_list(std::vector<ReplaySeed> seeds, const std::vector<BoundSite> &sites, const AssembleOptions &opt,
	bool has_probe, int assay_index, int assay_id, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs)
{
	// match_oligo_to_*_strand (bind_oligo.cpp:84-122): per (oligo, strand) one seed per diagonal,
	// ordered by q - t; the merge of all-unbound lists is an append
	std::stable_sort(seeds.begin(), seeds.end(), [](const ReplaySeed &a, const ReplaySeed &b) {
		if (a.cat != b.cat) return a.cat < b.cat;
		return ((int)a.q - (int)a.t) < ((int)b.q - (int)b.t);
	});
	static const uint8_t kMask[6] = {M_F | M_MINUS, M_R | M_MINUS, M_F | M_PLUS, M_R | M_PLUS, M_P | M_MINUS, M_P | M_PLUS};
	std::list<El> ml;
	size_t k = 0;
	auto append = [&](int cat) {
		for (; k < seeds.size() && seeds[k].cat == cat; ++k) {
			El e;
			e.q = seeds[k].q; e.t = seeds[k].t;
			e.mask = kMask[cat];
			e.site = seeds[k].site;
			ml.push_back(e);
		}
	};
	append(0); append(1);
	const size_t n_minus = ml.size();
	if (n_minus == 0) return;                       // amplicon_search.cpp:103-105
	append(2); append(3);
	const size_t n_plus = ml.size();
	if (n_plus == n_minus) return;                  // :113-115
	if (has_probe) {
		append(4); append(5);
		if (ml.size() == n_plus) return;            // :123-125
	}

	unsigned cm = 0, cp = 0;
	cull_list(ml, opt.max_len, has_probe, opt.single_primer_pcr, &cm, &cp);
	const bool first_plus = !(cm < cp);             // :131 vs :218
	for (int stage = 0; stage < 4; ++stage) {
		const bool plus = (stage < 2) ? first_plus : !first_plus;
		const bool is_r = stage & 1;
		bind_masked_list(ml, (uint8_t)((is_r ? M_R : M_F) | (plus ? M_PLUS : M_MINUS)), sites);
		if (stage < 3) {
			cull_list(ml, opt.max_len, has_probe, opt.single_primer_pcr, nullptr, nullptr);
			// early exits at :153, :177, :240, :264, :288 -- none after the third bind of the minus-first path (:199)
			if (ml.empty() && !(stage == 2 && !first_plus)) return;
		}
	}
	if (has_probe) {
		cull_list(ml, opt.max_len, has_probe, opt.single_primer_pcr, nullptr, nullptr);
		if (ml.empty()) return;
		bind_masked_list(ml, (uint8_t)(M_P | M_MINUS), sites);
		bind_masked_list(ml, (uint8_t)(M_P | M_PLUS), sites);
	}
	ml.sort(less_oligo_loc);                        // :353
	std::vector<const BoundSite *> all;
	all.reserve(ml.size());
	for (const El &e : ml) if (e.site >= 0) all.push_back(&sites[(size_t)e.site]);
	join_pcr_sorted(all, assay_index, assay_id, has_probe, opt, sites.data(), hits, refs);
}

void hit_sequence_plan(const tnt_hit &h, int assay_format, int &start, int &stop, SeqMode &mode)
{
	const bool primers = h.forward.oligo != TNT_OLIGO_NONE;
	if (!primers) {
		start = h.probe_first;
		stop = h.probe_last;
		mode = h.probe_strand == TNT_PLUS ? SeqMode::ProbePlus : SeqMode::ProbeMinus;
		return;
	}
	start = h.amp_first;
	stop = h.amp_last;
	if (assay_format == TNT_ASSAY_PADLOCK || assay_format == TNT_ASSAY_MIPS)
		mode = h.primer_strand == TNT_MINUS ? SeqMode::PadlockMinusStrand : SeqMode::PadlockPlusStrand;
	else mode = h.primer_strand == TNT_PLUS ? SeqMode::PcrPlus : SeqMode::PcrMinus;
}

// first_i of render_hit_sequence and the reading direction, per mode
static void hit_sequence_walk(int start, int stop, SeqMode mode, int seq_len, bool &forward, int &first_i)
{
	forward = true;
	first_i = 0;
	switch (mode) {
	case SeqMode::PcrPlus: forward = true; first_i = std::max(0, -start); break;                 // amplicon_search.cpp:512-523
	case SeqMode::PcrMinus: forward = false; first_i = std::max(0, stop - seq_len + 1); break;  // :526-537
	case SeqMode::ProbePlus: forward = true; first_i = 0; break;                                // probe_search.cpp:205-219
	case SeqMode::ProbeMinus: forward = false; first_i = 0; break;                              // :129-142
	case SeqMode::PadlockMinusStrand: forward = true; first_i = std::max(0, 1 - start); break;  // padlock_search.cpp:206-218
	case SeqMode::PadlockPlusStrand: forward = false; first_i = std::max(0, stop - seq_len - 1); break; // :341-352
	}
}

// The fragment positions [lo, hi] render_hit_sequence reads for this hit (hi < lo: none).  Not simply the
// overlap of [start, stop] with the fragment: a probe site that overhangs the end of a fragment is printed by
// the reference from the clamped end for the full length (probe_search.cpp:129-142, :205-219), i.e. from
// bases outside [start, stop] (found by tools/fuzz_parity.py: such a hit made the text fetch throw).
void hit_sequence_fetch_range(int start, int stop, SeqMode mode, int seq_len, int &lo, int &hi)
{
	bool forward;
	int first_i;
	hit_sequence_walk(start, stop, mode, seq_len, forward, first_i);
	const long count = (long)(stop - start + 1) - first_i;
	lo = 0;
	hi = -1;
	if (count <= 0 || seq_len <= 0) return;
	if (forward) {
		const long p = std::max(0, start);
		lo = (int)p;
		hi = (int)std::min<long>(p + count - 1, (long)seq_len - 1);
	}
	else {
		const long p = std::min(stop, seq_len - 1);
		hi = (int)p;
		lo = (int)std::max<long>(p - (count - 1), 0);
	}
}

std::string render_hit_sequence(int start, int stop, SeqMode mode, int seq_len, int lo, const std::vector<uint8_t> &codes)
{
	static const char fwd[] = "ACGTIMRSVWYHKDBN-";   // hash_base_to_ascii (seq.h:58-101)
	static const char cmp[] = "TGCAIKYSBWRDMHVN-";   // hash_base_to_ascii_complement (seq.h:103-146)
	const int n = stop - start + 1;
	std::string s((size_t)n, '-');
	auto code_at = [&](long p) -> int {
		const long k = p - lo;
		if (k < 0 || k >= (long)codes.size()) throw std::runtime_error("internal: sequence fetch out of range");
		const int c = codes[(size_t)k];
		if (c > 16) throw std::runtime_error(":hash_base_to_ascii: Illegal base");
		return c;
	};
	bool forward = true;
	int first_i = 0;
	hit_sequence_walk(start, stop, mode, seq_len, forward, first_i);
	if (forward) {
		long p = std::max(0, start);
		for (int i = first_i; i < n; ++i, ++p) {
			if (p >= seq_len) break;
			s[(size_t)i] = fwd[code_at(p)];
		}
	}
	else {
		long p = std::min(stop, seq_len - 1);
		for (int i = first_i; i < n; ++i, --p) {
			if (p < 0) break;
			s[(size_t)i] = cmp[code_at(p)];
		}
	}
	return s;
}

// Random match lists through both replays; returns the number of cases whose hit lists differ.
long replay_selftest(uint32_t seed, int cases, long *hits_out)
{
	std::mt19937 rng(seed);
	long ndiff = 0, nhits = 0;
	for (int it = 0; it < cases; ++it) {
		const bool has_probe = it & 1;
		const int ncat = has_probe ? 6 : 4;
		const int L = 20, span = 300 + (int)(rng() % 3000);
		std::vector<ReplaySeed> seeds;
		std::vector<BoundSite> sites;
		const int nseed = 4 + (int)(rng() % 60);
		for (int s = 0; s < nseed; ++s) {
			ReplaySeed r;
			r.cat = (int)(rng() % ncat);
			r.q = rng() % 14;
			// clusters, so that bound sites overlap and their two orders disagree
			const int base = (rng() % 4) ? (int)(rng() % span) : 100 + (int)(rng() % 40);
			r.t = (uint32_t)(base + 50);
			r.site = -1;
			if (rng() % 3) {
				BoundSite b{};
				b.role = r.cat >= 4 ? 2 : (r.cat & 1);
				b.plus = r.cat >= 4 ? (r.cat & 1) : (r.cat >= 2);
				b.os_index = (uint32_t)r.cat;
				b.index = (uint32_t)sites.size();
				const int shift = (int)(rng() % 5) - 2;
				b.loc5 = (int)r.t - (int)r.q + shift - ((rng() % 7) == 0 ? (int)(rng() % 15) : 0);
				b.loc3 = b.loc5 + L - 1 + (int)(rng() % 3);
				b.tm = 40.0f + (float)(rng() % 200)/10.0f;
				b.dH = -100.0f; b.dS = -0.3f;
				b.anchor5 = 3; b.anchor3 = (int)(rng() % 10);
				b.num_mm = (int)(rng() % 5);
				b.query_loc = r.q; b.target_loc = r.t;
				b.align_len = 60 + 3*(uint32_t)(rng() % 4);
				r.site = (int)sites.size();
				sites.push_back(b);
			}
			seeds.push_back(r);
		}
		AssembleOptions o{0, (uint32_t)(200 + rng() % 1800), (bool)(rng() & 1), (rng() % 3) ? -1 : 2};
		std::vector<tnt_hit> h1, h2;
		std::vector<HitSites> r1, r2;
		replay_pcr_group_list(seeds, sites, o, has_probe, 0, 7, h1, r1);
		replay_pcr_group(seeds, sites, o, has_probe, 0, 7, h2, r2);
		bool same = r1.size() == r2.size();
		for (size_t k = 0; same && k < r1.size(); ++k)
			same = r1[k].forward == r2[k].forward && r1[k].reverse == r2[k].reverse && r1[k].probe == r2[k].probe;
		nhits += (long)r1.size();
		if (!same) ++ndiff;
	}
	if (hits_out) *hits_out = nhits;
	return ndiff;
}

} // namespace tnt
