// Stage C on the host: bound oligos -> amplicon / padlock / probe hits, alignment text.
#pragma once

#include <string>
#include <vector>

#include "../../include/tntb200.h"
#include "tnt_types.h"

namespace tnt {

// One bound oligo site == oligo_info of the reference (tntblast.h:145-243), without the alignment
// text: only its length takes part in the ordering (bind_oligo.cpp:69-74).
struct BoundSite {
	int assay, role, plus;
	uint32_t os_index;   // global oligo-strand index (BoundHead::os)
	uint32_t index;      // record index in the device-side bound-site buffer
	uint32_t target;
	int loc5, loc3;
	float tm, dH, dS;
	int anchor5, anchor3, num_mm, num_gap;
	unsigned flags;
	uint32_t query_loc, target_loc;   // the seed
	uint32_t align_len;
};

BoundSite make_site(const BoundHead &h, uint32_t index, const OligoStrand &os);

// NucCruc operator<< text (nuc_cruc_output.cpp:74-205) of one record
std::string render_alignment(const BoundRec &r, const OligoStrand &os);

// which bound sites a hit is made of (indices into the `sites` array, -1: none)
struct HitSites { int forward, reverse, probe; };

struct AssembleOptions {
	int assay_format;
	uint32_t max_len;
	bool single_primer_pcr;
	int min_max_primer_clamp;
};

// amplicon() join (amplicon_search.cpp:355-674), padlock() joins (padlock_search.cpp:130-358),
// hybrid() (probe_search.cpp:103-227); `sites` holds everything that passed the per-oligo filters.
// The alignment-text offsets of the hits are left at 0; `refs` tells the caller which sites to render.
// `primers_joined_on_device`: the PCR join of assays with primers was done by k_pair_join (the caller turns
// its index triples into hits with make_pcr_hit); only probe-only assays are assembled here.
void assemble_hits(const std::vector<BoundSite> &sites, const AssembleOptions &opt, const std::vector<int> &assay_ids,
	const std::vector<int> &assay_has_primers, const std::vector<int> &assay_has_probe,
	std::vector<tnt_hit> &hits, std::vector<HitSites> &refs, bool primers_joined_on_device = false);

// One PCR hit from its sites: f = the minus-strand primer site, r = the plus-strand primer site, p = probe
// site or nullptr (amplicon_search.cpp:447-555, :565-671)
void make_pcr_hit(const BoundSite &f, const BoundSite &r, const BoundSite *p, int assay_index, int assay_id,
	const BoundSite *base, tnt_hit &hit, HitSites &ref);

// ---- exact replay of the reference's staged PCR search for one (fragment, assay) group ----------
// amplicon() binds the four primer categories one after the other and culls its match list in
// between (amplicon_search.cpp:131-355).  The cull sorts a list of bound sites and not-yet-bound
// seeds with a comparator that is no strict weak ordering once the two kinds mix, and ends its
// partner scan on an unsigned difference of seed positions (:12-26, :709): when two bound sites
// of one assay overlap, a site that is part of a real amplicon can be dropped.  For the groups
// where that can happen the engine aligns *every* seed of the group and hands seeds + bound sites
// to this function, which walks through the reference's sequence of list operations literally
// (std::list and its sort included: the outcome depends on the merge order).
struct ReplaySeed {
	int cat;             // 0 F-minus, 1 R-minus, 2 F-plus, 3 R-plus, 4 P-minus, 5 P-plus
	uint32_t q, t;       // seed: word index (oligo_info::query_loc), target position
	int site;            // index into `sites` when the seed's window passes every filter, else -1
};

void replay_pcr_group(std::vector<ReplaySeed> seeds, const std::vector<BoundSite> &sites, const AssembleOptions &opt,
	bool has_probe, int assay_index, int assay_id, std::vector<tnt_hit> &hits, std::vector<HitSites> &refs);

// the replay above against its literal std::list form on random match lists (CPU self-test)
long replay_selftest(uint32_t seed, int cases, long *hits_out);

enum class SeqMode { PcrPlus, PcrMinus, ProbePlus, ProbeMinus, PadlockMinusStrand, PadlockPlusStrand };

void hit_sequence_plan(const tnt_hit &h, int assay_format, int &start, int &stop, SeqMode &mode);

// `codes` are the seq.h codes of fragment positions [lo, lo + codes.size())
// fragment positions [lo, hi] the renderer reads (hi < lo: none)
void hit_sequence_fetch_range(int start, int stop, SeqMode mode, int seq_len, int &lo, int &hi);
std::string render_hit_sequence(int start, int stop, SeqMode mode, int seq_len, int lo, const std::vector<uint8_t> &codes);

} // namespace tnt
