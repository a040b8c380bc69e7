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
void assemble_hits(const std::vector<BoundSite> &sites, const AssembleOptions &opt, const std::vector<int> &assay_ids,
	const std::vector<int> &assay_has_primers, const std::vector<int> &assay_has_probe,
	std::vector<tnt_hit> &hits, std::vector<HitSites> &refs);

enum class SeqMode { PcrPlus, PcrMinus, ProbePlus, ProbeMinus, PadlockMinusStrand, PadlockPlusStrand };

void hit_sequence_plan(const tnt_hit &h, int assay_format, int &start, int &stop, SeqMode &mode);

// `codes` are the seq.h codes of fragment positions [lo, lo + codes.size())
std::string render_hit_sequence(int start, int stop, SeqMode mode, int seq_len, int lo, const std::vector<uint8_t> &codes);

} // namespace tnt
