// Stage C on the host: bound oligos -> amplicon / padlock / probe hits, alignment text.
#pragma once

#include <string>
#include <vector>

#include "../../include/tntb200.h"
#include "tnt_types.h"

namespace tnt {

// One bound oligo site == oligo_info of the reference (tntblast.h:145-243)
struct BoundSite {
	int assay, role, plus;
	uint32_t target;
	int loc5, loc3;
	float tm, dH, dS, dG;
	int anchor5, anchor3, num_mm, num_gap, poly_degen;
	int valid;
	unsigned flags;
	uint32_t query_loc, target_loc;   // the seed
	int win_start, win_stop;
	int q_first, q_last, t_first, t_last;
	std::string alignment;            // NucCruc operator<< text (nuc_cruc_output.cpp:74-205)
};

BoundSite make_site(const BoundRec &rec, const OligoStrand &os);

struct AssembleOptions {
	int assay_format;
	uint32_t max_len;
	bool single_primer_pcr;
	int min_max_primer_clamp;
};

// amplicon() join (amplicon_search.cpp:355-674), padlock() joins (padlock_search.cpp:130-358),
// hybrid() (probe_search.cpp:103-227); `sites` holds everything that passed the per-oligo filters.
void assemble_hits(std::vector<BoundSite> &sites, const AssembleOptions &opt, const std::vector<int> &assay_ids,
	const std::vector<int> &assay_has_primers, const std::vector<int> &assay_has_probe,
	std::vector<tnt_hit> &hits, std::string &arena);

enum class SeqMode { PcrPlus, PcrMinus, ProbePlus, ProbeMinus, PadlockMinusStrand, PadlockPlusStrand };

void hit_sequence_plan(const tnt_hit &h, int assay_format, int &start, int &stop, SeqMode &mode);

// `codes` are the seq.h codes of fragment positions [lo, lo + codes.size())
std::string render_hit_sequence(int start, int stop, SeqMode mode, int seq_len, int lo, const std::vector<uint8_t> &codes);

} // namespace tnt
