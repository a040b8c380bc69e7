/* tntb200 -- B200-native engine for the tntblast search hot path (C ABI).
 *
 * Drop-in boundary for ONE path of jgans/thermonucleotideBLAST v2.77: what happens between
 * "a target fragment has been read" and "a list of hybrid_sig hits comes back", i.e. the calls
 *
 *     dbase.hash(bio_seq.second, ...)                         tntblast_local.cpp:534
 *     amplicon(...) / padlock(...) / hybrid(...)              tntblast_local.cpp:566,584,598,616
 *                                                             (tntblast_worker.cpp:296,316,331,350)
 *
 * declared in tntblast.h:409-472.  The reference calls them once per (fragment, assay) from a
 * per-thread loop; this ABI is the batched equivalent: register every fragment once, register
 * every assay once, search, fetch flat hit records.  No torch / C++ types cross the boundary.
 *
 * All functions return 0 on success and a negative value on error; tnt_last_error() then holds
 * the message (the reference throws `const char*`, throw.h:16-17 -- the C++ shim in
 * INTEGRATION.md re-throws it).  The engine never falls back to a CPU path: if no CUDA device
 * is usable, tnt_engine_create() fails.
 */
#ifndef TNTB200_H
#define TNTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNTB200_ABI_VERSION 3

/* hybrid_sig.h:19 */
enum { TNT_ASSAY_PCR = 0, TNT_ASSAY_PROBE = 1, TNT_ASSAY_PADLOCK = 2, TNT_ASSAY_MIPS = 3 };
/* seq.h:36-40 */
enum { TNT_STRAND_PLUS = 1, TNT_STRAND_MINUS = 2, TNT_STRAND_BOTH = 3 };
/* hybrid_sig.h:118 */
enum { TNT_PLUS = 0, TNT_MINUS = 1 };
/* which input oligo an output slot refers to */
enum { TNT_OLIGO_F = 0, TNT_OLIGO_R = 1, TNT_OLIGO_P = 2, TNT_OLIGO_NONE = -1 };

#define TNT_MAX_OLIGO_LEN 56  /* longest oligo the sm_100a kernels accept (reference: 1024) */

typedef struct tnt_engine tnt_engine;

/* NucCruc state that the reference sets once per thread (tntblast_local.cpp:363-367):
 * NucCruc melt(param_set, opt.target_t); melt.Salt(opt.salt); melt.dangle(..); melt.dinkelbach(..)
 * plus DNAHash dbase(opt.hash_word_size) (:345). */
typedef struct {
	float target_T;      /* K, default 310.15 (tntblast.h:64) */
	float salt;          /* [Na+] M, default 0.05 (tntblast.h:61) */
	int32_t dangle5;     /* default 0 (tntblast.h:72-73) */
	int32_t dangle3;
	int32_t dinkelbach;  /* default 0; 1: NucCruc::dinkelbach(true), the iterative Tm of nuc_cruc.cpp:2399-2440,
	                      * :2459-2500, :2548-2588 -- every window then runs through the generic alignment kernel */
	int32_t word_size;   /* hash word size W, 3..8, default 7 (tntblast.h:68) */
	int32_t device;      /* CUDA device ordinal */
	int32_t reserved;    /* flags, TNT_ENGINE_*; 0 = behave exactly like the reference */
} tnt_engine_params;

/* amplicon() culls its match list between the binding steps (cull_oligo_match,
 * amplicon_search.cpp:679-765).  When two bound sites of one assay overlap (e.g. a primer that
 * binds both strands of a palindromic site), the cull's mixed ordering can drop a site that is part
 * of a real amplicon.  By default the engine reproduces this (the affected groups are searched again
 * step by step like the reference does); with this flag it reports every amplicon its sites allow. */
#define TNT_ENGINE_KEEP_CULLED_SITES 1

/* Scalars the reference passes to amplicon()/padlock()/hybrid() on every call
 * (tntblast.h:409-472; values from Options, options.h:25-76). */
typedef struct {
	int32_t assay_format;            /* TNT_ASSAY_* (opt.assay_format) */
	float forward_primer_strand;     /* opt.asymmetric_strand_ratio*opt.primer_strand */
	float reverse_primer_strand;     /* opt.primer_strand */
	float probe_strand;              /* opt.probe_strand */
	float min_primer_tm, max_primer_tm, min_primer_dg, max_primer_dg;
	float min_probe_tm, max_probe_tm, min_probe_dg, max_probe_dg;
	uint32_t primer_clamp;
	int32_t min_max_primer_clamp;    /* -1 disables */
	uint32_t probe_clamp_5, probe_clamp_3;
	uint32_t max_gap, max_mismatch, max_poly_degen;
	uint32_t max_len;                /* opt.max_len (PCR amplicon / MIPS gap) */
	int32_t single_primer_pcr;
	int32_t target_strand;           /* TNT_STRAND_* (probe / padlock modes) */
} tnt_search_options;

/* One assay == one hybrid_sig after degenerate expansion (hybrid_sig.h:28-446): NUL-terminated
 * ASCII oligos (NULL or "" when absent) and the *_degen multiplicities that divide the strand
 * concentrations (amplicon_search.cpp:85-87). */
typedef struct {
	int32_t id;                      /* hybrid_sig::my_id(), echoed back in every hit */
	const char *forward;
	const char *reverse;
	const char *probe;
	int32_t forward_degen, reverse_degen, probe_degen;
} tnt_assay;

/* One bound oligo (oligo_info, tntblast.h:145-243). */
typedef struct {
	int32_t oligo;                   /* TNT_OLIGO_* : which input oligo sits in this slot */
	int32_t loc_5, loc_3;            /* fragment-local target coordinates */
	float tm, dH, dS;
	int32_t num_mm, num_gap;
	int32_t anchor_5, anchor_3;
	uint32_t align_off;              /* offset of the NUL-terminated alignment text in the arena */
} tnt_bound_oligo;

/* One hit == the fields amplicon()/padlock()/hybrid() fill in a hybrid_sig
 * (amplicon_search.cpp:447-555, padlock_search.cpp:155-222, probe_search.cpp:103-151).
 * Coordinates are fragment-local, as in the reference (the caller adds the fragment offset,
 * tntblast_local.cpp:654). */
typedef struct {
	int32_t assay_index;             /* index into the array given to tnt_engine_set_assays */
	int32_t assay_id;
	uint32_t target_id;              /* value returned by tnt_engine_add_target */
	int32_t primer_strand;           /* TNT_PLUS / TNT_MINUS */
	int32_t probe_strand;
	int32_t amp_first, amp_last;     /* amplicon_range */
	int32_t probe_first, probe_last; /* probe_range */
	tnt_bound_oligo forward;         /* output "forward" slot (after the display swap) */
	tnt_bound_oligo reverse;
	tnt_bound_oligo probe;
	int32_t forward_clamp;           /* forward_primer_clamp (anchor_3; anchor_3 of P1 for padlock) */
	int32_t reverse_clamp;           /* reverse_primer_clamp (anchor_3; anchor_5 of P2 for padlock) */
} tnt_hit;

/* Counters of the last tnt_engine_search call. */
typedef struct {
	uint64_t db_bases;               /* bases searched (sum of fragment lengths) */
	uint64_t seeds;                  /* unique (oligo, strand, diagonal) candidates emitted */
	uint64_t alignments;             /* NucCruc heterodimer alignments executed */
	uint64_t dp_cells;               /* sum of Lq*Lt over those alignments */
	uint64_t bound_sites;            /* alignments that passed every per-oligo filter */
	uint64_t hits;
	uint64_t kernel_launches;        /* CUDA kernels launched by the call */
	double scan_ms, align_ms, pair_ms, total_ms;   /* device time (CUDA events) */
	uint64_t scan_bytes;             /* algorithmic bytes of the seed scan */
	uint64_t d2h_bytes;              /* result bytes copied device -> host by the search (counters excluded) */
	uint64_t replayed_groups;        /* (fragment, assay) groups searched again step by step like amplicon() does
	                                  * (its culls can lose a site there, see TNT_ENGINE_KEEP_CULLED_SITES) */
	uint64_t undefined_dropped;      /* windows whose traceback leaves the DP matrix: the reference reads unchecked
	                                  * ring-buffer memory there (nuc_cruc.cpp:1497-1541; only when a terminal penalty
	                                  * clamps to zero, T < ~205 K), no defined answer exists; dropped and counted */
	uint64_t nonbinding_dropped;     /* windows without any alignment (Tm = 0, dG = 0) that the bounds would have
	                                  * accepted: the reference reports them with the coordinates of an earlier
	                                  * alignment (nuc_cruc.h:360-371); the engine drops and counts them */
} tnt_stats;

const char *tnt_last_error(void);
int tnt_abi_version(void);

int tnt_engine_create(const tnt_engine_params *params, tnt_engine **out);
void tnt_engine_destroy(tnt_engine *e);

/* Replaces reading a fragment + DNAHash::hash (tntblast_local.cpp:510-534).  `codes` is the
 * SEQPTR payload (seq.h:44-56): one byte per base, values 0..17 (seq.h:12-33).  The bytes are
 * staged through pinned memory, packed on the device to 2 bit/base + a 1 bit/base non-ACGT
 * mask + a sparse list of the non-ACGT codes, and stay resident in HBM until cleared. */
int tnt_engine_add_target(tnt_engine *e, const uint8_t *codes, uint32_t len, uint32_t *target_id);
/* The same for n fragments in one call (ids first_target_id .. first_target_id + n - 1): what the
 * reference's work loop (tntblast_local.cpp:400-534) does for a run of queue entries.
 * Transfers are asynchronous: when a `codes` buffer is page-locked (cudaHostAlloc /
 * cudaHostRegister) it is read by DMA after the call returns and must stay valid and unchanged
 * until the next tnt_engine_search (or any other call that reads the fragments) has returned;
 * pageable buffers are copied before the call returns. */
int tnt_engine_add_targets(tnt_engine *e, const uint8_t *const *codes, const uint32_t *lens, uint32_t n, uint32_t *first_target_id);
int tnt_engine_clear_targets(tnt_engine *e);

/* ---- FASTA text -> resident fragments (host ingest) ----
 * Replaces, for a FASTA database, sequence_data::load_fasta (sequence_data_fastx.cpp:13-79: a
 * record starts at the first '>' of a line), read_bio_seq_fasta_slow (:190-382: defline up to
 * the first '\n' / '\r'; every later character of the record that is not white space, '*' or
 * '-' is a base), ascii_to_hash_base (seq.h:148-189) and the fragment loop of the driver
 * (tntblast_local.cpp:282-289,448-468 with seq_len_increment, sequence_data.cpp:739-754: a
 * record whose *byte* length exceeds `fragment_threshold` is cut into n equal pieces, each read
 * with `overlap` extra bases on the right, tntblast_local.cpp:174,510-511).  The text is parsed
 * on the device; the caller's text must stay valid until the call returns (page-locked text is
 * read by DMA while earlier parts are being parsed).  Text in front of the first record is
 * ignored like the reference does.  A defline that is empty or not terminated (the reference
 * throws "Truncated fasta file detected!" or takes the next line for the defline) is refused;
 * after an error the set of registered fragments is undefined: call tnt_engine_clear_targets. */
typedef struct {
	uint64_t text_offset;      /* byte offset of the record's '>' */
	uint64_t text_bytes;       /* bytes up to the next record: the reference's approx_seq_len */
	uint64_t defline_offset;   /* first byte of the defline (after '>' and leading white space) */
	uint32_t defline_len;
	uint32_t n_fragments;
	uint64_t bases;            /* sequence characters of the record */
	uint32_t first_fragment;   /* index into the fragment table of this call */
	uint32_t pad;
} tnt_fasta_record;

typedef struct {
	uint32_t record;           /* index into the record table of this call */
	uint32_t start;            /* first base of the piece (reference local_target_start) */
	uint32_t stop;             /* nominal inclusive stop (local_target_stop) */
	uint32_t max_stop;         /* local_target_max_stop = text_bytes - 1 */
	uint32_t len;              /* bases held: piece + right overlap, clipped at the record end */
	uint32_t target_id;        /* engine fragment id; 0xffffffff for an empty piece (not registered) */
} tnt_fasta_fragment;

/* `fragment_threshold` 0: records are never cut.  Tables stay valid until the next
 * tnt_engine_add_fasta, clear or destroy. */
int tnt_engine_add_fasta(tnt_engine *e, const char *text, size_t nbytes, uint32_t fragment_threshold, uint32_t overlap,
	const tnt_fasta_record **records, size_t *n_records, const tnt_fasta_fragment **fragments, size_t *n_fragments);

/* Counters of the last tnt_engine_add_fasta call.  parse_ms = device time of the parser kernels
 * (k_fa_summary + k_fa_scan + k_fa_emit, CUDA events on the upload stream, summed over slabs);
 * call_ms = host wall clock of the call (text transfer included, fragment packing enqueued). */
typedef struct {
	uint64_t text_bytes, bases, records, fragments;
	uint64_t slabs, launches;
	double parse_ms, call_ms;
} tnt_ingest_stats;
int tnt_engine_get_ingest_stats(tnt_engine *e, tnt_ingest_stats *out);

/* ---- Packed database snapshot (persistent packed-DB cache) ----
 * The resident form of the registered fragments -- 2 bit/base words, 1 bit/base non-ACGT mask,
 * sparse list of the non-ACGT codes, fragment table -- copied out to caller memory and back in,
 * so that a database parsed once (tnt_engine_add_target[s] / tnt_engine_add_fasta) can be kept
 * on disk at 0.375 B/base and re-loaded without reading, parsing or packing sequence text: the
 * counterpart of DNAHash::hash (seq_hash.h:524-642) being re-run for every fragment of every run.
 * Export with NULL buffers fills `info` only (sizes to allocate).  Import needs an engine without
 * fragments (tnt_engine_clear_targets) and registers the fragments under the same ids; buffers
 * that are page-locked are read by DMA after the call returns (same rule as
 * tnt_engine_add_targets). */
typedef struct {
	uint64_t base;             /* global base index of position 0 (multiple of 64) */
	uint32_t len;
	uint32_t pad;
	uint64_t exc_begin, exc_end;
} tnt_packed_target;

typedef struct {
	uint32_t format;           /* TNTB200_PACKED_FORMAT */
	uint32_t word_size;        /* informational: the engine's seed word size */
	uint64_t n_targets;        /* tnt_packed_target records */
	uint64_t n_words;          /* uint64 words of db2 == uint32 words of nmask (32 bases each) */
	uint64_t n_exceptions;     /* entries of exc_pos / exc_code */
	uint64_t next_base;        /* first free global base index */
	uint64_t total_bases;
} tnt_packed_info;
#define TNTB200_PACKED_FORMAT 1u

int tnt_engine_export_packed(tnt_engine *e, tnt_packed_info *info, tnt_packed_target *targets, uint64_t *db2,
	uint32_t *nmask, uint64_t *exc_pos, uint8_t *exc_code);
int tnt_engine_import_packed(tnt_engine *e, const tnt_packed_info *info, const tnt_packed_target *targets,
	const uint64_t *db2, const uint32_t *nmask, const uint64_t *exc_pos, const uint8_t *exc_code);

/* seq.h codes of bases [start, start+n) of a registered fragment, read back from the packed
 * database (parity tests of the ingest path). */
int tnt_engine_target_codes(tnt_engine *e, uint32_t target_id, uint32_t start, uint32_t n, uint8_t *out);

int tnt_engine_set_assays(tnt_engine *e, const tnt_assay *assays, int32_t n);

/* All registered assays against all registered fragments: seed scan -> NucCruc alignment of
 * every candidate window -> per-oligo filters -> amplicon / padlock / probe assembly. */
int tnt_engine_search(tnt_engine *e, const tnt_search_options *opt);

/* Hits of the last search, ordered by (target_id, assay_index) and, inside one pair, in the
 * order the reference's join loops emit them.  Pointers stay valid until the next search,
 * clear or destroy.  `arena` holds the alignment strings. */
int tnt_engine_get_hits(tnt_engine *e, const tnt_hit **hits, size_t *n, const char **arena, size_t *arena_size);
int tnt_engine_get_stats(tnt_engine *e, tnt_stats *out);

/* Hits of the last search that sit on a threshold: some bound oligo has its Tm within `tm_tol`
 * (deg C) of the min / max Tm bound it was filtered with, or dH - T*dS within `dg_tol` (kcal/mol)
 * of the min / max dG bound.  These are the hits BASELINE.json's north_star wants listed
 * separately when two implementations are compared at 0.01 C / 0.001 kcal/mol: a last-digit
 * difference in Tm may move them across the filter (bind_oligo.cpp:598-640).  Writes at most
 * `cap` indices into the hit array, returns the full count. */
long tnt_engine_hits_near_threshold(tnt_engine *e, float tm_tol, float dg_tol, uint32_t *indices, size_t cap);

/* Amplicon / probe-site text of a hit as the reference builds it from the fragment
 * (amplicon_search.cpp:508-537, padlock_search.cpp:203-218,338-352, probe_search.cpp:127-143):
 * writes at most cap-1 characters + NUL, returns the full length. */
long tnt_engine_hit_sequence(tnt_engine *e, const tnt_hit *hit, char *out, size_t cap);

/* The same text for every hit of the last search at once: the fragment ranges are read back from
 * the packed database with one kernel launch and one device-to-host copy (the per-hit call above
 * costs a launch and a synchronisation each).  Hit i occupies text[offsets[i] .. offsets[i+1] - 1),
 * NUL-terminated; `offsets` has n_hits + 1 entries.  Pointers stay valid until the next search,
 * clear or destroy. */
int tnt_engine_hit_sequences(tnt_engine *e, const char **text, const uint64_t **offsets, size_t *n_hits);

/* ---- Stage-level entry points (used by the parity tests and the roofline benchmark) ---- */

/* Stage A: unique seeds of one oligo against one registered fragment, i.e. the list
 * match_oligo_to_{minus,plus}_strand builds (bind_oligo.cpp:84-122): (query_loc, target_loc),
 * sorted by diagonal.  Returns the number of seeds (which may exceed cap). */
long tnt_engine_seeds(tnt_engine *e, uint32_t target_id, const char *oligo, int32_t plus_strand,
	uint32_t *query_loc, uint32_t *target_loc, long cap);

/* Stage B: one NucCruc evaluation per explicit candidate (what bind_oligo_to_*_strand does for
 * one seed, bind_oligo.cpp:502-748 / :1205-1451, before the threshold filters). */
typedef struct {
	float tm, dH, dS, dG;
	int32_t valid;
	int32_t anchor5, anchor3;
	int32_t num_mismatch, num_gap, max_poly_degen;
	int32_t q_first, q_last, t_first, t_last;
	int32_t target_start, target_stop;
	int32_t loc_5, loc_3;
	char alignment[512];
} tnt_align_result;

int tnt_engine_align(tnt_engine *e, uint32_t target_id, const char *oligo, int32_t plus_strand,
	float strand_conc, const uint32_t *query_loc, const uint32_t *target_loc, long n,
	tnt_align_result *out);

/* ---- Oligo-only duplexes (the dimer part) ----
 * The temperatures the driver attaches to every hit but that depend on the assay's oligos only
 * (tntblast_local.cpp:657-686): `target` NULL or "": approximate_tm_homodimer of `query`
 * (nuc_cruc.cpp:2457-2516: the query against itself, symmetry entropy in the initiation term,
 * :1632), as called for forward_dimer_tm / reverse_dimer_tm / probe_dimer_tm; otherwise
 * approximate_tm_heterodimer with `target` as the second strand, 5'->3' (:2397-2455), as called
 * for primer_dimer_tm (query = forward primer, target = reverse primer).  The strand
 * concentration follows NucCruc::strand(c_a, c_b) (nuc_cruc.h:893-910).  Runs the generic
 * NucCruc kernel on the device with the second oligo as an explicit target (<= 64 bases). */
int tnt_engine_oligo_dimer(tnt_engine *e, const char *query, const char *target, float conc_a, float conc_b,
	tnt_align_result *out);


/* approximate_tm_hairpin of one oligo (nuc_cruc.cpp:2542-2618: align_hairpin :771-971, the same
 * recurrence for the oligo against itself over the triangle the steric limit leaves;
 * enumerate_hairpin_alignments :1172-1407; evaluate_hairpin_alignment :2301-2394 with the loop
 * entropy, the special tri- / tetra-loop bonuses and the terminal mismatch; Tm = dH/dS). */
int tnt_engine_oligo_hairpin(tnt_engine *e, const char *query, tnt_align_result *out);

/* Every secondary-structure temperature the driver attaches to the hits of an assay
 * (tntblast_local.cpp:657-686), for all registered assays with one kernel launch.  They depend on
 * the oligos and the strand concentrations only; -1 where the assay has no such oligo
 * (hybrid_sig::init, hybrid_sig.h:71-79).  Index 0 / 1 / 2 = forward primer / reverse primer / probe:
 *   hairpin_tm[k]      -> forward_hairpin_tm, reverse_hairpin_tm, probe_hairpin_tm
 *   homodimer_tm[k]    -> forward_dimer_tm, reverse_dimer_tm, probe_dimer_tm      (strand(c, c))
 *   heterodimer_tm[0]  -> primer_dimer_tm of a hit made by F and R                  (strand(c_f, c_r))
 *   heterodimer_tm[1], [2] -> primer_dimer_tm of a single-primer hit (F with F, R with R: the driver
 *                         runs approximate_tm_heterodimer on the oligos in the hit's two primer slots)
 * A hit whose forward slot holds the reverse primer (single-primer amplicon) takes the values of
 * that oligo, exactly as the driver looks them up by the hit's oligo strings. */
typedef struct {
	float hairpin_tm[3];
	float homodimer_tm[3];
	float heterodimer_tm[3];
} tnt_assay_structures;
int tnt_engine_assay_structures(tnt_engine *e, const tnt_search_options *opt, tnt_assay_structures *out /* one per assay */);

/* Measured 32-bit integer throughput of the device in TOP/s (lane operations), the denominator of
 * the NucCruc roofline: tops[0] independent adds, tops[1] independent min / max, tops[2] the
 * subtract-then-max pair of the DP recurrence (two operations per pair).  A few milliseconds. */
int tnt_engine_alu_peak(tnt_engine *e, double *tops /* [3] */);

/* Seed-scan-only pass over every registered fragment with the registered assays (timing aid for
 * the HBM roofline): returns the number of unique candidates and the device time. */
int tnt_engine_scan_only(tnt_engine *e, const tnt_search_options *opt, uint64_t *candidates, double *ms);

/* ---- Gathering and finishing hit lists on the host (no GPU needed) ----
 * What the reference driver does with the lists its search calls return: hits that touch a cut
 * edge of their fragment are dropped (tntblast_local.cpp:635-648), coordinates become record
 * coordinates and target_id the record index (:650-654), the lists of one assay id are joined
 * (:701-706), and per assay id: select_best_match (tntblast_util.cpp:1482-1547) if `best_match`,
 * uniquify_results (:1555-1755) if any record was cut (`uniquify_mode` < 0: decide from the
 * fragment tables; 0 / 1: never / always), sort by hybrid_sig::operator< (tntblast_local.cpp:918-930).
 * Blocks are the hit lists of any number of engines -- the shards of a database spread over several
 * GPUs -- each with the table of the fragments it holds (by target_id); pass them in shard order. */
typedef struct {
	uint32_t record;           /* index of the database record the fragment was cut from */
	uint32_t start;            /* first base of the piece in the record (local_target_start) */
	uint32_t stop;             /* nominal inclusive stop (local_target_stop) */
	uint32_t max_stop;         /* local_target_max_stop */
	uint32_t len;              /* bases held: piece + right overlap */
} tnt_fragment;

typedef struct {
	const tnt_hit *hits;       /* as tnt_engine_get_hits hands them out: ordered by (target_id, assay_index) */
	size_t n_hits;
	const char *arena;
	const tnt_fragment *fragments;
	size_t n_fragments;
} tnt_hit_block;

typedef struct {
	uint32_t block, index;     /* where the hit came from: blocks[block].hits[index] (alignment text: that block's arena) */
	tnt_hit hit;               /* record coordinates, target_id = record index */
} tnt_final_hit;

/* `assays`: the array given to tnt_engine_set_assays (ids, oligo lengths).  The result is allocated
 * by the library: release it with tnt_free.  Error text: tnt_postprocess_error(). */
int tnt_finalize_hits(const tnt_hit_block *blocks, size_t n_blocks, const tnt_assay *assays, int32_t n_assays,
	int32_t best_match, int32_t uniquify_mode, tnt_final_hit **out, size_t *n_out);
void tnt_free(void *p);
const char *tnt_postprocess_error(void);

/* ---- Host-only diagnostics (no GPU needed; used by the CPU test-suite) ---- */

/* The integer penalty table of NucCruc::update_dp_param (nuc_cruc.cpp:340-487) and the
 * best_base_pair table (nuc_cruc.cpp:14-213) as the engine uploads them: dg[49*49], bbp[18*18]. */
int tnt_debug_thermo(float T, float na, int32_t *dg, uint8_t *bbp);

/* the delta_g table of (T, na) re-derived at T_eval from the rule table the Dinkelbach kernels use */
int tnt_debug_thermo_at(float T, float na, float T_eval, int32_t *dg);

/* Compacted seed word list of an oligo (DNAHash_iterator::build_word_list, seq_hash.h:287-374);
 * returns the number of words written (at most TNT_MAX_OLIGO_LEN). */
int tnt_debug_words(const char *oligo, int32_t word_size, int32_t complement, uint16_t *words);
/* Host-side bound of the lean alignment tier: the smallest number of columns a trimmed gapless
 * alignment of `oligo` needs to reach a melting temperature of min_tm at the given conditions
 * (shorter alignments are rejected without evaluating them).  Negative on error. */
int tnt_debug_min_columns(float T, float na, const char *oligo, float strand_concentration, float min_tm);
/* The replay of amplicon()'s staged bind / cull sequence (see TNT_ENGINE_KEEP_CULLED_SITES) in its
 * array form against its literal std::list form on `cases` random match lists: returns the number of
 * cases whose hit lists differ (0 expected), the hits compared in *hits. */
long tnt_debug_replay_selftest(uint32_t seed, int32_t cases, long *hits);

#ifdef __cplusplus
}
#endif
#endif /* TNTB200_H */
