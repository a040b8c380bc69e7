// tntb200 engine: host orchestration + C ABI (include/tntb200.h).
//
// Host mirror of the reference call surface for the search hot path:
//   tnt_engine_add_target   <- read_bio_seq + DNAHash::hash         tntblast_local.cpp:510-534
//   tnt_engine_search       <- amplicon()/padlock()/hybrid() calls  tntblast_local.cpp:554-626
// with the per-(fragment, assay) loop turned into three device stages:
//   A  k_seed_scan / k_region_scan   seeds, one per (oligo strand, diagonal)
//   B  k_align                       NucCruc Tm/dG of every candidate window + per-oligo filters
//   C  assembly of amplicons / padlock sites / probe sites from the bound oligos (assemble.cpp)
// There is no CPU fallback: every stage-A/B result comes from the kernels in kernels.cuh.

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <functional>
#include <vector>

#include "../../include/tntb200.h"
#include "assemble.h"
#include "kernels.cuh"
#include "fasta.cuh"
#include "thermo.h"

using namespace tnt;

namespace {

thread_local std::string g_error;

// TNT_PROFILE=1: wall-clock of the host-side sections of a search on stderr
struct HostTimer {
	const char *name;
	std::chrono::steady_clock::time_point t0;
	static bool enabled() { static const bool on = std::getenv("TNT_PROFILE") != nullptr; return on; }
	explicit HostTimer(const char *n) : name(n), t0(std::chrono::steady_clock::now()) {}
	~HostTimer()
	{
		if (!enabled()) return;
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "[tnt] %-28s %9.3f ms\n", name, ms);
	}
};

struct CudaError : std::runtime_error {
	using std::runtime_error::runtime_error;
};

#define CUDA_OK(expr)                                                                      \
	do {                                                                                   \
		cudaError_t _e = (expr);                                                           \
		if (_e != cudaSuccess) {                                                           \
			char _b[512];                                                                  \
			snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
			throw CudaError(_b);                                                           \
		}                                                                                  \
	} while (0)

template <class T>
struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
	~DevBuf() { if (p) cudaFree(p); }
	// grow to at least n elements; `keep` elements of the old contents survive
	void reserve(size_t n, size_t keep, cudaStream_t st)
	{
		if (n <= cap) return;
		size_t ncap = std::max(n, cap + cap/2);
		T *np = nullptr;
		if (cudaMalloc(&np, ncap*sizeof(T)) != cudaSuccess) {
			cudaGetLastError();
			ncap = n; // no headroom left: exactly what is needed
			const cudaError_t err = cudaMalloc(&np, ncap*sizeof(T));
			if (err != cudaSuccess) {
				size_t free_b = 0, total_b = 0;
				cudaMemGetInfo(&free_b, &total_b);
				char msg[256];
				snprintf(msg, sizeof(msg), "out of device memory: %.1f MB requested (%zu elements of %zu bytes), %.1f MB free of %.1f MB",
					(double)ncap*sizeof(T)/1e6, ncap, sizeof(T), (double)free_b/1e6, (double)total_b/1e6);
				throw CudaError(msg);
			}
		}
		if (keep && p) CUDA_OK(cudaMemcpyAsync(np, p, keep*sizeof(T), cudaMemcpyDeviceToDevice, st));
		if (p) { CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(p)); }
		p = np;
		cap = ncap;
	}
	void upload(const std::vector<T> &v, cudaStream_t st)
	{
		reserve(std::max<size_t>(v.size(), 1), 0, st);
		if (!v.empty()) CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size()*sizeof(T), cudaMemcpyHostToDevice, st));
	}
};

// Host -> device for parameter blocks of a few KB (multiples of 4 bytes): kernel arguments, not the
// copy engine (k_poke); larger blocks fall back to an ordinary copy.
void poke(void *dst, const void *src, size_t bytes, cudaStream_t st, uint64_t *launches = nullptr)
{
	if (bytes == 0) return;
	if ((bytes & 3u) || bytes > 8*sizeof(PokePayload)) {
		CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
		return;
	}
	PokePayload p;
	for (size_t off = 0; off < bytes; off += sizeof(PokePayload)) {
		const size_t n = std::min(sizeof(PokePayload), bytes - off);
		std::memcpy(p.w, (const char *)src + off, n);
		k_poke<<<1, 256, 0, st>>>((uint32_t *)((char *)dst + off), (uint32_t)(n/4), p);
		if (launches) ++*launches;
	}
	CUDA_OK(cudaGetLastError());
}

constexpr size_t STAGE_BYTES = 32u << 20; // one upload batch
constexpr size_t MAX_SLOTS = 128;         // device staging ring (slots are allocated as needed): up to 4 GB of fragment bytes in flight

struct AssayHost {
	int id;
	std::string F, R, P;
	int fdeg, rdeg, pdeg;
};

// A set of oligo strands searched together + its k-mer lookup table
struct OsSet {
	std::vector<OligoStrand> os;
	std::vector<uint16_t> keys;     // [nos][MAX_OLIGO] little-endian keys
	std::vector<uint32_t> present, offset, entry;
	DevBuf<OligoStrand> d_os;
	DevBuf<uint16_t> d_keys;
	std::vector<uint64_t> packed;       // [nos][2] seed-orientation oligos, 2 bit/base (bit 127: contiguous word list)
	std::vector<uint32_t> assay_present; // [n_assays][nkeys/32] k-mer bitmap per assay
	DevBuf<uint32_t> d_assay_present;
	DevBuf<uint64_t> d_packed;
	DevBuf<uint32_t> d_present, d_offset, d_entry;
	DevBuf<uint16_t> d_prefix, d_doff;  // rank-compressed table (k_seed_scan_smem)
	bool smem_table = false;
	uint32_t distinct = 0;
	size_t smem_table_bytes = 0;
	std::vector<int32_t> row_tab;       // fast-kernel penalty rows, [sum of len][ROW_WORDS]
	std::vector<int32_t> lean_tab;      // lean-tier rows, [sum of len][LEAN_WORDS]
	std::vector<uint32_t> row_tab_off;  // first row of each oligo strand
	std::vector<uint8_t> lean_ok;       // the table has the structure the lean tier relies on
	DevBuf<int32_t> d_lean_tab;
	DevBuf<uint32_t> d_group_present;   // sparse scan: bitmap over (W+G-1)-mers
	int group_G = 0;                    // 0: dense scan kernel
	std::vector<uint8_t> fast_ok;       // best possible DP score fits the fast kernel's packed maximum
	DevBuf<int32_t> d_row_tab;
	DevBuf<uint32_t> d_row_tab_off;
	uint32_t nkeys = 0;
	uint64_t total_words = 0;
	int max_len = 0, max_words = 0;
};

uint32_t le_key(uint16_t word, int W)
{
	// reference words carry the first base in the high digits; the packed database is read
	// with the first base in the low bits
	uint32_t k = 0;
	for (int i = 0; i < W; ++i) k |= ((word >> (2*(W - 1 - i))) & 3u) << (2*i);
	return k;
}

} // namespace

struct tnt_engine {
	tnt_engine_params prm{};
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[8]{};

	Thermo h_thermo{};
	DevBuf<Thermo> d_thermo;
	// oligo-only duplexes (tnt_engine_oligo_dimer): explicit target of the generic kernel, and the
	// tables with the symmetry entropy folded into the initiation term for homodimers
	DevBuf<Thermo> d_thermo_homo;
	// host scratch of replay_groups, kept across searches (180 MB of fresh pages per search cost ~90 ms
	// in page faults for a hit-dense PCR search)
	std::vector<ReplaySeedRec> h_replay_seeds;
	std::vector<int32_t> h_seed_group, h_seed_site;
	std::vector<ReplaySeed> h_flat_seeds;
	DevBuf<uint8_t> d_explicit;
	DevBuf<OligoJob> d_jobs;
	DevBuf<OligoJobResult> d_job_results;
	DevBuf<uint16_t> d_job_trace;
	const Thermo *thermo_override = nullptr;
	const uint8_t *explicit_tgt = nullptr;
	int explicit_len = 0;

	// resident database
	DevBuf<uint64_t> db2;
	DevBuf<uint32_t> nmask;
	DevBuf<uint64_t> exc_pos;
	DevBuf<uint8_t> exc_code;
	uint64_t next_base = 0;     // first free global base index
	uint64_t packed_words = 0;  // db2 words written so far
	uint64_t nexc = 0;
	// open upload batch: a contiguous range of the global base space mirrored in a staging buffer
	bool batch_open = false;
	int batch_slot = 0;
	uint64_t batch_base = 0;
	uint32_t batch_used = 0;
	// Upload pipeline (own stream): a ring of device staging slots, one per batch in flight.  A
	// slot is recycled once the exceptions (non-ACGT codes) of its batch have been emitted, which
	// needs the batch's exception count on the host; emission is therefore lazy (polled), and a
	// search can start on the first batches while later ones are still crossing PCIe.
	cudaStream_t up_stream = nullptr;
	cudaStream_t emit_stream = nullptr; // exception emission: ordered by per-batch events, not behind later copies
	cudaEvent_t pads_ev = nullptr;      // read-ahead pads behind the last batch written
	struct Slot {
		uint8_t *d_stage = nullptr;
		cudaEvent_t free_ev = nullptr;    // slot contents fully consumed (pack + exception emission)
		DevBuf<uint32_t> block_count;
	};
	std::vector<Slot> slots;
	size_t max_slots = 0;             // ring size (MAX_SLOTS; TNT_UPLOAD_SLOTS shrinks it for tests)
	struct Batch { uint64_t base; uint32_t used; int slot; uint32_t nblocks; };
	// per batch (not per slot: a slot is reused within one upload when the ring wraps)
	struct BatchEvents {
		cudaEvent_t count_ev = nullptr;   // exception count of the batch is on the host
		cudaEvent_t packed_ev = nullptr;  // db2 / nmask of the batch are written
	};
	std::vector<BatchEvents> batch_ev; // grows to the largest number of batches seen, reused across uploads
	std::vector<Batch> batches;       // of the registered fragments, ascending base
	// chunks of an imported snapshot still on their way (tnt_engine_import_packed): they gate stage 1
	// like upload batches do, in front of them (an import starts at base 0)
	struct ImportChunk { uint64_t end_base; cudaEvent_t ev; };
	std::vector<ImportChunk> import_chunks;
	std::vector<cudaEvent_t> import_ev_pool;
	size_t next_emit = 0;             // batches[next_emit..] still owe their exceptions
	// host -> staging copies of the open batch, issued as one batched copy when the batch is flushed
	std::vector<void *> piece_dst, piece_src;
	std::vector<size_t> piece_size;
	bool piece_uses_mirror = false;
	cudaEvent_t emit_done = nullptr;  // recorded on emit_stream after the latest emission
	uint64_t upload_launches = 0;
	uint64_t total_bases = 0;
	std::vector<Target> targets;
	DevBuf<Target> d_targets;
	bool targets_dirty = true;
	std::vector<ScanTile> tiles;      // SCAN_TILE-sized
	DevBuf<ScanTile> d_tiles;

	uint8_t *h_stage[2] = {nullptr, nullptr}; // pinned mirrors for pageable sources (batch k uses k & 1)
	cudaEvent_t h_free[2]{};
	size_t h_mirror_batch[2] = {~(size_t)0, ~(size_t)0}; // batch each mirror currently serves
	uint64_t *h_total = nullptr; // pinned [MAX_SLOTS]
	uint64_t *d_total = nullptr; // [MAX_SLOTS]

	std::vector<AssayHost> assays;

	// FASTA ingest (tnt_engine_add_fasta): text slabs in flight, the codes of the text (1 B/base,
	// transient: the source of the fragment copies), parse state, record table
	static constexpr int FA_BUFS = 3;
	cudaStream_t fa_copy_stream = nullptr;
	uint8_t *fa_text[FA_BUFS] = {nullptr, nullptr, nullptr};
	cudaEvent_t fa_copied[FA_BUFS]{}, fa_parsed[FA_BUFS]{};
	cudaEvent_t fa_scanned = nullptr;
	DevBuf<uint8_t> fa_codes;
	DevBuf<FaTables> fa_tables;
	DevBuf<uint4> fa_block_map;
	DevBuf<FaCarry> fa_block_entry;
	DevBuf<FaCarry> fa_carry;
	FaCarry *h_fa_carry = nullptr; // pinned
	DevBuf<uint64_t> fa_rec_pos, fa_rec_base;
	std::vector<tnt_fasta_record> fa_records;
	std::vector<tnt_fasta_fragment> fa_fragments;
	std::vector<cudaEvent_t> fa_ev;   // four per slab: [summary .. scan], [emit]
	size_t fa_ev_used = 0;
	tnt_ingest_stats fa_stats{};

	// CUDA-event pairs around the alignment launches of a pass (read after the pass's own synchronisation)
	std::vector<cudaEvent_t> tev;
	size_t tev_used = 0;

	// scratch of the search
	DevBuf<Candidate> d_cand;
	DevBuf<uint32_t> d_cand_count;
	DevBuf<uint32_t> d_tile_counter;  // dense seed scan: dynamic tile distribution
	DevBuf<AlignGroup> d_groups;
	DevBuf<uint16_t> d_trace;
	DevBuf<BoundRec> d_bound;    // every site that passed the per-oligo filters, all passes of a search
	uint32_t n_bound = 0;
	DevBuf<BoundRec> d_gather;
	DevBuf<uint32_t> d_gather_idx;
	BoundHead *h_heads = nullptr; // pinned
	size_t h_heads_cap = 0;
	DevBuf<uint32_t> d_out_count;
	DevBuf<unsigned long long> d_cells;
	DevBuf<Region> d_regions;
	DevBuf<unsigned long long> d_per_assay;
	DevBuf<uint32_t> d_live;        // bitmap over (fragment, assay) groups; word 0..1 of d_live_ctl = count, error flags
	DevBuf<uint32_t> d_live_ctl;
	DevBuf<BoundHead> d_live_heads;
	DevBuf<uint32_t> d_live_index;
	uint32_t *h_live_index = nullptr; // pinned
	size_t h_live_index_cap = 0;
	DevBuf<SlowItem> d_slow;
	DevBuf<Candidate> d_retry_cand;   // hand-over to the full-trace tier, one segment per oligo strand
	DevBuf<uint32_t> d_retry_slot;
	DevBuf<uint32_t> d_retry_ctl;     // base[nos] | cap[nos] | fill[nos*COUNT_STRIDE]
	DevBuf<Candidate> d_slow_cand;
	DevBuf<uint32_t> d_slot_map;
	DevBuf<uint32_t> d_group;
	DevBuf<int32_t> d_p5;
	DevBuf<uint8_t> d_extract;

	// amplicon pairing on the device (k_pair_*)
	DevBuf<uint64_t> d_pair_key[4];   // group / location keys, each with its sort double
	DevBuf<uint32_t> d_pair_order[2];
	DevBuf<uint8_t> d_pair_alive, d_pair_tmp, d_assay_probe;
	DevBuf<PairRec> d_pairs;
	DevBuf<uint32_t> d_pair_count;
	uint64_t assay_probe_version = ~(uint64_t)0;

	// exact replay of the reference's staged PCR search (groups where its culls can lose a site)
	DevBuf<uint32_t> d_crowd_bits;
	DevBuf<CrowdRec> d_crowd_out;
	DevBuf<CandSpan> d_spans;
	DevBuf<ReplaySeedRec> d_replay_seeds;

	// oligo-strand sets of the last search, reused while assays and options stay the same
	std::unique_ptr<OsSet> set1, set2, set_all;
	tnt_search_options set_opt{};
	uint64_t assays_version = 0, set_version = ~(uint64_t)0;

	std::vector<tnt_hit> hits;
	std::string arena;
	// text of every hit (tnt_engine_hit_sequences), built on the first request after a search
	bool seq_ready = false;
	std::string seq_text;
	std::vector<uint64_t> seq_off;
	DevBuf<ExtractItem> d_extract_items;
	tnt_stats stats{};
	tnt_search_options last_opt{};

	~tnt_engine()
	{
		if (up_stream) cudaStreamSynchronize(up_stream);
		if (emit_stream) cudaStreamSynchronize(emit_stream);
		for (int i = 0; i < 2; ++i) {
			if (h_stage[i]) cudaFreeHost(h_stage[i]);
			if (h_free[i]) cudaEventDestroy(h_free[i]);
		}
		for (Slot &sl : slots) {
			if (sl.d_stage) cudaFree(sl.d_stage);
			if (sl.free_ev) cudaEventDestroy(sl.free_ev);
		}
		for (BatchEvents &b : batch_ev) {
			if (b.count_ev) cudaEventDestroy(b.count_ev);
			if (b.packed_ev) cudaEventDestroy(b.packed_ev);
		}
		if (fa_copy_stream) { cudaStreamSynchronize(fa_copy_stream); cudaStreamDestroy(fa_copy_stream); }
		for (int i = 0; i < FA_BUFS; ++i) {
			if (fa_text[i]) cudaFree(fa_text[i]);
			if (fa_copied[i]) cudaEventDestroy(fa_copied[i]);
			if (fa_parsed[i]) cudaEventDestroy(fa_parsed[i]);
		}
		if (fa_scanned) cudaEventDestroy(fa_scanned);
		for (cudaEvent_t ev : fa_ev) cudaEventDestroy(ev);
		for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
		if (h_fa_carry) cudaFreeHost(h_fa_carry);
		for (cudaEvent_t ev : import_ev_pool) cudaEventDestroy(ev);
		if (emit_done) cudaEventDestroy(emit_done);
		if (pads_ev) cudaEventDestroy(pads_ev);
		if (emit_stream) cudaStreamDestroy(emit_stream);
		if (up_stream) cudaStreamDestroy(up_stream);
		if (h_total) cudaFreeHost(h_total);
		if (h_heads) cudaFreeHost(h_heads);
		if (h_live_index) cudaFreeHost(h_live_index);
		if (d_total) cudaFree(d_total);
		for (auto &e : ev) if (e) cudaEventDestroy(e);
		if (stream) cudaStreamDestroy(stream);
	}

	DbView view() const
	{
		DbView v;
		v.db2 = db2.p; v.nmask = nmask.p; v.exc_pos = exc_pos.p; v.exc_code = exc_code.p; v.targets = d_targets.p;
		v.nexc = nexc;
		v.nwords = db2.cap;
		return v;
	}

	void finish_upload();
	void settle_upload();
	bool upload_settled = false;
	void reserve_heads(size_t n);
	// `wait_upload` false: the caller (search stage 1) orders its work behind the per-batch events
	void sync_targets(bool wait_upload = true)
	{
		if (wait_upload) finish_upload();
		if (!targets_dirty) return;
		d_targets.upload(targets, stream);
		tiles.clear();
		for (uint32_t t = 0; t < targets.size(); ++t)
			for (uint32_t s = 0; s < targets[t].len; s += SCAN_TILE) tiles.push_back(ScanTile{t, s});
		d_tiles.upload(tiles, stream);
		targets_dirty = false;
	}
};

namespace {

// ------------------------------------------------------------------------------------------
// Target upload.  Fragments are laid out back to back (64-base aligned) in a global base space;
// a staging buffer mirrors a contiguous 32 MB range of it ("batch").  Every fragment is copied
// into the device staging buffer asynchronously (straight from the caller's buffer when that is
// pinned, through the pinned host mirror otherwise); a full batch is packed by one k_pack launch.
// The sparse non-ACGT list needs the exception count of a batch: it is read back asynchronously
// and the ordered emission pass of batch i is issued while batch i+1 is being filled.
// ------------------------------------------------------------------------------------------
// Emit the exceptions of batches whose count has arrived (all of them up to `through` when
// `block`), in batch order: the list stays sorted by global position.
void emit_ready(tnt_engine *e, bool block, size_t through = ~(size_t)0)
{
	bool any = false;
	while (e->next_emit < e->batches.size()) {
		if (!block || e->next_emit > through) {
			if (!block && cudaEventQuery(e->batch_ev[e->next_emit].count_ev) != cudaSuccess) { cudaGetLastError(); break; }
			if (block && e->next_emit > through) break;
		}
		const tnt_engine::Batch &bt = e->batches[e->next_emit];
		tnt_engine::Slot &sl = e->slots[(size_t)bt.slot];
		CUDA_OK(cudaEventSynchronize(e->batch_ev[e->next_emit].count_ev));
		const uint64_t nexc = e->h_total[bt.slot];
		if (nexc) {
			if (e->nexc + nexc > e->exc_pos.cap) {
				// growing frees the old arrays: nothing on the search stream may still read them
				CUDA_OK(cudaStreamSynchronize(e->stream));
				e->exc_pos.reserve(e->nexc + nexc, e->nexc, e->emit_stream);
				e->exc_code.reserve(e->exc_pos.cap, e->nexc, e->emit_stream);
			}
			k_emit_exceptions<<<bt.nblocks, PACK_THREADS, 0, e->emit_stream>>>(sl.d_stage, bt.used, sl.block_count.p,
				e->nexc, bt.base, e->exc_pos.p, e->exc_code.p);
			CUDA_OK(cudaGetLastError());
			e->upload_launches++;
			e->nexc += nexc;
		}
		// the staging slot may be refilled once its emission has run (the count event already
		// implies that the pack kernel is done with it)
		CUDA_OK(cudaEventRecord(sl.free_ev, e->emit_stream));
		++e->next_emit;
		any = true;
	}
	if (any) CUDA_OK(cudaEventRecord(e->emit_done, e->emit_stream));
}

void flush_batch(tnt_engine *e)
{
	if (!e->batch_open) return;
	const int slot = e->batch_slot;
	const uint32_t n = e->batch_used;
	if (n) {
		tnt_engine::Slot &sl = e->slots[(size_t)slot];
		if (!e->piece_dst.empty()) {
			// all fragment pieces of the batch in one call (a 1 Gbp database is ~2000 fragments: one
			// cudaMemcpyAsync each costs more host and copy-engine time than the bytes themselves)
			cudaMemcpyAttributes attr{};
			attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
			size_t attr_idx = 0, fail = 0;
			CUDA_OK(cudaMemcpyBatchAsync(e->piece_dst.data(), e->piece_src.data(), e->piece_size.data(), e->piece_dst.size(),
				&attr, &attr_idx, 1, &fail, e->up_stream));
			if (e->piece_uses_mirror) CUDA_OK(cudaEventRecord(e->h_free[e->batches.size() & 1u], e->up_stream));
			e->piece_dst.clear();
			e->piece_src.clear();
			e->piece_size.clear();
			e->piece_uses_mirror = false;
		}
		const uint64_t first_word = e->batch_base/32u;
		const uint64_t need_words = first_word + ((uint64_t)n + 31u)/32u + 8;
		if (need_words > e->db2.cap) {
			CUDA_OK(cudaStreamSynchronize(e->stream));
			e->db2.reserve(need_words, e->packed_words, e->up_stream);
			e->nmask.reserve(e->db2.cap, e->packed_words, e->up_stream);
		}
		const uint32_t nblocks = (n + PACK_BASES_PER_BLOCK - 1)/PACK_BASES_PER_BLOCK;
		sl.block_count.reserve(nblocks, 0, e->up_stream);
		k_pack<<<nblocks, PACK_THREADS, 0, e->up_stream>>>(sl.d_stage, n, e->db2.p, e->nmask.p, first_word, sl.block_count.p);
		const size_t bi = e->batches.size();
		if (bi >= e->batch_ev.size()) {
			e->batch_ev.emplace_back();
			CUDA_OK(cudaEventCreateWithFlags(&e->batch_ev.back().count_ev, cudaEventDisableTiming));
			CUDA_OK(cudaEventCreateWithFlags(&e->batch_ev.back().packed_ev, cudaEventDisableTiming));
		}
		CUDA_OK(cudaEventRecord(e->batch_ev[bi].packed_ev, e->up_stream));
		k_scan_counts<<<1, 1024, 0, e->up_stream>>>(sl.block_count.p, nblocks, e->d_total + slot);
		CUDA_OK(cudaGetLastError());
		e->upload_launches += 2;
		CUDA_OK(cudaMemcpyAsync(e->h_total + slot, e->d_total + slot, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->up_stream));
		CUDA_OK(cudaEventRecord(e->batch_ev[bi].count_ev, e->up_stream));
		e->packed_words = first_word + ((uint64_t)n + 31u)/32u;
		e->batches.push_back(tnt_engine::Batch{e->batch_base, n, slot, nblocks});
	}
	e->batch_open = false;
	e->batch_used = 0;
	emit_ready(e, false);
}

void open_batch(tnt_engine *e, uint64_t base)
{
	// next ring slot; a slot still owned by an older batch is released by emitting that batch
	const size_t k = e->batches.size();
	const size_t slot = k % e->max_slots;
	if (slot >= e->slots.size()) {
		e->slots.emplace_back();
		tnt_engine::Slot &sl = e->slots.back();
		CUDA_OK(cudaMalloc(&sl.d_stage, STAGE_BYTES));
		CUDA_OK(cudaEventCreateWithFlags(&sl.free_ev, cudaEventDisableTiming));
	}
	else if (k >= e->max_slots) {
		emit_ready(e, true, k - e->max_slots);
		CUDA_OK(cudaEventSynchronize(e->slots[slot].free_ev));
	}
	CUDA_OK(cudaMemsetAsync(e->slots[slot].d_stage, 0, STAGE_BYTES, e->up_stream)); // alignment gaps pack as zero words
	e->batch_open = true;
	e->batch_slot = (int)slot;
	e->batch_base = base;
	e->batch_used = 0;
}

// page-locked host memory or device memory: the copy engine reads the source itself
bool is_pinned(const void *p)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice;
}

void add_target(tnt_engine *e, const uint8_t *codes, uint32_t len, uint32_t *id_out)
{
	CUDA_OK(cudaSetDevice(e->prm.device));
	if (e->targets.size() >= (1u << 24)) throw std::runtime_error("tnt_engine_add_target: too many fragments (limit 2^24)");

	Target tg{};
	tg.base = (e->next_base + 63u) & ~(uint64_t)63u; // fragments start on 64-base boundaries
	tg.len = len;

	const bool pinned = len && is_pinned(codes);
	uint64_t pos = tg.base;
	const uint8_t *src = codes;
	uint32_t rem = len;
	while (rem) {
		if (e->batch_open && pos >= e->batch_base + STAGE_BYTES) flush_batch(e);
		if (!e->batch_open) open_batch(e, pos);
		uint8_t *dst = e->slots[(size_t)e->batch_slot].d_stage;
		const uint32_t off = (uint32_t)(pos - e->batch_base);
		const uint32_t n = (uint32_t)std::min<uint64_t>(rem, STAGE_BYTES - off);
		const uint8_t *from = src;
		if (!pinned) {
			// pageable source: through one of two pinned mirrors (alternating per batch)
			const int hs = (int)(e->batches.size() & 1u);
			if (e->h_mirror_batch[hs] != e->batches.size()) {
				CUDA_OK(cudaEventSynchronize(e->h_free[hs]));
				e->h_mirror_batch[hs] = e->batches.size();
			}
			std::memcpy(e->h_stage[hs] + off, src, n);
			from = e->h_stage[hs] + off;
			e->piece_uses_mirror = true;
		}
		e->piece_dst.push_back(dst + off);
		e->piece_src.push_back(const_cast<uint8_t *>(from));
		e->piece_size.push_back(n);
		e->batch_used = off + n;
		pos += n;
		src += n;
		rem -= n;
		if (e->batch_used == STAGE_BYTES) flush_batch(e);
	}
	e->next_base = tg.base + len;
	e->total_bases += len;
	e->targets.push_back(tg);
	e->targets_dirty = true;
	e->upload_settled = false;
	if (id_out) *id_out = (uint32_t)e->targets.size() - 1;
}

// ------------------------------------------------------------------------------------------
// FASTA ingest (fasta.cuh).  The text crosses PCIe in 64 MB slabs on its own copy stream (three
// slab buffers); every slab is parsed on the upload stream: k_fa_summary, k_fa_scan, [record
// count to the host], k_fa_emit.  Records whose end is known are cut like the reference driver
// cuts them and their pieces registered straight from the device-resident codes (device ->
// staging copies, then the ordinary k_pack path), so packing overlaps the transfer of the
// following slabs.
// ------------------------------------------------------------------------------------------
void build_fasta_tables(FaTables &t)
{
	for (int c = 0; c < 256; ++c) {
		const bool lf = c == '\n', cr = c == '\r', gt = c == '>';
		const bool blank = c == ' ' || c == '\t' || c == '\v' || c == '\f';
		const bool skip = c == '*' || c == '-';
		// sequence_data_fastx.cpp:369: !isspace && != '*' && != '-' && != '\r'
		const bool base = !(lf || cr || blank || skip);
		t.lut[FS_SEQ][c] = gt ? (FS_LEAD | FA_REC_ONE) : (FS_SEQ | (base ? FA_BASE_ONE : 0u));
		t.lut[FS_LEAD][c] = blank ? FS_LEAD : lf ? (FS_SEQ | FA_ERR) : cr ? (FS_SAMELINE | FA_ERR) : FS_DEFLINE;
		t.lut[FS_DEFLINE][c] = lf ? FS_SEQ : cr ? FS_SAMELINE : FS_DEFLINE;
		t.lut[FS_SAMELINE][c] = lf ? FS_SEQ : (FS_SAMELINE | (base ? FA_BASE_ONE : 0u));
		// ascii_to_hash_base (seq.h:148-189)
		int code = 17;
		switch (std::toupper(c)) {
			case 'A': code = 0; break;  case 'C': code = 1; break;  case 'G': code = 2; break;
			case 'T': case 'U': code = 3; break;
			case 'I': code = 4; break;  case 'M': code = 5; break;  case 'R': code = 6; break;
			case 'S': code = 7; break;  case 'V': code = 8; break;  case 'W': code = 9; break;
			case 'Y': code = 10; break; case 'H': code = 11; break; case 'K': code = 12; break;
			case 'D': code = 13; break; case 'B': code = 14; break; case 'N': code = 15; break;
			case '-': code = 16; break;
			default: break;
		}
		t.code[c] = (uint8_t)code;
	}
}

// seq_len_increment(...).first (sequence_data.cpp:739-754)
uint32_t fragment_delta(uint32_t len, uint32_t max_len)
{
	if (max_len == 0 || len <= max_len) return len - 1;
	uint64_t n = 2;
	while ((uint64_t)len > n*max_len) ++n;
	return (uint32_t)(len/n + ((len % n) ? 1 : 0));
}

// Cut one record like the driver's work queue (tntblast_local.cpp:282-289,448-468) and register
// the pieces from the device-resident codes.
void register_fasta_record(tnt_engine *e, const char *text, uint64_t pos, uint64_t end, uint64_t base0, uint64_t base1,
	uint32_t threshold, uint32_t overlap)
{
	if (end - pos >= (1ull << 32)) throw std::runtime_error("tnt_engine_add_fasta: a record of 4 GB or more (the reference keeps record sizes in 32 bits)");
	tnt_fasta_record r{};
	r.text_offset = pos;
	r.text_bytes = end - pos;
	r.bases = base1 - base0;
	uint64_t p = pos + 1;
	while (p < end && std::isspace((unsigned char)text[p])) ++p;
	r.defline_offset = p;
	while (p < end && text[p] != '\n' && text[p] != '\r') ++p;
	r.defline_len = (uint32_t)(p - r.defline_offset);
	r.first_fragment = (uint32_t)e->fa_fragments.size();
	const uint32_t len = (uint32_t)r.text_bytes, max_stop = len - 1;
	const uint32_t delta = fragment_delta(len, threshold);
	uint32_t start = 0, stop = delta;
	while (true) {
		tnt_fasta_fragment f{};
		f.record = (uint32_t)e->fa_records.size();
		f.start = start;
		f.stop = stop;
		f.max_stop = max_stop;
		f.target_id = 0xffffffffu;
		if (start < r.bases) {
			const uint64_t last = std::min<uint64_t>((uint64_t)stop + overlap, r.bases - 1);
			f.len = (uint32_t)(last - start + 1);
			add_target(e, e->fa_codes.p + base0 + start, f.len, &f.target_id);
		}
		e->fa_fragments.push_back(f);
		++r.n_fragments;
		if (stop == max_stop) break;
		start = stop + 1;
		stop = (uint32_t)std::min<uint64_t>((uint64_t)stop + delta, max_stop);
	}
	e->fa_records.push_back(r);
}

void add_fasta(tnt_engine *e, const char *text, size_t nbytes, uint32_t threshold, uint32_t overlap)
{
	CUDA_OK(cudaSetDevice(e->prm.device));
	const auto wall0 = std::chrono::steady_clock::now();
	e->fa_records.clear();
	e->fa_fragments.clear();
	e->fa_ev_used = 0;
	e->fa_stats = tnt_ingest_stats{};
	e->fa_stats.text_bytes = nbytes;
	const char *gt = nbytes ? (const char *)std::memchr(text, '>', nbytes) : nullptr;
	if (!gt) return;
	const size_t first = (size_t)(gt - text), n = nbytes - first;
	const uint8_t *src = (const uint8_t *)text + first;

	if (!e->fa_copy_stream) {
		CUDA_OK(cudaStreamCreateWithFlags(&e->fa_copy_stream, cudaStreamNonBlocking));
		for (int i = 0; i < tnt_engine::FA_BUFS; ++i) {
			CUDA_OK(cudaEventCreateWithFlags(&e->fa_copied[i], cudaEventDisableTiming));
			CUDA_OK(cudaEventCreateWithFlags(&e->fa_parsed[i], cudaEventDisableTiming));
		}
		CUDA_OK(cudaEventCreateWithFlags(&e->fa_scanned, cudaEventDisableTiming));
		CUDA_OK(cudaMallocHost(&e->h_fa_carry, sizeof(FaCarry)));
		std::unique_ptr<FaTables> t(new FaTables);
		build_fasta_tables(*t);
		e->fa_tables.reserve(1, 0, e->up_stream);
		CUDA_OK(cudaMemcpyAsync(e->fa_tables.p, t.get(), sizeof(FaTables), cudaMemcpyHostToDevice, e->up_stream));
		CUDA_OK(cudaStreamSynchronize(e->up_stream));
		e->fa_carry.reserve(1, 0, e->up_stream);
	}
	// the slab buffers may still be read by the parser kernels of an earlier call
	for (int i = 0; i < tnt_engine::FA_BUFS; ++i) CUDA_OK(cudaStreamWaitEvent(e->fa_copy_stream, e->fa_parsed[i], 0));
	const size_t nslabs = (n + FA_SLAB_BYTES - 1)/FA_SLAB_BYTES;
	for (int i = 0; i < tnt_engine::FA_BUFS && (size_t)i < nslabs; ++i)
		if (!e->fa_text[i]) CUDA_OK(cudaMalloc(&e->fa_text[i], FA_SLAB_BYTES));
	// Pieces registered by an earlier call are only *queued* (piece_src points into fa_codes) until
	// their batch is flushed: issue those copies now, before fa_codes is overwritten or replaced
	// (reserve() then waits for the stream before it frees the old array).
	flush_batch(e);
	e->fa_codes.reserve(n + 16, 0, e->up_stream);
	const uint32_t max_blocks = (uint32_t)((std::min(n, FA_SLAB_BYTES) + FA_BLOCK_BYTES - 1)/FA_BLOCK_BYTES);
	e->fa_block_map.reserve(max_blocks, 0, e->up_stream);
	e->fa_block_entry.reserve(max_blocks, 0, e->up_stream);
	*e->h_fa_carry = FaCarry{0, 0, FS_SEQ, 0};
	CUDA_OK(cudaMemcpyAsync(e->fa_carry.p, e->h_fa_carry, sizeof(FaCarry), cudaMemcpyHostToDevice, e->up_stream));
	CUDA_OK(cudaStreamSynchronize(e->up_stream)); // the pinned carry is reused as the download target

	auto slab_bytes = [&](size_t k) { return std::min(FA_SLAB_BYTES, n - k*FA_SLAB_BYTES); };
	auto enqueue_copy = [&](size_t k) {
		const int b = (int)(k % tnt_engine::FA_BUFS);
		if (k >= (size_t)tnt_engine::FA_BUFS) CUDA_OK(cudaStreamWaitEvent(e->fa_copy_stream, e->fa_parsed[b], 0));
		CUDA_OK(cudaMemcpyAsync(e->fa_text[b], src + k*FA_SLAB_BYTES, slab_bytes(k), cudaMemcpyHostToDevice, e->fa_copy_stream));
		CUDA_OK(cudaEventRecord(e->fa_copied[b], e->fa_copy_stream));
	};

	std::vector<uint64_t> rpos, rbase; // record table on the host, grows slab by slab
	size_t registered = 0;             // records already cut and registered
	auto register_known = [&](uint64_t end_pos, uint64_t end_base, bool final_record) {
		// a record ends where the next one starts; the last one at the end of the text
		while (registered + 1 < rpos.size() || (final_record && registered < rpos.size())) {
			const bool last = registered + 1 == rpos.size();
			register_fasta_record(e, text, rpos[registered], last ? end_pos : rpos[registered + 1],
				rbase[registered], last ? end_base : rbase[registered + 1], threshold, overlap);
			++registered;
		}
	};

	auto mark = [&]() {
		if (e->fa_ev_used == e->fa_ev.size()) {
			e->fa_ev.emplace_back();
			CUDA_OK(cudaEventCreate(&e->fa_ev.back()));
		}
		CUDA_OK(cudaEventRecord(e->fa_ev[e->fa_ev_used++], e->up_stream));
	};
	size_t copies = 0;
	for (; copies < nslabs && copies < 2; ++copies) enqueue_copy(copies);
	for (size_t k = 0; k < nslabs; ++k) {
		const int b = (int)(k % tnt_engine::FA_BUFS);
		const uint32_t m = (uint32_t)slab_bytes(k);
		const uint32_t nblocks = (m + FA_BLOCK_BYTES - 1)/FA_BLOCK_BYTES;
		CUDA_OK(cudaStreamWaitEvent(e->up_stream, e->fa_copied[b], 0));
		mark();
		k_fa_summary<<<nblocks, FA_THREADS, 0, e->up_stream>>>(e->fa_text[b], m, e->fa_tables.p, e->fa_block_map.p);
		k_fa_scan<<<1, FA_SCAN_THREADS, 0, e->up_stream>>>(e->fa_block_map.p, nblocks, e->fa_carry.p, e->fa_block_entry.p);
		CUDA_OK(cudaGetLastError());
		mark();
		CUDA_OK(cudaMemcpyAsync(e->h_fa_carry, e->fa_carry.p, sizeof(FaCarry), cudaMemcpyDeviceToHost, e->up_stream));
		CUDA_OK(cudaEventRecord(e->fa_scanned, e->up_stream));
		CUDA_OK(cudaEventSynchronize(e->fa_scanned));
		const FaCarry after = *e->h_fa_carry;
		const size_t had = rpos.size();
		if (after.recs > had) {
			e->fa_rec_pos.reserve(after.recs, had, e->up_stream);
			e->fa_rec_base.reserve(after.recs, had, e->up_stream);
		}
		mark();
		k_fa_emit<<<nblocks, FA_THREADS, 0, e->up_stream>>>(e->fa_text[b], m, (uint64_t)first + k*FA_SLAB_BYTES, e->fa_tables.p,
			e->fa_block_entry.p, e->fa_codes.p, e->fa_rec_pos.p, e->fa_rec_base.p);
		CUDA_OK(cudaGetLastError());
		mark();
		CUDA_OK(cudaEventRecord(e->fa_parsed[b], e->up_stream));
		e->upload_launches += 3;
		e->fa_stats.launches += 3;
		e->fa_stats.slabs++;
		if (copies < nslabs) enqueue_copy(copies++);
		if (after.recs > had) {
			rpos.resize(after.recs);
			rbase.resize(after.recs);
			CUDA_OK(cudaMemcpyAsync(rpos.data() + had, e->fa_rec_pos.p + had, (after.recs - had)*sizeof(uint64_t), cudaMemcpyDeviceToHost, e->up_stream));
			CUDA_OK(cudaMemcpyAsync(rbase.data() + had, e->fa_rec_base.p + had, (after.recs - had)*sizeof(uint64_t), cudaMemcpyDeviceToHost, e->up_stream));
			CUDA_OK(cudaStreamSynchronize(e->up_stream));
		}
		const bool final_slab = k + 1 == nslabs;
		if (final_slab) {
			if (after.err) throw std::runtime_error("tnt_engine_add_fasta: empty defline (the reference reader takes the next line for the defline or throws \"Truncated fasta file detected!\")");
			if (after.state == FS_DEFLINE || after.state == FS_LEAD) throw std::runtime_error("Truncated fasta file detected!");
		}
		register_known((uint64_t)nbytes, after.bases, final_slab);
		e->fa_stats.bases = after.bases;
	}
	// the queued device-to-staging copies of the last pieces source fa_codes: put them on the stream
	// before the caller can start another ingest (or anything else that reuses the code buffer)
	flush_batch(e);
	e->fa_stats.records = e->fa_records.size();
	e->fa_stats.fragments = e->fa_fragments.size();
	e->fa_stats.call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
}

} // namespace

void tnt_engine::reserve_heads(size_t n)
{
	if (n > h_heads_cap) {
		const size_t ncap = std::max<size_t>(n, h_heads_cap*2 + 4096);
		if (h_heads) cudaFreeHost(h_heads);
		h_heads = nullptr;
		CUDA_OK(cudaMallocHost(&h_heads, ncap*sizeof(BoundHead)));
		h_heads_cap = ncap;
	}
	if (n > h_live_index_cap) {
		const size_t ncap = std::max<size_t>(n, h_live_index_cap*2 + 4096);
		if (h_live_index) cudaFreeHost(h_live_index);
		h_live_index = nullptr;
		CUDA_OK(cudaMallocHost(&h_live_index, ncap*sizeof(uint32_t)));
		h_live_index_cap = ncap;
	}
}

// Everything registered so far packed, exceptions emitted, visible to the search stream.
void tnt_engine::finish_upload()
{
	if (!batch_open && next_emit == batches.size() && upload_settled) return;
	flush_batch(this);
	emit_ready(this, true);
	import_chunks.clear(); // callers of finish_upload want everything: the search stream follows the whole upload stream
	settle_upload();
}

// Pads and ordering once all batches are enqueued (does not wait for the emission of exceptions).
void tnt_engine::settle_upload()
{
	// read-ahead pad of the scan / window loads
	const uint64_t need = packed_words + 8;
	if (need > db2.cap) {
		CUDA_OK(cudaStreamSynchronize(stream));
		db2.reserve(need, packed_words, up_stream);
		nmask.reserve(db2.cap, packed_words, up_stream);
	}
	CUDA_OK(cudaMemsetAsync(db2.p + packed_words, 0, 8*sizeof(uint64_t), up_stream));
	CUDA_OK(cudaMemsetAsync(nmask.p + packed_words, 0, 8*sizeof(uint32_t), up_stream));
	CUDA_OK(cudaEventRecord(pads_ev, up_stream));
	if (exc_pos.cap == 0) { exc_pos.reserve(16, 0, emit_stream); exc_code.reserve(16, 0, emit_stream); }
	if (next_emit == batches.size() && import_chunks.empty()) {
		// fully uploaded: later work on the search stream simply follows the upload streams
		CUDA_OK(cudaEventRecord(emit_done, emit_stream));
		CUDA_OK(cudaStreamWaitEvent(stream, emit_done, 0));
		CUDA_OK(cudaStreamWaitEvent(stream, pads_ev, 0));
		upload_settled = true;
	}
}

namespace {

// ------------------------------------------------------------------------------------------
// Oligo strands
// ------------------------------------------------------------------------------------------
OligoStrand make_os(const tnt_engine *e, int assay_index, int role, bool plus, const std::string &oligo,
	float ct, float min_tm, float max_tm, float min_dg, float max_dg, uint32_t clamp5, uint32_t clamp3,
	const tnt_search_options &o)
{
	OligoStrand s{};
	const int L = (int)oligo.size();
	if (L > MAX_OLIGO) throw std::runtime_error("oligo longer than TNT_MAX_OLIGO_LEN (56) bases");
	if (L == 0) throw std::runtime_error("empty oligo");
	for (int i = 0; i < L; ++i) {
		const int b = base_from_ascii(oligo[i]);
		if (b < 0) throw std::runtime_error(":char_to_nucleic_acid: Illegal base");
		s.seq[i] = (uint8_t)b;
	}
	s.len = L;
	s.nwords = build_words(oligo.c_str(), e->prm.word_size, plus, s.words);
	s.plus = plus ? 1 : 0;
	s.assay = assay_index;
	s.role = role;
	if (!(ct > 0.0f)) throw std::runtime_error(":NucCruc::tm_dimer: Invalid strand_concentration");
	s.r_log_ct = r_log_ct(ct);
	s.min_tm = min_tm; s.max_tm = max_tm; s.min_dg = min_dg; s.max_dg = max_dg;
	s.lean_min_cols = std::getenv("TNT_NO_LEAN_SKIP") ? 0 : lean_min_columns(e->h_thermo, s, min_tm);
	s.clamp5 = clamp5; s.clamp3 = clamp3;
	s.max_gap = o.max_gap; s.max_mismatch = o.max_mismatch; s.max_poly_degen = o.max_poly_degen;
	// Bounds that accept Tm = 0 and dG = 0 also accept windows without any alignment; the kernels
	// drop those and count them (finish_alignment, tnt_stats::nonbinding_dropped).
	return s;
}

void finish_set(tnt_engine *e, OsSet &set)
{
	const int W = e->prm.word_size;
	set.nkeys = 1u << (2*W);
	const size_t nos = set.os.size();
	if (nos >= (1u << 24)) throw std::runtime_error("too many oligo strands");
	set.keys.assign(nos*MAX_OLIGO, 0);
	set.offset.assign((size_t)set.nkeys + 1, 0);
	set.present.assign((set.nkeys + 31)/32, 0);
	set.total_words = 0;
	set.max_len = 0;
	set.max_words = 0;
	for (size_t s = 0; s < nos; ++s) {
		const OligoStrand &o = set.os[s];
		set.max_len = std::max(set.max_len, o.len);
		set.max_words = std::max(set.max_words, o.nwords);
		for (int k = 0; k < o.nwords; ++k) {
			const uint32_t key = le_key(o.words[k], W);
			set.keys[s*MAX_OLIGO + k] = (uint16_t)key;
			set.offset[key + 1]++;
			set.present[key >> 5] |= 1u << (key & 31u);
			set.total_words++;
		}
	}
	for (uint32_t k = 0; k < set.nkeys; ++k) set.offset[k + 1] += set.offset[k];
	set.entry.assign(std::max<size_t>(set.total_words, 1), 0);
	std::vector<uint32_t> fill(set.offset.begin(), set.offset.end() - 1);
	for (size_t s = 0; s < nos; ++s)
		for (int k = 0; k < set.os[s].nwords; ++k)
			set.entry[fill[set.keys[s*MAX_OLIGO + k]]++] = (uint32_t)(s << 8) | (uint32_t)k;
	set.row_tab_off.assign(nos, 0);
	size_t rows = 0;
	for (size_t s = 0; s < nos; ++s) { set.row_tab_off[s] = (uint32_t)rows; rows += (size_t)set.os[s].len; }
	set.row_tab.assign(std::max<size_t>(rows*ROW_WORDS, 1), 0);
	set.lean_tab.assign(std::max<size_t>(rows*LEAN_WORDS, 1), 0);
	set.fast_ok.assign(nos, 1);
	set.lean_ok.assign(nos, 1);
	const bool no_lean = std::getenv("TNT_NO_LEAN") != nullptr; // test hook: everything through the full-trace tier
	for (size_t s = 0; s < nos; ++s) {
		int32_t *rows_s = set.row_tab.data() + (size_t)set.row_tab_off[s]*ROW_WORDS;
		build_row_tables(e->h_thermo, set.os[s], rows_s);
		set.lean_ok[s] = build_lean_tables(e->h_thermo, set.os[s], rows_s, set.lean_tab.data() + (size_t)set.row_tab_off[s]*LEAN_WORDS) && !no_lean;
		// upper bound of any cell: every row contributes at most its most favourable M-from-M term
		int64_t bound = 0;
		for (int r = 0; r < set.os[s].len; ++r) {
			int32_t best = 0;
			for (int td = 0; td < 20; ++td) best = std::max(best, -rows_s[r*ROW_WORDS + ROW_P1 + td]);
			bound += best;
		}
		if (bound >= (1 << 20)) set.fast_ok[s] = 0;
		// Dinkelbach mode: every window iterates at temperatures of its own -> generic kernel only
		if (e->h_thermo.dinkelbach) set.fast_ok[s] = 0;
	}
	set.d_row_tab.upload(set.row_tab, e->stream);
	set.d_lean_tab.upload(set.lean_tab, e->stream);
	set.d_row_tab_off.upload(set.row_tab_off, e->stream);
	set.d_os.upload(set.os, e->stream);
	// 2-bit form of every oligo whose word list has no holes (no degenerate letter): base i is the
	// first base of word i, the tail comes from the last word
	set.packed.assign(2*nos, 0);
	for (size_t s = 0; s < nos; ++s) {
		const OligoStrand &o = set.os[s];
		if (o.nwords <= 0 || o.nwords != o.len - W + 1) continue;
		uint64_t w[2] = {0, 0};
		for (int i = 0; i < o.len; ++i) {
			const int word = std::min(i, o.nwords - 1);
			const uint64_t b = (set.keys[s*MAX_OLIGO + (size_t)word] >> (2*(i - word))) & 3u;
			w[i >> 5] |= b << (2*(i & 31));
		}
		set.packed[2*s] = w[0];
		set.packed[2*s + 1] = w[1] | ((uint64_t)1 << 63);
	}
	set.d_packed.upload(set.packed, e->stream);
	if (set.nkeys >= 32) {
		const size_t words = set.nkeys/32, na = e->assays.size();
		set.assay_present.assign(words*std::max<size_t>(na, 1), 0);
		for (size_t s = 0; s < nos; ++s)
			for (int k = 0; k < set.os[s].nwords; ++k) {
				const uint32_t key = set.keys[s*MAX_OLIGO + (size_t)k];
				set.assay_present[(size_t)set.os[s].assay*words + (key >> 5)] |= 1u << (key & 31u);
			}
		set.d_assay_present.upload(set.assay_present, e->stream);
	}
	set.d_keys.upload(set.keys, e->stream);
	set.d_present.upload(set.present, e->stream);
	set.d_offset.upload(set.offset, e->stream);
	set.d_entry.upload(set.entry, e->stream);
	// rank-compressed form of the table for k_seed_scan_smem: set bits in front of each bitmap word, and
	// offsets of the keys that occur only
	set.smem_table = false;
	if (set.nkeys >= 32 && set.total_words < 65536) {
		const size_t bm_words = set.nkeys/32;
		std::vector<uint16_t> prefix(bm_words), doff;
		uint32_t rank = 0;
		for (size_t w = 0; w < bm_words; ++w) {
			prefix[w] = (uint16_t)rank;
			rank += (uint32_t)__builtin_popcount(set.present[w]);
		}
		doff.reserve(rank + 1);
		for (uint32_t k = 0; k < set.nkeys; ++k)
			if (set.offset[k + 1] > set.offset[k]) doff.push_back((uint16_t)set.offset[k]);
		doff.push_back((uint16_t)set.total_words);
		set.distinct = rank;
		set.smem_table_bytes = 16*nos + 4*std::max<size_t>(set.total_words, 1) + 4*bm_words + 2*bm_words + 2*(rank + 2) + 16;
		if (rank < 65536 && doff.size() == (size_t)rank + 1 && set.smem_table_bytes <= SCAN_SMEM_TABLE_MAX) {
			set.d_prefix.upload(prefix, e->stream);
			set.d_doff.upload(doff, e->stream);
			set.smem_table = std::getenv("TNT_SCAN_GLOBAL_TABLE") == nullptr; // test hook: the kernel with the table in L2
		}
	}

	// Few words in the table -> most positions miss -> grouped pre-filter scan (k_seed_scan_sparse)
	{
		const int G = std::max(1, std::min(4, 11 - W));
		uint64_t distinct = 0;
		for (uint32_t k = 0; k < set.nkeys; ++k) distinct += set.offset[k + 1] > set.offset[k];
		bool sparse = (double)distinct*G/(double)set.nkeys < 0.30;
		if (const char *m = std::getenv("TNT_SCAN_MODE")) sparse = std::strcmp(m, "sparse") == 0 ? true : (std::strcmp(m, "dense") == 0 ? false : sparse);
		set.group_G = 0;
		if (sparse && set.total_words) {
			const uint32_t nkeys2 = 1u << (2*(W + G - 1));
			set.d_group_present.reserve(nkeys2/32, 0, e->stream);
			k_build_group_bitmap<<<std::min<uint32_t>(nkeys2/256, 4096u), 256, 0, e->stream>>>(set.d_present.p, set.nkeys - 1, G, nkeys2, set.d_group_present.p);
			CUDA_OK(cudaGetLastError());
			set.group_G = G;
		}
	}
}

ScanArgs scan_args(tnt_engine *e, OsSet &set, uint32_t cap)
{
	ScanArgs a{};
	a.db = e->view();
	a.wt.present = set.d_present.p;
	a.wt.offset = set.d_offset.p;
	a.wt.entry = set.d_entry.p;
	a.wt.nkeys = set.nkeys;
	a.os = set.d_os.p;
	a.os_keys = set.d_keys.p;
	a.os_packed = set.d_packed.p;
	a.assay_present = set.nkeys >= 32 ? set.d_assay_present.p : nullptr;
	a.tiles = e->d_tiles.p;
	a.W = e->prm.word_size;
	a.cand = e->d_cand.p;
	a.cand_count = e->d_cand_count.p;
	a.cap = cap;
	return a;
}

// Which seed-scan kernel serves this set, and over which tile list
struct ScanPlan {
	bool sparse;
	uint32_t tile_bases;
	const std::vector<ScanTile> *tiles;
};

ScanPlan scan_plan(tnt_engine *e, const OsSet &set)
{
	ScanPlan p;
	p.sparse = set.group_G > 0;
	p.tile_bases = (uint32_t)SCAN_TILE; // both kernels walk the same 8192-base tiles
	p.tiles = &e->tiles;
	return p;
}

// Scan tiles [t0, t1) of the plan's tile list into the candidate buckets (a.cap, a.cand set by the caller)
void launch_scan(tnt_engine *e, OsSet &set, ScanArgs a, uint32_t t0, uint32_t t1)
{
	const ScanPlan plan = scan_plan(e, set);
	const uint32_t ntiles = t1 - t0;
	if (plan.sparse) {
		SparseScanArgs sa{};
		sa.s = a;
		sa.group_present = set.d_group_present.p;
		sa.G = set.group_G;
		sa.s.tiles = e->d_tiles.p;
		sa.s.tile_begin = t0;
		sa.s.tile_end = t1;
		const size_t smem = ((size_t)1 << (2*(e->prm.word_size + set.group_G - 1)))/8 + ((set.nkeys + 31)/32)*sizeof(uint32_t) +
			(size_t)(SPARSE_THREADS/32)*(SPARSE_QUEUE*sizeof(uint32_t) + 64*sizeof(StagedCand));
		static size_t sparse_attr = 0; // the opt-in is sticky per function: set it only when it has to grow
		if (smem > sparse_attr) { CUDA_OK(cudaFuncSetAttribute(k_seed_scan_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); sparse_attr = smem; }
		const uint32_t grid = std::min<uint32_t>((ntiles + SPARSE_THREADS/32 - 1)/(SPARSE_THREADS/32), (uint32_t)e->sm_count);
		k_seed_scan_sparse<<<grid, SPARSE_THREADS, smem, e->stream>>>(sa);
	}
	else {
		a.tiles = e->d_tiles.p;
		a.tile_begin = t0;
		a.tile_end = t1;
		const size_t smem = ((set.nkeys + 31)/32)*sizeof(uint32_t);
		static size_t dense_attr = 0;
		if (smem > dense_attr) { CUDA_OK(cudaFuncSetAttribute(k_seed_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); dense_attr = smem; }
		// one resident wave of CTAs; tiles beyond the first are drawn from a counter (k_seed_scan)
		static int scan_ctas_per_sm = 0;
		static size_t scan_ctas_smem = ~(size_t)0;
		if (scan_ctas_smem != smem) {
			CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&scan_ctas_per_sm, k_seed_scan, SCAN_THREADS, smem));
			scan_ctas_per_sm = std::max(scan_ctas_per_sm, 1);
			scan_ctas_smem = smem;
		}
		e->d_tile_counter.reserve(1, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_tile_counter.p, 0, sizeof(uint32_t), e->stream));
		a.tile_counter = e->d_tile_counter.p;
		if (set.smem_table) {
			// tile, k-mer table and packed oligos in shared memory: the hit path stays on the SM
			SmemScanArgs sa{};
			sa.s = a;
			sa.prefix = set.d_prefix.p;
			sa.doff = set.d_doff.p;
			sa.distinct = set.distinct;
			sa.nentries = (uint32_t)std::max<size_t>(set.total_words, 1);
			sa.nos = (uint32_t)set.os.size();
			const size_t smem2 = set.smem_table_bytes;
			static size_t smem_attr = 0;
			if (smem2 > smem_attr) { CUDA_OK(cudaFuncSetAttribute(k_seed_scan_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); smem_attr = smem2; }
			static int ctas2 = 0;
			static size_t ctas2_smem = ~(size_t)0;
			if (ctas2_smem != smem2) {
				CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas2, k_seed_scan_smem, SCAN_THREADS, smem2));
				ctas2 = std::max(ctas2, 1);
				ctas2_smem = smem2;
			}
			const uint32_t grid2 = std::min<uint32_t>(ntiles, (uint32_t)(e->sm_count*ctas2));
			k_seed_scan_smem<<<grid2, SCAN_THREADS, smem2, e->stream>>>(sa);
			CUDA_OK(cudaGetLastError());
			return;
		}
		const uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)(e->sm_count*scan_ctas_per_sm));
		k_seed_scan<<<grid, SCAN_THREADS, smem, e->stream>>>(a);
	}
	CUDA_OK(cudaGetLastError());
}

// Oligo length classes of the fast kernels (rows held in registers)
#define TNT_FAST_CLASSES(X) X(20) X(22) X(24) X(26) X(28) X(32) X(40) X(56)
// The lean tier also has the odd classes and 18 / 19 below 28: a row of padding is 4-5 % of a 20-row fill, and the
// lean fill is two thirds of a step.  The full-trace tier (two rows per trace store, 3 % of the windows) keeps the
// even classes.
#define TNT_LEAN_CLASSES(X) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(30) X(32) X(36) X(40) X(48) X(56)
#define TNT_CLASS_ENTRY(L) L,
const int kFastClasses[] = {TNT_FAST_CLASSES(TNT_CLASS_ENTRY)};
const int kLeanClasses[] = {TNT_LEAN_CLASSES(TNT_CLASS_ENTRY)};
#undef TNT_CLASS_ENTRY

// 32-bit trace words per thread of the fast tiers
uint32_t fast_trace_words(int lq, bool full)
{
	const uint32_t cols = (uint32_t)(lq + 2*NUM_FLANK);
	return full ? cols*(uint32_t)(lq/2) : cols*(uint32_t)((lq + 15)/16);
}

// dynamic shared memory of a fast kernel: the lean tier keeps its trace there
size_t fast_smem(int lq, bool full)
{
	return full ? 0 : (size_t)fast_trace_words(lq, false)*ALIGN_THREADS*sizeof(uint32_t);
}

template <int LQ>
void launch_full(const AlignArgs &a, uint32_t grid, cudaStream_t st)
{
	k_align_fast<LQ, true><<<grid, ALIGN_THREADS, 0, st>>>(a);
}

template <int LQ>
void launch_lean(const AlignArgs &a, uint32_t grid, cudaStream_t st)
{
	k_align_fast<LQ, false><<<grid, ALIGN_THREADS, fast_smem(LQ, false), st>>>(a);
}

template <int LQ>
int full_occupancy()
{
	static int n = 0;
	if (n == 0) {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_align_fast<LQ, true>, ALIGN_THREADS, 0);
		n = std::max(n, 1);
	}
	return n;
}

template <int LQ>
int lean_occupancy()
{
	static int n = 0;
	if (n == 0) {
		const size_t smem = fast_smem(LQ, false);
		CUDA_OK(cudaFuncSetAttribute(k_align_fast<LQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_align_fast<LQ, false>, ALIGN_THREADS, smem);
		n = std::max(n, 1);
	}
	return n;
}

int fast_blocks_per_sm(int lq, bool full)
{
	if (full)
		switch (lq) {
#define TNT_CLASS_CASE(L) case L: return full_occupancy<L>();
		TNT_FAST_CLASSES(TNT_CLASS_CASE)
#undef TNT_CLASS_CASE
		default: throw std::runtime_error("internal: unknown oligo length class (full-trace tier)");
		}
	switch (lq) {
#define TNT_CLASS_CASE(L) case L: return lean_occupancy<L>();
	TNT_LEAN_CLASSES(TNT_CLASS_CASE)
#undef TNT_CLASS_CASE
	default: throw std::runtime_error("internal: unknown oligo length class (lean tier)");
	}
}

// Run one alignment kernel (fast class `lq`, or the generic kernel when lq == 0) over `units`,
// appending / writing results into e->d_out.  Returns the device time.
float run_align_kernel(tnt_engine *e, OsSet &set, AlignArgs a, std::vector<AlignGroup> &groups, int lq, int max_len, bool full = false)
{
	// unit prefix over the groups
	uint32_t nunits = 0;
	for (AlignGroup &g : groups) { g.unit_prefix = nunits; nunits += (g.count + ALIGN_THREADS - 1)/ALIGN_THREADS; }
	e->d_groups.reserve(std::max<size_t>(groups.size(), 1), 0, e->stream);
	poke(e->d_groups.p, groups.data(), groups.size()*sizeof(AlignGroup), e->stream, &e->stats.kernel_launches);
	a.groups = e->d_groups.p;
	a.ngroups = (uint32_t)groups.size();
	a.nunits = nunits;
	uint32_t grid;
	size_t smem = 0;
	if (lq == 0) {
		const int max_lt = std::max(max_len + 2*NUM_FLANK, e->explicit_len);
		smem = ((TABLE*4 + NB*NB + 52 + MAX_OLIGO + 15) & ~15) + (size_t)3*(max_lt + 1)*ALIGN_THREADS*sizeof(int32_t);
		CUDA_OK(cudaFuncSetAttribute(k_align, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		int per_sm = 1;
		CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align, ALIGN_THREADS, smem));
		grid = (uint32_t)std::min<size_t>(nunits, (size_t)e->sm_count*std::max(per_sm, 1));
		a.max_lt = max_lt;
		a.trace_cells = (uint32_t)max_len*(uint32_t)max_lt;
	}
	else {
		const size_t resident = (size_t)e->sm_count*fast_blocks_per_sm(lq, full);
		grid = (uint32_t)std::min<size_t>(nunits, resident);
		// The lean tier (trace in shared memory) runs as several waves of CTAs with 4 consecutive units each
		// (2 / 3 / 4 / 8 / 16 units: 87.0 / 87.2 / 86.8 / 87.6 / 88.0 ms end to end, where the arrival-gated chunks
		// make the launches small and their last waves count)
		// instead of one persistent wave: CTAs retire every few hundred microseconds, so the
		// pack kernels of the (higher-priority) upload stream are not locked out for the whole
		// launch while fragments are still arriving.
		static const size_t per_cta = []() { const char *v = std::getenv("TNT_LEAN_UNITS"); return (size_t)(v ? std::max(1L, std::atol(v)) : 4L); }();
		if (!full && nunits > per_cta*resident) grid = (uint32_t)((nunits + per_cta - 1)/per_cta);
		a.trace_cells = fast_trace_words(lq, full);
	}
	// d_trace counts 16-bit units; the lean tier needs none (shared memory)
	if (lq == 0 || full) e->d_trace.reserve((size_t)grid*a.trace_cells*ALIGN_THREADS*(lq == 0 ? 1 : 2) + 64, 0, e->stream);
	a.trace = e->d_trace.p;
	// timed with an event pair of its own; nobody waits for the launch here (collect_align_ms)
	while (e->tev.size() < e->tev_used + 2) {
		e->tev.emplace_back();
		CUDA_OK(cudaEventCreate(&e->tev.back()));
	}
	CUDA_OK(cudaEventRecord(e->tev[e->tev_used], e->stream));
	if (lq == 0) k_align<<<grid, ALIGN_THREADS, smem, e->stream>>>(a);
	else if (full)
		switch (lq) {
#define TNT_CLASS_CASE(L) case L: launch_full<L>(a, grid, e->stream); break;
		TNT_FAST_CLASSES(TNT_CLASS_CASE)
#undef TNT_CLASS_CASE
		default: throw std::runtime_error("internal: unknown oligo length class (full-trace tier)");
		}
	else
		switch (lq) {
#define TNT_CLASS_CASE(L) case L: launch_lean<L>(a, grid, e->stream); break;
		TNT_LEAN_CLASSES(TNT_CLASS_CASE)
#undef TNT_CLASS_CASE
		default: throw std::runtime_error("internal: unknown oligo length class (lean tier)");
		}
	CUDA_OK(cudaGetLastError());
	CUDA_OK(cudaEventRecord(e->tev[e->tev_used + 1], e->stream));
	e->tev_used += 2;
	e->stats.kernel_launches++;
	(void)set;
	return 0.0f;
}

// Device time of the alignment launches since the last call; the stream must have been synchronised.
float collect_align_ms(tnt_engine *e)
{
	float total = 0;
	for (size_t i = 0; i + 1 < e->tev_used; i += 2) {
		float ms = 0;
		CUDA_OK(cudaEventElapsedTime(&ms, e->tev[i], e->tev[i + 1]));
		total += ms;
	}
	e->tev_used = 0;
	return total;
}

// Align every candidate currently in the buckets of `set`; append the survivors to `out`.
// Returns false if a bucket overflowed (the caller shrinks the chunk and retries).
bool align_buckets(tnt_engine *e, OsSet &set, uint32_t cap, uint32_t os_base, bool emit_all,
	std::vector<uint32_t> *counts_out = nullptr)
{
	const size_t nos = set.os.size();
	std::vector<uint32_t> counts(nos);
	CUDA_OK(cudaMemcpy2DAsync(counts.data(), sizeof(uint32_t), e->d_cand_count.p, COUNT_STRIDE*sizeof(uint32_t),
		sizeof(uint32_t), nos, cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	if (counts_out) *counts_out = counts;
	uint64_t total = 0;
	for (size_t s = 0; s < nos; ++s) {
		if (counts[s] > cap) return false;
		total += counts[s];
	}
	e->stats.seeds += total;
	if (total == 0) return true;
	if (emit_all && (nos != 1 || e->n_bound != 0)) throw std::runtime_error("emit_all needs a single oligo strand and an empty site buffer");

	HostTimer t_ab("  align_buckets");
	// units per fast class
	const int nclass = (int)(sizeof(kFastClasses)/sizeof(kFastClasses[0]));   // full-trace tier (even classes)
	const int nlean = (int)(sizeof(kLeanClasses)/sizeof(kLeanClasses[0]));    // lean tier (odd ones too)
	std::vector<std::vector<AlignGroup>> by_lean(nlean);
	std::vector<AlignGroup> by_generic;
	std::vector<std::vector<AlignGroup>> by_class_full(nclass); // oligo strands the lean tier cannot take
	auto class_of = [&](size_t s) { // class index of the full-trace tier; nclass: generic kernel
		int c = 0;
		while (c < nclass - 1 && kFastClasses[c] < set.os[s].len) ++c;
		return set.fast_ok[s] ? c : nclass;
	};
	auto lean_class_of = [&](size_t s) {
		int c = 0;
		while (c < nlean - 1 && kLeanClasses[c] < set.os[s].len) ++c;
		return c;
	};
	for (size_t s = 0; s < nos; ++s)
		if (counts[s]) {
			const int c = class_of(s);
			const AlignGroup g{(uint32_t)s, 0u, counts[s], 0u};
			if (c == nclass) by_generic.push_back(g);
			else if (!set.lean_ok[s]) by_class_full[c].push_back(g);
			else by_lean[lean_class_of(s)].push_back(g);
		}

	const uint32_t base_count = e->n_bound;
	// room for 1 % of the candidates to pass (0.3 % do at -e 45 -E 50): a first search should not have
	// to redo its largest pass because the site buffer started small (seen in the ncu launch list of a
	// cold search: every lean-tier launch twice)
	size_t out_cap = emit_all ? (size_t)base_count + total
		: std::max<size_t>(e->d_bound.cap, (size_t)base_count + std::max<size_t>(1u << 16, (size_t)(total/100)));
	size_t slow_cap = std::max<size_t>(e->d_slow.cap, 1u << 16);
	// Full-trace hand-over segments.  A few per cent of a strand's candidates are typical, but an
	// oligo with an internal repeat ties its maximal cells in most windows: every segment can take
	// all candidates of its strand (12 bytes each; only what is handed over is ever touched).
	std::vector<uint32_t> retry_ctl(2*nos, 0); // base | cap
	size_t retry_total = 0;
	for (size_t s = 0; s < nos; ++s) {
		retry_ctl[s] = (uint32_t)retry_total;
		retry_ctl[nos + s] = counts[s];
		retry_total += counts[s];
	}
	if (retry_total >= ((uint64_t)1 << 32)) throw std::runtime_error("internal: hand-over list too large");
	// snapshot of the DP-cell and dropped-window counters so that a retried pass is not counted twice
	unsigned long long cells_before = 0;
	uint32_t dropped_before[2] = {0, 0}; // non-binding windows, windows without a defined answer
	CUDA_OK(cudaMemcpyAsync(&cells_before, e->d_cells.p, sizeof(cells_before), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaMemcpyAsync(dropped_before, e->d_out_count.p + 3, sizeof(dropped_before), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));

	for (;;) {
		e->d_bound.reserve(out_cap, base_count, e->stream);
		out_cap = e->d_bound.cap;
		e->d_slow.reserve(slow_cap, 0, e->stream);
		e->d_retry_cand.reserve(std::max<size_t>(retry_total, 1), 0, e->stream);
		e->d_retry_slot.reserve(std::max<size_t>(retry_total, 1), 0, e->stream);
		e->d_retry_ctl.reserve(2*nos + nos*COUNT_STRIDE, 0, e->stream);
		poke(e->d_retry_ctl.p, retry_ctl.data(), 2*nos*sizeof(uint32_t), e->stream, &e->stats.kernel_launches);
		CUDA_OK(cudaMemsetAsync(e->d_retry_ctl.p + 2*nos, 0, nos*COUNT_STRIDE*sizeof(uint32_t), e->stream));
		{
			const uint32_t init[5] = {base_count, 0, 0, dropped_before[0], dropped_before[1]};
			poke(e->d_out_count.p, init, sizeof(init), e->stream, &e->stats.kernel_launches);
		}
		poke(e->d_cells.p, &cells_before, sizeof(cells_before), e->stream, &e->stats.kernel_launches);
		AlignArgs a{};
		a.db = e->view();
		a.thermo = e->thermo_override ? e->thermo_override : e->d_thermo.p;
		a.explicit_tgt = e->explicit_tgt;
		a.explicit_len = e->explicit_len;
		a.os = set.d_os.p;
		a.cand = e->d_cand.p;
		a.cap = cap;
		a.out = e->d_bound.p;
		a.os_base = os_base;
		a.out_count = e->d_out_count.p;
		a.out_cap = (uint32_t)out_cap;
		a.emit_all = emit_all ? 1 : 0;
		a.slot_map = nullptr;
		a.cells = e->d_cells.p;
		a.row_tab = set.d_row_tab.p;
		a.lean_tab = set.d_lean_tab.p;
		a.row_off = set.d_row_tab_off.p;
		a.p5_tab = e->d_p5.p;
		a.slow = e->d_slow.p;
		a.slow_count = e->d_out_count.p + 1;
		a.retry_cand = e->d_retry_cand.p;
		a.retry_slot = e->d_retry_slot.p;
		a.retry_base = e->d_retry_ctl.p;
		a.retry_cap = e->d_retry_ctl.p + nos;
		a.retry_fill = e->d_retry_ctl.p + 2*nos;
		a.slow_cap = (uint32_t)slow_cap;

		float ms = 0;
		std::unique_ptr<HostTimer> t_phase(new HostTimer("    lean / first pass"));
		for (int c = 0; c < nlean; ++c)
			if (!by_lean[c].empty()) ms += run_align_kernel(e, set, a, by_lean[c], kLeanClasses[c], set.max_len);
		for (int c = 0; c < nclass; ++c)
			if (!by_class_full[c].empty()) ms += run_align_kernel(e, set, a, by_class_full[c], kFastClasses[c], set.max_len, true);
		if (!by_generic.empty()) ms += run_align_kernel(e, set, a, by_generic, 0, set.max_len);

		uint32_t cnt[3] = {0, 0, 0};
		std::vector<uint32_t> retry_fill(nos);
		CUDA_OK(cudaMemcpyAsync(cnt, e->d_out_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaMemcpy2DAsync(retry_fill.data(), sizeof(uint32_t), e->d_retry_ctl.p + 2*nos, COUNT_STRIDE*sizeof(uint32_t),
			sizeof(uint32_t), nos, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		ms += collect_align_ms(e);
		if (cnt[1] > slow_cap) { slow_cap = (size_t)cnt[1]*5/4; continue; }
		{
			bool overflow = false;
			uint64_t n_retry = 0;
			for (size_t s = 0; s < nos; ++s) { overflow = overflow || retry_fill[s] > retry_ctl[nos + s]; n_retry += retry_fill[s]; }
			if (overflow) throw std::runtime_error("internal: hand-over segment overflow");
			cnt[2] = (uint32_t)n_retry;
		}
		if (HostTimer::enabled())
			fprintf(stderr, "[tnt]   candidates %llu, full-trace retry %u, generic %u, fast ms %.3f\n", (unsigned long long)total, cnt[2], cnt[1], ms);
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 7
		{
			uint32_t sk[2];
			CUDA_OK(cudaMemcpy(sk, e->d_out_count.p + 12, sizeof(sk), cudaMemcpyDeviceToHost));
			fprintf(stderr, "[tnt]   evaluations skipped (cumulative) %u, contradictions %u\n", sk[0], sk[1]);
		}
#endif
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 6
		{
			uint32_t why[4];
			CUDA_OK(cudaMemcpy(why, e->d_out_count.p + 8, sizeof(why), cudaMemcpyDeviceToHost));
			fprintf(stderr, "[tnt]   hand-over reasons (cumulative): tied maxima %u, tie/gap on path %u, M==0 on path %u, border %u\n", why[0], why[1], why[2], why[3]);
		}
#endif

		// Hand-over lists -> compact candidate arrays grouped by oligo strand (counting sort)
		auto regroup = [&](const DevBuf<SlowItem> &list, uint32_t n, std::vector<std::vector<AlignGroup>> *per_class,
			std::vector<AlignGroup> *flat) {
			e->d_group.reserve(3*nos + 2, 0, e->stream); // hist | start[nos+1] | fill
			uint32_t *hist = e->d_group.p, *start_d = hist + nos, *fill = start_d + nos + 1;
			e->d_slow_cand.reserve(n, 0, e->stream);
			e->d_slot_map.reserve(n, 0, e->stream);
			CUDA_OK(cudaMemsetAsync(hist, 0, nos*sizeof(uint32_t), e->stream));
			const unsigned gb = (unsigned)std::min<size_t>(((size_t)n + 255)/256, (size_t)e->sm_count*8);
			k_regroup_hist<<<gb, 256, 0, e->stream>>>(list.p, n, hist);
			k_regroup_scan<<<1, 1024, 0, e->stream>>>(hist, (uint32_t)nos, start_d, fill);
			k_regroup_scatter<<<gb, 256, 0, e->stream>>>(list.p, n, fill, e->d_slow_cand.p, e->d_slot_map.p);
			CUDA_OK(cudaGetLastError());
			e->stats.kernel_launches += 3;
			std::vector<uint32_t> start(nos + 1, 0);
			CUDA_OK(cudaMemcpyAsync(start.data(), start_d, (nos + 1)*sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			for (size_t s2 = 0; s2 < nos; ++s2) {
				if (start[s2 + 1] == start[s2]) continue;
				const AlignGroup g{(uint32_t)s2, start[s2], start[s2 + 1] - start[s2], 0u};
				if (per_class) (*per_class)[std::min(class_of(s2), nclass - 1)].push_back(g);
				if (flat) flat->push_back(g);
			}
		};

		t_phase.reset(new HostTimer("    full-trace tier"));
		if (cnt[2]) {
			// optimal path enters a gap state: full-trace variant of the fast kernel
			std::vector<std::vector<AlignGroup>> retry_units(nclass);
			for (size_t s2 = 0; s2 < nos; ++s2)
				if (retry_fill[s2]) retry_units[std::min(class_of(s2), nclass - 1)].push_back(AlignGroup{(uint32_t)s2, retry_ctl[s2], retry_fill[s2], 0u});
			AlignArgs g = a;
			g.cand = e->d_retry_cand.p;
			g.cap = 0; // units index the segments directly
			g.slot_map = emit_all ? e->d_retry_slot.p : nullptr;
			for (int c = 0; c < nclass; ++c)
				if (!retry_units[c].empty()) ms += run_align_kernel(e, set, g, retry_units[c], kFastClasses[c], set.max_len, true);
			CUDA_OK(cudaMemcpyAsync(cnt, e->d_out_count.p, 2*sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			ms += collect_align_ms(e);
			if (cnt[1] > slow_cap) { slow_cap = (size_t)cnt[1]*5/4; continue; }
		}
		t_phase.reset(new HostTimer("    generic tier"));
		if (cnt[1]) {
			// windows with IUPAC / inosine / N target bases (or no positive score): generic kernel
			std::vector<AlignGroup> units;
			regroup(e->d_slow, cnt[1], nullptr, &units);
			AlignArgs g = a;
			g.cand = e->d_slow_cand.p;
			g.cap = 0;
			g.slot_map = emit_all ? e->d_slot_map.p : nullptr;
			ms += run_align_kernel(e, set, g, units, 0, set.max_len);
			CUDA_OK(cudaMemcpyAsync(cnt, e->d_out_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			ms += collect_align_ms(e);
		}
		t_phase.reset();
		e->stats.align_ms += ms;

		uint32_t n = cnt[0];
		if (emit_all) n = base_count + (uint32_t)total;
		else if (n > out_cap) { out_cap = (size_t)n + n/4; continue; } // enlarge and redo this pass
		e->stats.alignments += total;
		e->n_bound = n;
		return true;
	}
}

// Stage A+B over all fragments for one oligo-strand set, chunked so the candidate buckets fit.
void scan_and_align(tnt_engine *e, OsSet &set, uint32_t os_base)
{
	const ScanPlan plan = scan_plan(e, set);
	const std::vector<ScanTile> &tiles = *plan.tiles;
	if (set.os.empty() || tiles.empty()) return;
	const size_t nos = set.os.size();
	const double keys = (double)set.nkeys;
	// expected candidates per base of database for the busiest oligo strand
	const double per_base = (double)set.max_words/keys;
	size_t budget_bytes = (size_t)4 << 30; // candidate buckets per pass (HBM is plentiful)
	if (const char *mb = std::getenv("TNT_CAND_BUDGET_MB")) budget_bytes = (size_t)std::max(1L, std::atol(mb)) << 20; // test hook: force many passes
	const size_t cap_budget = std::max<size_t>(budget_bytes/sizeof(Candidate)/nos, 4096);
	uint32_t tiles_per_chunk = (uint32_t)tiles.size();
	{
		// cap = 2 * expected + slack
		const double exp_per_tile = per_base*plan.tile_bases;
		const double max_tiles = ((double)cap_budget - 2048.0)/(2.0*std::max(exp_per_tile, 1e-9));
		if (max_tiles < (double)tiles_per_chunk) tiles_per_chunk = (uint32_t)std::max(1.0, max_tiles);
	}

	e->d_cand_count.reserve(nos*COUNT_STRIDE, 0, e->stream);

	// Fragments may still be on their way to the device (upload stream).  Then the pass is cut at
	// groups of upload batches and every chunk is ordered behind the batches it reads, so the
	// first chunks are scanned and aligned while the rest of the database crosses PCIe.
	// what stage 1 may have to wait for, ascending in the base space: the chunks of an imported
	// snapshot, then the upload batches
	struct Arrival { uint64_t end_base; cudaEvent_t arrived, ready; long batch; };
	std::vector<Arrival> arrivals;
	if (!e->upload_settled) {
		for (const tnt_engine::ImportChunk &c : e->import_chunks) arrivals.push_back(Arrival{c.end_base, c.ev, c.ev, -1});
		for (size_t bi = 0; bi < e->batches.size(); ++bi)
			arrivals.push_back(Arrival{e->batches[bi].base + e->batches[bi].used, e->batch_ev[bi].count_ev, e->batch_ev[bi].packed_ev, (long)bi});
	}
	struct Gate { uint32_t tile_end; size_t batch; };
	std::vector<Gate> gates;
	if (!arrivals.empty()) {
		const size_t group = 1; // arrivals per gate (32 MB of bases); a chunk takes every gate that is already open
		uint32_t t = 0;
		for (size_t b = group - 1;; b += group) {
			const size_t bi = std::min(b, arrivals.size() - 1);
			const bool last = bi + 1 == arrivals.size();
			const uint64_t gend = arrivals[bi].end_base;
			while (t < tiles.size()) {
				const Target &tg = e->targets[tiles[t].target];
				// k-mers and alignment windows read a little past the tile, never past the fragment
				const uint64_t need = tg.base + std::min<uint64_t>((uint64_t)tiles[t].start + plan.tile_bases + 192, tg.len);
				if (!last && need > gend) break;
				++t;
			}
			if (last) t = (uint32_t)tiles.size();
			if (gates.empty() || t > gates.back().tile_end) gates.push_back(Gate{t, bi});
			else gates.back().batch = bi;
			if (last) break;
		}
	}
	size_t gate = 0;

	uint32_t t0 = 0;
	while (t0 < tiles.size()) {
		uint32_t t1 = (uint32_t)std::min<size_t>(tiles.size(), (size_t)t0 + tiles_per_chunk);
		if (!gates.empty()) {
			while (gates[gate].tile_end <= t0) ++gate;
			// as far as the upload has come (at least one gate: the stream then waits for it)
			while (gate + 1 < gates.size() &&
				cudaEventQuery(arrivals[gates[gate + 1].batch].arrived) == cudaSuccess) ++gate;
			cudaGetLastError();
			t1 = std::min(t1, gates[gate].tile_end);
			const Arrival &ar = arrivals[gates[gate].batch];
			HostTimer t_gate("  upload gate (host wait)");
			if (ar.batch >= 0) emit_ready(e, true, (size_t)ar.batch); // exceptions of these batches (the host waits for their counts only)
			if (gates[gate].batch + 1 == arrivals.size()) CUDA_OK(cudaStreamWaitEvent(e->stream, e->pads_ev, 0)); // + the read-ahead pads
			CUDA_OK(cudaStreamWaitEvent(e->stream, ar.ready, 0));
			if (ar.batch >= 0) CUDA_OK(cudaStreamWaitEvent(e->stream, e->emit_done, 0));
		}
		uint32_t forced_cap = 0; // set after an overflow, when the real bucket sizes are known
		HostTimer t_chunk("  chunk (scan + align)");
		for (;;) {
			const uint32_t ntiles = t1 - t0;
			uint32_t cap = (uint32_t)std::min<double>(4.0e9, 2.0*per_base*plan.tile_bases*ntiles + 2048.0);
			cap = (uint32_t)std::min<size_t>(cap, std::max<size_t>(cap_budget, 4096));
			if (forced_cap) cap = forced_cap;
			e->d_cand.reserve(nos*(size_t)cap, 0, e->stream);
			CUDA_OK(cudaMemsetAsync(e->d_cand_count.p, 0, nos*COUNT_STRIDE*sizeof(uint32_t), e->stream));
			CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
			launch_scan(e, set, scan_args(e, set, cap), t0, t1);
			CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
			e->stats.kernel_launches++;
			const uint64_t seeds_before = e->stats.seeds;
			std::vector<uint32_t> counts;
			const bool ok = align_buckets(e, set, cap, os_base, false, &counts);
			float ms = 0;
			CUDA_OK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
			e->stats.scan_ms += ms;
			if (HostTimer::enabled()) fprintf(stderr, "[tnt]   scan of tiles [%u, %u): %.3f ms, bucket capacity %u\n", t0, t1, ms, cap);
			if (ok) {
				uint64_t bases = 0;
				for (uint32_t t = t0; t < t1; ++t)
					bases += std::min<uint32_t>(plan.tile_bases, e->targets[tiles[t].target].len - tiles[t].start);
				e->stats.scan_bytes += bases/4 + bases/8 + (e->stats.seeds - seeds_before)*sizeof(Candidate);
				break;
			}
			// A bucket overflowed (repeats, low-complexity sequence): the counters kept counting, so
			// the true sizes are known.  Re-run the pass with exact capacity if that fits the
			// budget, otherwise halve the pass; a single tile always gets what it needs.
			uint32_t need = 0;
			for (uint32_t c : counts) need = std::max(need, c);
			need += need/16 + 64;
			if (ntiles == 1 || (size_t)need*nos*sizeof(Candidate) <= std::max<size_t>(budget_bytes, (size_t)256 << 20)) forced_cap = need;
			else { t1 = t0 + std::max<uint32_t>(1, ntiles/2); forced_cap = 0; }
		}
		t0 = t1;
	}
	if (!gates.empty() && e->next_emit == e->batches.size()) { e->upload_settled = true; e->import_chunks.clear(); } // everything waited for
}

// Stage-2: scan regions [r0, r1) of e->d_regions with the given set and align what they hold
// (`worst` = most region positions of any one assay, an estimate for the bucket size).  Regions
// of one promiscuous oligo can hold far more seeds than any estimate; when the buckets of a
// pass would not fit the budget the region list is halved.
constexpr size_t REGION_CAND_BUDGET = (size_t)24 << 30;   // HBM is plentiful (180 GB); a 100 Gbp shard packs into 37.5 GB

void region_scan_and_align(tnt_engine *e, OsSet &set, uint32_t r0, uint32_t r1, uint64_t worst, uint32_t os_base)
{
	if (set.os.empty() || r1 <= r0) return;
	const size_t nos = set.os.size();
	const uint32_t nregions = r1 - r0;
	e->d_cand_count.reserve(nos*COUNT_STRIDE, 0, e->stream);
	// Size the buckets for the busiest assay: expected seeds = positions x words / 4^W; start with
	// generous slack and use the exact counts after an overflow (repeat-rich fragments can exceed
	// any estimate).
	const double expect = (double)worst*(double)set.max_words/(double)set.nkeys;
	uint32_t cap = (uint32_t)std::min<double>(4.0*expect + 4096.0, (double)REGION_CAND_BUDGET/(double)(nos*sizeof(Candidate)));
	cap = std::max<uint32_t>(cap, 4096u);
	for (;;) {
		e->d_cand.reserve(nos*(size_t)cap, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_cand_count.p, 0, nos*COUNT_STRIDE*sizeof(uint32_t), e->stream));
		RegionScanArgs ra{};
		ra.s = scan_args(e, set, cap);
		ra.regions = e->d_regions.p + r0;
		ra.nregions = nregions;
		// Regions are dealt out by a static stride, so the grid is made several times larger than the
		// resident set and the block scheduler evens out the tail (scan + region scan per step: 8.4 ms
		// with 8 CTAs per SM, 7.9 ms with 32-128; one region per CTA -- TNT_REGION_GRID=0 -- 8.2 ms)
		static const int region_mult = []() { const char *v = std::getenv("TNT_REGION_GRID"); return v ? std::atoi(v) : 64; }();
		const uint32_t grid = region_mult > 0 ? std::min<uint32_t>(nregions, (uint32_t)e->sm_count*(uint32_t)region_mult) : nregions;
		CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
		k_region_scan<<<grid, SCAN_THREADS, 0, e->stream>>>(ra);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
		e->stats.kernel_launches++;
		std::vector<uint32_t> counts;
		const bool ok = align_buckets(e, set, cap, os_base, false, &counts);
		float ms = 0;
		CUDA_OK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
		e->stats.scan_ms += ms;
		if (ok) return;
		uint32_t need = 0;
		for (uint32_t c : counts) need = std::max(need, c);
		need += need/16 + 64; // exact sizes are known after the overflow
		if (HostTimer::enabled()) fprintf(stderr, "[tnt]   region buckets overflowed: %u regions, capacity %u, needed %u\n", nregions, cap, need);
		if ((size_t)need*nos*sizeof(Candidate) > REGION_CAND_BUDGET && nregions > 1) {
			const uint32_t mid = r0 + nregions/2;
			region_scan_and_align(e, set, r0, mid, worst/2, os_base);
			region_scan_and_align(e, set, mid, r1, worst/2, os_base);
			return;
		}
		cap = need;
	}
}

// Heads (48 B) of the bound-site records [from, to) -> pinned host array
void fetch_heads(tnt_engine *e, uint32_t from, uint32_t to)
{
	e->reserve_heads(to);
	if (to > from) {
		CUDA_OK(cudaMemcpy2DAsync(e->h_heads + from, sizeof(BoundHead), e->d_bound.p + from, sizeof(BoundRec),
			sizeof(BoundHead), to - from, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
	}
}

// Oligo-strand sets for the registered assays under the given options: stage 1 = what is scanned
// over the whole database, stage 2 = what is only searched around bound stage-1 sites (PCR).
// Rebuilt only when assays or options changed.
void prepare_sets(tnt_engine *e, const tnt_search_options &o)
{
	if (o.assay_format != TNT_ASSAY_PCR && o.assay_format != TNT_ASSAY_PROBE &&
		o.assay_format != TNT_ASSAY_PADLOCK && o.assay_format != TNT_ASSAY_MIPS)
		throw std::runtime_error("unsupported assay format");
	const bool reuse_sets = e->set1 && e->set2 && e->set_version == e->assays_version &&
		std::memcmp(&e->set_opt, &o, sizeof(o)) == 0;
	if (!reuse_sets) {
		e->set1.reset(new OsSet);
		e->set2.reset(new OsSet);
		e->set_all.reset();
	}
	OsSet &stage1 = *e->set1, &stage2 = *e->set2;

	if (!reuse_sets) {

		for (size_t ai = 0; ai < e->assays.size(); ++ai) {
			const AssayHost &as = e->assays[ai];
			const bool has_primers = !as.F.empty() && !as.R.empty();
			const bool has_probe = !as.P.empty();
			const float fct = o.forward_primer_strand/as.fdeg;
			const float rct = o.reverse_primer_strand/as.rdeg;
			const float pct = o.probe_strand/as.pdeg;
			if (has_primers) {
				if (o.assay_format == TNT_ASSAY_PCR) {
					for (int plus = 0; plus < 2; ++plus) {
						OsSet &dst = plus ? stage2 : stage1;
						dst.os.push_back(make_os(e, (int)ai, TNT_OLIGO_F, plus, as.F, fct, o.min_primer_tm, o.max_primer_tm,
							o.min_primer_dg, o.max_primer_dg, 0, o.primer_clamp, o));
						dst.os.push_back(make_os(e, (int)ai, TNT_OLIGO_R, plus, as.R, rct, o.min_primer_tm, o.max_primer_tm,
							o.min_primer_dg, o.max_primer_dg, 0, o.primer_clamp, o));
					}
					if (has_probe)
						for (int plus = 0; plus < 2; ++plus)
							stage2.os.push_back(make_os(e, (int)ai, TNT_OLIGO_P, plus, as.P, pct, o.min_probe_tm, o.max_probe_tm,
								o.min_probe_dg, o.max_probe_dg, o.probe_clamp_5, o.probe_clamp_3, o));
				}
				else if (o.assay_format == TNT_ASSAY_PADLOCK || o.assay_format == TNT_ASSAY_MIPS) {
					for (int plus = 0; plus < 2; ++plus) {
						if (!(o.target_strand & (plus ? TNT_STRAND_PLUS : TNT_STRAND_MINUS))) continue;
						// upstream probe = "reverse" oligo with a 5' clamp, downstream = "forward" with a 3' clamp
						stage1.os.push_back(make_os(e, (int)ai, TNT_OLIGO_R, plus, as.R, rct, o.min_probe_tm, o.max_probe_tm,
							o.min_probe_dg, o.max_probe_dg, o.probe_clamp_5, 0, o));
						stage1.os.push_back(make_os(e, (int)ai, TNT_OLIGO_F, plus, as.F, fct, o.min_probe_tm, o.max_probe_tm,
							o.min_probe_dg, o.max_probe_dg, 0, o.probe_clamp_3, o));
					}
				}
				else throw std::runtime_error("assay with primers in PROBE format");
			}
			else if (has_probe) {
				for (int plus = 0; plus < 2; ++plus) {
					if (!(o.target_strand & (plus ? TNT_STRAND_PLUS : TNT_STRAND_MINUS))) continue;
					stage1.os.push_back(make_os(e, (int)ai, TNT_OLIGO_P, plus, as.P, pct, o.min_probe_tm, o.max_probe_tm,
						o.min_probe_dg, o.max_probe_dg, o.probe_clamp_5, o.probe_clamp_3, o));
				}
			}
		}


		{ HostTimer t("finish_set stage1"); finish_set(e, stage1); }
		if (!stage2.os.empty()) finish_set(e, stage2);
		e->set_opt = o;
		e->set_version = e->assays_version;
	}
}

// ------------------------------------------------------------------------------------------
// Exact replay of the reference's staged PCR search for a list of (fragment, assay) groups
// (assemble.h: replay_pcr_group).  Device side: every oligo strand of the group's assay is
// scanned over the whole fragment (k_region_scan with one region per group), every seed is
// aligned (the survivors join the bound-site buffer), and the seeds themselves come back to the
// host as one dense array.
// ------------------------------------------------------------------------------------------
struct GroupKey {
	uint32_t target;
	int assay;
	bool operator<(const GroupKey &o) const { return target != o.target ? target < o.target : assay < o.assay; }
	bool operator==(const GroupKey &o) const { return target == o.target && assay == o.assay; }
};

void replay_groups(tnt_engine *e, const tnt_search_options &o, const AssembleOptions &ao, const std::vector<GroupKey> &groups,
	std::vector<BoundSite> &sites, std::vector<tnt_hit> &out_hits, std::vector<HitSites> &out_refs)
{
	if (groups.empty()) return;
	OsSet &stage1 = *e->set1, &stage2 = *e->set2;
	if (!e->set_all) {
		e->set_all.reset(new OsSet);
		e->set_all->os = stage1.os;
		e->set_all->os.insert(e->set_all->os.end(), stage2.os.begin(), stage2.os.end());
		finish_set(e, *e->set_all);
	}
	OsSet &set = *e->set_all;
	const size_t nos = set.os.size();
	const size_t n_assays = e->assays.size();

	// one whole fragment per group, cut into pieces so that a few groups still fill the device
	const uint32_t piece = 16384;
	std::vector<Region> regions;
	std::vector<uint64_t> per_assay(n_assays, 0);
	for (size_t g = 0; g < groups.size(); ++g) {
		const uint32_t len = e->targets[groups[g].target].len;
		for (uint32_t s0 = 0; s0 < len; s0 += piece) regions.push_back(Region{groups[g].target, s0, std::min(len, s0 + piece), groups[g].assay});
		per_assay[(size_t)groups[g].assay] += len;
	}
	uint64_t worst = 0;
	for (uint64_t v : per_assay) worst = std::max(worst, v);
	// many groups of one assay: the seeds of a pass have to fit the candidate budget
	if (groups.size() > 1 && (double)worst*(double)set.max_words/(double)set.nkeys*4.0*(double)nos*sizeof(Candidate) > (double)REGION_CAND_BUDGET) {
		const size_t mid = groups.size()/2;
		replay_groups(e, o, ao, std::vector<GroupKey>(groups.begin(), groups.begin() + (ptrdiff_t)mid), sites, out_hits, out_refs);
		replay_groups(e, o, ao, std::vector<GroupKey>(groups.begin() + (ptrdiff_t)mid, groups.end()), sites, out_hits, out_refs);
		return;
	}
	e->d_regions.upload(regions, e->stream);

	const uint32_t rec_before = e->n_bound;
	std::unique_ptr<HostTimer> t_part(new HostTimer("  replay: scan + align"));
	e->d_cand_count.reserve(nos*COUNT_STRIDE, 0, e->stream);
	const double expect = (double)worst*(double)set.max_words/(double)set.nkeys;
	uint32_t cap = (uint32_t)std::min<double>(4.0*expect + 4096.0, (double)(1u << 28));
	std::vector<uint32_t> counts;
	for (;;) {
		e->d_cand.reserve(nos*(size_t)cap, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_cand_count.p, 0, nos*COUNT_STRIDE*sizeof(uint32_t), e->stream));
		RegionScanArgs ra{};
		ra.s = scan_args(e, set, cap);
		ra.regions = e->d_regions.p;
		ra.nregions = (uint32_t)regions.size();
		ra.whole_fragment = 1;
		const uint32_t grid = std::min<uint32_t>((uint32_t)regions.size(), (uint32_t)e->sm_count*64u);
		CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
		k_region_scan<<<grid, SCAN_THREADS, 0, e->stream>>>(ra);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
		e->stats.kernel_launches++;
		const bool ok = align_buckets(e, set, cap, 0, false, &counts);
		float ms = 0;
		CUDA_OK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
		e->stats.scan_ms += ms;
		if (ok) break;
		uint32_t need = 0;
		for (uint32_t c : counts) need = std::max(need, c);
		need += need/16 + 64;
		if ((size_t)need*nos*sizeof(Candidate) > REGION_CAND_BUDGET && groups.size() > 1) {
			t_part.reset();
			const size_t mid = groups.size()/2;
			replay_groups(e, o, ao, std::vector<GroupKey>(groups.begin(), groups.begin() + (ptrdiff_t)mid), sites, out_hits, out_refs);
			replay_groups(e, o, ao, std::vector<GroupKey>(groups.begin() + (ptrdiff_t)mid, groups.end()), sites, out_hits, out_refs);
			return;
		}
		cap = need;
	}

	// seeds -> host
	t_part.reset(new HostTimer("  replay: seeds and sites to the host"));
	std::vector<CandSpan> spans;
	uint64_t nseeds = 0;
	for (size_t s = 0; s < nos; ++s)
		if (counts[s]) { spans.push_back(CandSpan{(uint32_t)s, counts[s], nseeds}); nseeds += counts[s]; }
	std::vector<ReplaySeedRec> &seeds = e->h_replay_seeds;
	seeds.resize(nseeds);
	if (nseeds) {
		e->d_spans.upload(spans, e->stream);
		e->d_replay_seeds.reserve(nseeds, 0, e->stream);
		k_compact_cands<<<(unsigned)std::min<size_t>(spans.size(), (size_t)e->sm_count*8), 256, 0, e->stream>>>(e->d_cand.p, cap,
			e->d_spans.p, (uint32_t)spans.size(), e->d_replay_seeds.p);
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches++;
		CUDA_OK(cudaMemcpyAsync(seeds.data(), e->d_replay_seeds.p, nseeds*sizeof(ReplaySeedRec), cudaMemcpyDeviceToHost, e->stream));
		e->stats.d2h_bytes += nseeds*sizeof(ReplaySeedRec);
	}
	// heads of the sites bound in this pass
	const uint32_t rec_after = e->n_bound;
	std::vector<BoundHead> heads(rec_after - rec_before);
	if (!heads.empty()) {
		CUDA_OK(cudaMemcpy2DAsync(heads.data(), sizeof(BoundHead), e->d_bound.p + rec_before, sizeof(BoundRec),
			sizeof(BoundHead), heads.size(), cudaMemcpyDeviceToHost, e->stream));
		e->stats.d2h_bytes += heads.size()*sizeof(BoundHead);
	}
	CUDA_OK(cudaStreamSynchronize(e->stream));

	// bound sites of this pass -> `sites`, and per group a (oligo strand, position, word) -> site table
	t_part.reset(new HostTimer("  replay: grouping (sites)"));
	struct Key { uint32_t os, t, k; int idx; };
	auto key_less = [](const Key &a, const Key &b) {
		if (a.os != b.os) return a.os < b.os;
		if (a.t != b.t) return a.t < b.t;
		return a.k < b.k;
	};
	// groups of one assay, by fragment: a seed finds its group through its oligo strand's assay
	std::vector<std::vector<std::pair<uint32_t, uint32_t>>> groups_of_assay(n_assays); // (target, group index), ascending
	for (size_t g = 0; g < groups.size(); ++g) groups_of_assay[(size_t)groups[g].assay].push_back(std::make_pair(groups[g].target, (uint32_t)g));
	auto group_index = [&](int assay, uint32_t target) -> long {
		const auto &v = groups_of_assay[(size_t)assay];
		const auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(target, 0u));
		return (it != v.end() && it->first == target) ? (long)it->second : -1;
	};
	const size_t site_base = sites.size();
	std::vector<std::vector<Key>> group_sites(groups.size());
	for (size_t i = 0; i < heads.size(); ++i) {
		const BoundHead &b = heads[i];
		if (b.flags & (F_OOB | F_STACK | F_TRUNC)) throw std::runtime_error("internal: a window without a defined answer reached the host");
		sites.push_back(make_site(b, rec_before + (uint32_t)i, set.os[b.os]));
		const long g = group_index(set.os[b.os].assay, b.target);
		if (g >= 0) group_sites[(size_t)g].push_back(Key{b.os, b.t, b.k, (int)(site_base + i)});
	}
	for (std::vector<Key> &v : group_sites) std::sort(v.begin(), v.end(), key_less);

	// Seeds per group, as one array sorted by group (counting sort; the order inside a group stays span by span,
	// seed by seed, which the replay's stable list sorts depend on).  Labelling (group of a seed, and whether its
	// window bound: two binary searches) is the expensive part and runs on a few host threads over disjoint,
	// contiguous index ranges; an earlier form appended to one vector per group from all threads, and the
	// adjacent vector headers ping-ponged between the cores (270 ms for 5 M seeds of 100 PCR assays x 1 Gbp).
	t_part.reset(new HostTimer("  replay: grouping (labels)"));
	std::vector<int32_t> &seed_group = e->h_seed_group, &seed_site = e->h_seed_site;
	seed_group.assign(nseeds, -1);
	seed_site.assign(nseeds, -1);
	{
		const unsigned nt = nseeds < 200000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		auto label = [&](unsigned t) {
			for (size_t si = t; si < spans.size(); si += nt) {
				const CandSpan &sp = spans[si];
				const OligoStrand &os = set.os[sp.os];
				long last_g = -1;
				uint32_t last_target = 0xffffffffu;
				for (uint64_t i = sp.out_off; i < sp.out_off + sp.count; ++i) {
					const ReplaySeedRec &sd = seeds[i];
					const uint32_t target = sd.target_k & 0xffffffu;
					if (target != last_target) { last_g = group_index(os.assay, target); last_target = target; }
					if (last_g < 0) continue;
					seed_group[i] = (int32_t)last_g;
					const std::vector<Key> &sk = group_sites[(size_t)last_g];
					if (!sk.empty()) {
						const Key k{sp.os, sd.t, sd.target_k >> 24, 0};
						const auto it = std::lower_bound(sk.begin(), sk.end(), k, key_less);
						if (it != sk.end() && !key_less(k, *it)) seed_site[i] = it->idx;
					}
				}
			}
		};
		if (nt == 1) label(0);
		else {
			std::vector<std::thread> pool;
			for (unsigned t = 0; t < nt; ++t) pool.emplace_back(label, t);
			for (std::thread &t : pool) t.join();
		}
	}
	t_part.reset(new HostTimer("  replay: grouping (sort)"));
	// counting sort by group; every thread owns a range of groups (it reads all labels, which is cheap and
	// sequential, and writes only inside its own part of the sorted array)
	std::vector<uint64_t> group_start(groups.size() + 1, 0);
	std::vector<ReplaySeed> &flat_seeds = e->h_flat_seeds;
	{
		const unsigned nt = nseeds < 200000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		auto run = [&](const std::function<void(unsigned)> &f) {
			if (nt == 1) { f(0); return; }
			std::vector<std::thread> pool;
			for (unsigned t = 0; t < nt; ++t) pool.emplace_back(f, t);
			for (std::thread &t : pool) t.join();
		};
		auto g_lo = [&](unsigned t) { return (int32_t)(groups.size()*(size_t)t/nt); };
		run([&](unsigned t) {
			const int32_t g0 = g_lo(t), g1 = g_lo(t + 1);
			for (uint64_t i = 0; i < nseeds; ++i) {
				const int32_t g = seed_group[i];
				if (g >= g0 && g < g1) ++group_start[(size_t)g + 1];
			}
		});
		for (size_t g = 0; g < groups.size(); ++g) group_start[g + 1] += group_start[g];
		flat_seeds.resize(group_start[groups.size()]);
		run([&](unsigned t) {
			const int32_t g0 = g_lo(t), g1 = g_lo(t + 1);
			if (g0 == g1) return;
			std::vector<uint64_t> fill(group_start.begin() + g0, group_start.begin() + g1);
			for (const CandSpan &sp : spans) {
				const OligoStrand &os = set.os[sp.os];
				const int cat = os.role == TNT_OLIGO_P ? (os.plus ? 5 : 4) : (os.plus ? 2 : 0) + (os.role == TNT_OLIGO_R ? 1 : 0);
				for (uint64_t i = sp.out_off; i < sp.out_off + sp.count; ++i) {
					const int32_t g = seed_group[i];
					if (g < g0 || g >= g1) continue;
					ReplaySeed &r = flat_seeds[fill[(size_t)(g - g0)]++];
					r.cat = cat;
					r.q = seeds[i].target_k >> 24;
					r.t = seeds[i].t;
					r.site = seed_site[i];
				}
			}
		});
	}
	// the groups are independent: a few host threads when there are many, results in group order
	t_part.reset(new HostTimer("  replay: list operations"));
	const unsigned nthreads = groups.size() < 16 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
	std::vector<std::vector<tnt_hit>> part_hits(nthreads);
	std::vector<std::vector<HitSites>> part_refs(nthreads);
	std::vector<std::string> errors(nthreads);
	auto work = [&](unsigned t) {
		const size_t per = (groups.size() + nthreads - 1)/nthreads;
		try {
			for (size_t gidx = t*per; gidx < std::min(groups.size(), (t + 1)*per); ++gidx) {
				const AssayHost &as = e->assays[(size_t)groups[gidx].assay];
				replay_pcr_group(std::vector<ReplaySeed>(flat_seeds.begin() + (ptrdiff_t)group_start[gidx], flat_seeds.begin() + (ptrdiff_t)group_start[gidx + 1]),
					sites, ao, !as.P.empty(), groups[gidx].assay, as.id, part_hits[t], part_refs[t]);
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
	for (unsigned t = 0; t < nthreads; ++t) {
		out_hits.insert(out_hits.end(), part_hits[t].begin(), part_hits[t].end());
		out_refs.insert(out_refs.end(), part_refs[t].begin(), part_refs[t].end());
	}
	(void)o;
}

// Amplicon pairing of the live sites on the device (kernels.cuh: k_pair_*).  Returns the sorted order
// (position -> index into the live heads) and the index triples, ordered like the reference emits
// its hits: by (fragment, assay), then forward site, reverse site, probe site in list order.
void pair_on_device(tnt_engine *e, const tnt_search_options &o, uint32_t n_live, std::vector<uint32_t> &order, std::vector<PairRec> &pairs)
{
	order.clear();
	pairs.clear();
	if (n_live == 0) return;
	HostTimer t("pairing on the device");
	OsSet &stage1 = *e->set1, &stage2 = *e->set2;
	const size_t n = n_live;
	if (e->assay_probe_version != e->assays_version) {
		std::vector<uint8_t> has_probe(e->assays.size());
		for (size_t a = 0; a < e->assays.size(); ++a) has_probe[a] = !e->assays[a].P.empty();
		e->d_assay_probe.upload(has_probe, e->stream);
		e->assay_probe_version = e->assays_version;
	}
	for (int k = 0; k < 4; ++k) e->d_pair_key[k].reserve(n, 0, e->stream);
	for (int k = 0; k < 2; ++k) e->d_pair_order[k].reserve(n, 0, e->stream);
	e->d_pair_alive.reserve(n, 0, e->stream);
	e->d_pair_count.reserve(1, 0, e->stream);
	PairArgs a{};
	a.heads = e->d_live_heads.p;
	a.n = n_live;
	a.os1 = stage1.d_os.p;
	a.os2 = stage2.d_os.p;
	a.nos1 = (uint32_t)stage1.os.size();
	a.assay_has_probe = e->d_assay_probe.p;
	a.key_group = e->d_pair_key[0].p;
	a.key_loc = e->d_pair_key[2].p;
	a.order = e->d_pair_order[0].p;
	a.alive = e->d_pair_alive.p;
	a.pair_count = e->d_pair_count.p;
	a.max_len = (int32_t)o.max_len;
	a.single_primer_pcr = o.single_primer_pcr;
	a.min_max_primer_clamp = o.min_max_primer_clamp;
	const unsigned grid = (unsigned)std::min<size_t>((n + 255)/256, (size_t)e->sm_count*8);
	k_pair_keys<<<grid, 256, 0, e->stream>>>(a);
	// stable LSD order: by location first, then by (fragment, assay)
	size_t tmp_bytes = 0;
	CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, e->d_pair_key[2].p, e->d_pair_key[3].p, e->d_pair_order[0].p, e->d_pair_order[1].p,
		(int)n, 0, 64, e->stream));
	e->d_pair_tmp.reserve(tmp_bytes + 16, 0, e->stream);
	CUDA_OK(cub::DeviceRadixSort::SortPairs(e->d_pair_tmp.p, tmp_bytes, e->d_pair_key[2].p, e->d_pair_key[3].p, e->d_pair_order[0].p, e->d_pair_order[1].p,
		(int)n, 0, 64, e->stream));
	k_pair_gather<<<grid, 256, 0, e->stream>>>(e->d_pair_key[0].p, e->d_pair_order[1].p, n_live, e->d_pair_key[1].p);
	CUDA_OK(cub::DeviceRadixSort::SortPairs(e->d_pair_tmp.p, tmp_bytes, e->d_pair_key[1].p, e->d_pair_key[0].p, e->d_pair_order[1].p, e->d_pair_order[0].p,
		(int)n, 0, 64, e->stream));
	a.order = e->d_pair_order[0].p;
	k_pair_unique<<<grid, 256, 0, e->stream>>>(a);
	CUDA_OK(cudaGetLastError());
	e->stats.kernel_launches += 3;
	uint32_t cap = (uint32_t)std::max<size_t>(e->d_pairs.cap, 1u << 16);
	for (;;) {
		e->d_pairs.reserve(cap, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_pair_count.p, 0, sizeof(uint32_t), e->stream));
		a.pairs = e->d_pairs.p;
		a.pair_cap = cap;
		k_pair_join<<<grid, 256, 0, e->stream>>>(a);
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches++;
		uint32_t cnt = 0;
		CUDA_OK(cudaMemcpyAsync(&cnt, e->d_pair_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		if (cnt > cap) { cap = cnt + cnt/8; continue; }
		order.resize(n);
		pairs.resize(cnt);
		CUDA_OK(cudaMemcpyAsync(order.data(), e->d_pair_order[0].p, n*sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
		if (cnt) CUDA_OK(cudaMemcpyAsync(pairs.data(), e->d_pairs.p, (size_t)cnt*sizeof(PairRec), cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		e->stats.d2h_bytes += n*sizeof(uint32_t) + (uint64_t)cnt*sizeof(PairRec);
		break;
	}
	// the reference's order: groups ascending, inside a group the loops' order
	std::sort(pairs.begin(), pairs.end(), [](const PairRec &x, const PairRec &y) {
		if (x.f != y.f) return x.f < y.f;
		if (x.r != y.r) return x.r < y.r;
		return x.p < y.p;
	});
}

// Which groups with a hit does the reference possibly treat differently from a join over all bound
// sites?  Two ways the culls of amplicon() can lose a site (assemble.h):
//  (1) two bound sites of the group whose order by (loc_5, loc_3) is not strictly their order by
//      seed position (only possible within CROWD_REACH bases; `crowded` holds every site with such a
//      neighbour).  Exact duplicates -- one oligo strand, one (loc_5, loc_3), two seed diagonals --
//      are no such pair: a bind step hands back one site per range, so they never meet as bound
//      sites; they count as alternative seeds of their site in (2).
//  (2) a hit whose own sites do not validate each other in every cull: the culls see the seed
//      positions, so forward site, probe site and reverse site must follow each other strictly by
//      seed position and within reach -- for every duplicate seed of each site.
// Everything else is reported by the reference exactly as the join over all sites reports it.
std::vector<GroupKey> groups_to_replay(tnt_engine *e, const tnt_search_options &o, const std::vector<BoundSite> &sites,
	const std::vector<HitSites> &refs, std::vector<CrowdRec> &crowded, bool replay_all)
{
	auto group_of_rec = [&](const CrowdRec &c) { return GroupKey{c.target, (int)c.assay}; };
	{
		// only the groups with a hit matter: drop the rest before sorting
		std::vector<GroupKey> hit_groups;
		for (const tnt_hit &h : e->hits)
			if (hit_groups.empty() || !(hit_groups.back() == GroupKey{h.target_id, h.assay_index})) hit_groups.push_back(GroupKey{h.target_id, h.assay_index});
		size_t m = 0;
		for (const CrowdRec &c : crowded)
			if (std::binary_search(hit_groups.begin(), hit_groups.end(), group_of_rec(c))) crowded[m++] = c;
		crowded.resize(m);
	}
	std::sort(crowded.begin(), crowded.end(), [&](const CrowdRec &a, const CrowdRec &b) {
		if (a.target != b.target) return a.target < b.target;
		if (a.assay != b.assay) return a.assay < b.assay;
		if (a.loc5 != b.loc5) return a.loc5 < b.loc5;
		if (a.loc3 != b.loc3) return a.loc3 < b.loc3;
		return a.rec < b.rec;
	});
	auto group_range = [&](const GroupKey &g) {
		const auto lo = std::lower_bound(crowded.begin(), crowded.end(), g, [&](const CrowdRec &c, const GroupKey &k) { return group_of_rec(c) < k; });
		auto hi = lo;
		while (hi != crowded.end() && group_of_rec(*hi) == g) ++hi;
		return std::make_pair(lo, hi);
	};
	std::vector<GroupKey> groups;
	for (size_t i = 0; i < e->hits.size();) {
		const tnt_hit &h = e->hits[i];
		const GroupKey g{h.target_id, h.assay_index};
		size_t j = i;
		while (j < e->hits.size() && e->hits[j].target_id == g.target && e->hits[j].assay_index == g.assay) ++j;
		if (h.forward.oligo == TNT_OLIGO_NONE) { i = j; continue; } // probe-only assay in a PCR run
		bool need = replay_all;
		const auto range = group_range(g);
		// (1) neighbours whose two orders disagree
		for (auto a = range.first; !need && a != range.second; ++a)
			for (auto b = a + 1; !need && b != range.second && b->loc5 - a->loc5 <= CROWD_REACH; ++b) {
				if (a->rec == b->rec) continue; // reported twice (hash collision)
				const bool same_range = a->loc5 == b->loc5 && a->loc3 == b->loc3;
				if (same_range && a->os == b->os) continue; // duplicate seeds of one site
				if (same_range || !(a->t < b->t)) need = true; // sorted by (loc_5, loc_3): the seeds must ascend strictly
			}
		// (2) every hit validates itself, whichever duplicate seed represents a site
		auto seeds_of = [&](const BoundSite &s, uint32_t os_index, std::vector<uint32_t> &out) {
			out.assign(1, s.target_loc);
			for (auto c = range.first; c != range.second; ++c)
				if (c->os == os_index && c->loc5 == s.loc5 && c->loc3 == s.loc3 && c->t != s.target_loc) out.push_back(c->t);
		};
		std::vector<uint32_t> tf, tr, tp;
		const uint32_t threshold = o.max_len + 50;
		for (size_t k = i; !need && k < j; ++k) {
			const BoundSite &a = sites[(size_t)refs[k].forward], &b = sites[(size_t)refs[k].reverse];
			const BoundSite &f = a.plus ? b : a, &r = a.plus ? a : b;
			const BoundSite *p = refs[k].probe >= 0 ? &sites[(size_t)refs[k].probe] : nullptr;
			seeds_of(f, f.os_index, tf);
			seeds_of(r, r.os_index, tr);
			if (p) seeds_of(*p, p->os_index, tp);
			for (uint32_t x : tf)
				for (uint32_t y : tr) {
					if (!(x < y) || y - x > threshold) need = true;
					if (p) for (uint32_t z : tp) if (!(x < z && z < y)) need = true;
				}
		}
		if (need) groups.push_back(g);
		i = j;
	}
	return groups;
}

// ------------------------------------------------------------------------------------------
// Search
// ------------------------------------------------------------------------------------------
void search(tnt_engine *e, const tnt_search_options &o)
{
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->hits.clear();
	e->seq_ready = false;
	e->arena.clear();
	e->arena.push_back('\0'); // offset 0 == empty string
	e->stats = tnt_stats{};
	e->last_opt = o;
	e->stats.db_bases = e->total_bases;
	// no waiting for fragments still in flight: stage 1 orders itself behind the upload batches
	flush_batch(e);
	e->settle_upload();
	e->sync_targets(false);
	CUDA_OK(cudaMemsetAsync(e->d_cells.p, 0, sizeof(unsigned long long), e->stream));
	CUDA_OK(cudaMemsetAsync(e->d_out_count.p + 3, 0, 2*sizeof(uint32_t), e->stream));
	cudaEvent_t t_begin = e->ev[4], t_end = e->ev[5];
	CUDA_OK(cudaEventRecord(t_begin, e->stream));

	if (o.assay_format != TNT_ASSAY_PCR && o.assay_format != TNT_ASSAY_PROBE &&
		o.assay_format != TNT_ASSAY_PADLOCK && o.assay_format != TNT_ASSAY_MIPS)
		throw std::runtime_error("unsupported assay format");

	prepare_sets(e, o);
	OsSet &stage1 = *e->set1, &stage2 = *e->set2;

	e->n_bound = 0;
	{ HostTimer t("scan_and_align stage1"); scan_and_align(e, stage1, 0); }
	const uint32_t n1 = e->n_bound;
	const uint32_t nos1 = (uint32_t)stage1.os.size();
	const uint32_t n_assays = (uint32_t)e->assays.size();
	const unsigned gen_grid = (unsigned)e->sm_count*8u;

	if (!stage2.os.empty() && n1 != 0) {
		HostTimer t_s2("stage2 total");
		// Regions around the bound minus-strand primer sites, built on the device.  Overlapping
		// regions are not merged: a seed found twice is aligned twice and collapses again in the
		// per-(loc_5, loc_3) uniqueness step, exactly like duplicate windows do in the reference
		// (bind_oligo.cpp:810-826).
		e->d_regions.reserve(n1, 0, e->stream);
		e->d_per_assay.reserve(n_assays + 1, 0, e->stream);
		e->d_live_ctl.reserve(2, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_per_assay.p, 0, (n_assays + 1)*sizeof(unsigned long long), e->stream));
		CUDA_OK(cudaMemsetAsync(e->d_live_ctl.p, 0, 2*sizeof(uint32_t), e->stream));
		k_make_regions<<<gen_grid, 256, 0, e->stream>>>(e->d_bound.p, n1, stage1.d_os.p, e->d_targets.p, o.max_len,
			e->d_regions.p, e->d_per_assay.p, e->d_live_ctl.p + 1);
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches++;
		std::vector<unsigned long long> per_assay(n_assays + 1, 0);
		CUDA_OK(cudaMemcpyAsync(per_assay.data(), e->d_per_assay.p, per_assay.size()*sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		uint64_t worst = 0;
		for (unsigned long long v : per_assay) worst = std::max<uint64_t>(worst, v);
		{ HostTimer t("region_scan_and_align"); region_scan_and_align(e, stage2, 0, n1, worst, nos1); }
	}
	const uint32_t n2 = e->n_bound;

	// Bound sites with a neighbour of their own group close by (k_crowd): input of the decision which
	// groups have to be searched again step by step (groups_to_replay)
	const bool pcr_primers = o.assay_format == TNT_ASSAY_PCR && !stage2.os.empty();
	std::vector<CrowdRec> crowded;
	if (pcr_primers && n2 != 0 && !(e->prm.reserved & TNT_ENGINE_KEEP_CULLED_SITES)) {
		HostTimer t_crowd("crowded sites (k_crowd)");
		uint32_t log2_bits = 22;
		while (log2_bits < 32 && ((uint64_t)1 << log2_bits) < (uint64_t)n2*1024u) ++log2_bits;
		const size_t words = (size_t)(((uint64_t)1 << log2_bits)/32u);
		uint32_t out_cap = (uint32_t)std::max<size_t>(e->d_crowd_out.cap, 1u << 16);
		for (;;) {
			e->d_crowd_bits.reserve(2*words + 1, 0, e->stream);
			e->d_crowd_out.reserve(out_cap, 0, e->stream);
			CUDA_OK(cudaMemsetAsync(e->d_crowd_bits.p, 0, (2*words + 1)*sizeof(uint32_t), e->stream));
			CrowdArgs ca{};
			ca.recs = e->d_bound.p;
			ca.n = n2;
			ca.os1 = stage1.d_os.p;
			ca.os2 = stage2.d_os.p;
			ca.nos1 = nos1;
			ca.bits = e->d_crowd_bits.p;
			ca.log2_bits = log2_bits;
			ca.out = e->d_crowd_out.p;
			ca.out_count = e->d_crowd_bits.p + 2*words;
			ca.out_cap = out_cap;
			k_crowd<<<gen_grid, 256, 0, e->stream>>>(ca, 0);
			k_crowd<<<gen_grid, 256, 0, e->stream>>>(ca, 1);
			CUDA_OK(cudaGetLastError());
			e->stats.kernel_launches += 2;
			uint32_t cnt = 0;
			CUDA_OK(cudaMemcpyAsync(&cnt, ca.out_count, sizeof(cnt), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			if (cnt > out_cap) { out_cap = cnt + cnt/8; continue; }
			crowded.resize(cnt);
			if (cnt) {
				CUDA_OK(cudaMemcpyAsync(crowded.data(), e->d_crowd_out.p, (size_t)cnt*sizeof(CrowdRec), cudaMemcpyDeviceToHost, e->stream));
				CUDA_OK(cudaStreamSynchronize(e->stream));
				e->stats.d2h_bytes += (uint64_t)cnt*sizeof(CrowdRec);
			}
			break;
		}
	}

	// Stage C
	auto os_of = [&](uint32_t g) -> const OligoStrand & {
		return g < stage1.os.size() ? stage1.os[g] : stage2.os[g - stage1.os.size()];
	};
	// PCR: an amplicon needs a minus-strand primer site with a plus-strand primer site of the same
	// (fragment, assay) less than max_len downstream.  Only sites that can be part of such a pair
	// leave the device (k_live_*: position buckets at least max_len wide).
	const bool padlock_format = o.assay_format == TNT_ASSAY_PADLOCK || o.assay_format == TNT_ASSAY_MIPS;
	const uint64_t key_space = (uint64_t)e->targets.size()*n_assays*(padlock_format ? 2u : 1u);
	const bool prefilter = ((o.assay_format == TNT_ASSAY_PCR && !stage2.os.empty()) || padlock_format) && key_space <= ((uint64_t)1 << 31);
	uint32_t n_live = n2;
	const uint32_t *site_index = nullptr; // record index of each downloaded head (nullptr: identity)
	bool paired_on_device = false;
	std::vector<uint32_t> pair_order;     // sorted position -> live site
	std::vector<PairRec> pair_recs;
	if (prefilter && n2 != 0) {
		uint32_t max_target_len = 1;
		for (const Target &t : e->targets) max_target_len = std::max(max_target_len, t.len);
		// bucket width: >= max_len for PCR; >= gap limit + two oligos for padlock / MIPS
		const int64_t span = padlock_format ? (o.assay_format == TNT_ASSAY_MIPS ? (int64_t)o.max_len : 0) + 128 : std::max<int64_t>(o.max_len, 64);
		uint32_t shift = 6;
		while (((uint32_t)1 << shift) < (uint32_t)span && shift < 31) ++shift;
		uint32_t nbucket = (max_target_len >> shift) + 4;
		while (key_space*nbucket > ((uint64_t)1 << 33) && shift < 31) { ++shift; nbucket = (max_target_len >> shift) + 4; }
		const size_t words = (size_t)((key_space*nbucket + 31)/32) + 1;
		e->d_live.reserve(2*words, 0, e->stream);
		e->d_live_ctl.reserve(2, 0, e->stream);
		e->d_live_heads.reserve(n2, 0, e->stream);
		e->d_live_index.reserve(n2, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_live.p, 0, 2*words*sizeof(uint32_t), e->stream));
		CUDA_OK(cudaMemsetAsync(e->d_live_ctl.p, 0, 2*sizeof(uint32_t), e->stream));
		LiveArgs la{};
		la.recs = e->d_bound.p;
		la.os1 = stage1.d_os.p;
		la.os2 = stage2.d_os.p;
		la.nos1 = nos1;
		la.nassay = n_assays;
		la.nbucket = nbucket;
		la.shift = shift;
		la.live_r = e->d_live.p;
		la.live_f = e->d_live.p + words;
		la.out_heads = e->d_live_heads.p;
		la.out_index = e->d_live_index.p;
		la.count = e->d_live_ctl.p;
		la.err_flags = e->d_live_ctl.p + 1;
		if (padlock_format) {
			for (int pass = 0; pass < 3; ++pass) k_live_padlock<<<gen_grid, 256, 0, e->stream>>>(la, n2, pass);
		}
		else {
			if (n2 > n1) k_live_mark_plus<<<gen_grid, 256, 0, e->stream>>>(la, n1, n2);
			if (n1) k_live_compact<<<gen_grid, 256, 0, e->stream>>>(la, 0, n1, 1);
			if (n2 > n1) k_live_compact<<<gen_grid, 256, 0, e->stream>>>(la, n1, n2, 2);
		}
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches += 3;
		uint32_t ctl[2] = {0, 0};
		CUDA_OK(cudaMemcpyAsync(ctl, e->d_live_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		if (ctl[1] & F_TRUNC)
			throw std::runtime_error("more than 64 co-optimal DP cells or an alignment longer than the record (unsupported)");
		if (ctl[1] & (F_OOB | F_STACK))
			throw std::runtime_error("NucCruc traceback left the DP matrix (the reference reads unchecked ring-buffer memory here, SURVEY 8a/B4); unsupported parameters");
		n_live = ctl[0];
		e->reserve_heads(n_live);
		if (n_live) {
			CUDA_OK(cudaMemcpyAsync(e->h_heads, e->d_live_heads.p, (size_t)n_live*sizeof(BoundHead), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaMemcpyAsync(e->h_live_index, e->d_live_index.p, (size_t)n_live*sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
		}
		site_index = e->h_live_index;
		e->stats.d2h_bytes += (uint64_t)n_live*(sizeof(BoundHead) + sizeof(uint32_t));
		// PCR: the F x R (x P) join of the live sites runs on the device as well
		if (!padlock_format && !std::getenv("TNT_HOST_JOIN")) {
			pair_on_device(e, o, n_live, pair_order, pair_recs);
			paired_on_device = true;
		}
	}
	else {
		fetch_heads(e, 0, n2);
		e->stats.d2h_bytes += (uint64_t)n2*sizeof(BoundHead);
	}

	CUDA_OK(cudaEventRecord(t_end, e->stream));
	unsigned long long cells = 0;
	uint32_t dropped[2] = {0, 0};
	CUDA_OK(cudaMemcpyAsync(&cells, e->d_cells.p, sizeof(cells), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaMemcpyAsync(dropped, e->d_out_count.p + 3, sizeof(dropped), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	e->stats.dp_cells = cells;
	e->stats.nonbinding_dropped = dropped[0];
	e->stats.undefined_dropped = dropped[1];
	float ms = 0;
	CUDA_OK(cudaEventElapsedTime(&ms, t_begin, t_end));
	e->stats.total_ms = ms;

	HostTimer t_sites("stage C total");
	std::vector<BoundSite> sites;
	sites.reserve(n_live);
	for (uint32_t i = 0; i < n_live; ++i) {
		const BoundHead &b = e->h_heads[i];
		if (b.flags & F_TRUNC)
			throw std::runtime_error("more than 64 co-optimal DP cells or an alignment longer than the record (unsupported)");
		if (b.flags & (F_OOB | F_STACK))
			throw std::runtime_error("NucCruc traceback left the DP matrix (the reference reads unchecked ring-buffer memory here, SURVEY 8a/B4); unsupported parameters");
		sites.push_back(make_site(b, site_index ? site_index[i] : i, os_of(b.os)));
	}
	e->stats.bound_sites = n2;

	AssembleOptions ao;
	ao.assay_format = o.assay_format;
	ao.max_len = o.max_len;
	ao.single_primer_pcr = o.single_primer_pcr != 0;
	ao.min_max_primer_clamp = o.min_max_primer_clamp;
	std::vector<int> assay_ids, assay_has_primers, assay_has_probe;
	for (const AssayHost &a : e->assays) {
		assay_ids.push_back(a.id);
		assay_has_primers.push_back(!a.F.empty() && !a.R.empty());
		assay_has_probe.push_back(!a.P.empty());
	}
	std::vector<HitSites> refs;
	bool any_probe_only = false;
	for (size_t a = 0; a < assay_has_primers.size(); ++a) any_probe_only = any_probe_only || !assay_has_primers[a];
	if (!paired_on_device || any_probe_only) {
		HostTimer t("assemble_hits");
		assemble_hits(sites, ao, assay_ids, assay_has_primers, assay_has_probe, e->hits, refs, paired_on_device);
	}
	if (paired_on_device) {
		// index triples of k_pair_join -> hit records; merged with the hits of probe-only assays (assembled
		// above) in (fragment, assay) order
		std::vector<tnt_hit> dhits(pair_recs.size());
		std::vector<HitSites> drefs(pair_recs.size());
		for (size_t i = 0; i < pair_recs.size(); ++i) {
			const BoundSite &f = sites[pair_order[(size_t)pair_recs[i].f]], &r = sites[pair_order[(size_t)pair_recs[i].r]];
			const BoundSite *p = pair_recs[i].p >= 0 ? &sites[pair_order[(size_t)pair_recs[i].p]] : nullptr;
			make_pcr_hit(f, r, p, f.assay, assay_ids[(size_t)f.assay], sites.data(), dhits[i], drefs[i]);
		}
		if (e->hits.empty()) { e->hits.swap(dhits); refs.swap(drefs); }
		else if (!dhits.empty()) {
			std::vector<tnt_hit> merged(e->hits.size() + dhits.size());
			std::vector<HitSites> mrefs(merged.size());
			auto less_group = [](const tnt_hit &x, const tnt_hit &y) { return x.target_id != y.target_id ? x.target_id < y.target_id : x.assay_index < y.assay_index; };
			size_t i = 0, j = 0, k = 0;
			while (i < e->hits.size() || j < dhits.size()) {
				const bool take_dev = j < dhits.size() && (i >= e->hits.size() || less_group(dhits[j], e->hits[i]));
				if (take_dev) { merged[k] = dhits[j]; mrefs[k] = drefs[j]; ++j; }
				else { merged[k] = e->hits[i]; mrefs[k] = refs[i]; ++i; }
				++k;
			}
			e->hits.swap(merged);
			refs.swap(mrefs);
		}
	}

	// PCR: the join above runs over every bound site.  The reference culls its match list between
	// the binding steps and can lose a site when bound sites of one assay overlap
	// (amplicon_search.cpp:679-765, see assemble.h); groups in which that is possible are
	// searched again step by step, exactly like the reference does it, and their hits replaced.
	if (pcr_primers && !e->hits.empty() && !(e->prm.reserved & TNT_ENGINE_KEEP_CULLED_SITES)) {
		const bool replay_all = std::getenv("TNT_REPLAY_ALL") != nullptr; // verification: every group with a hit
		std::vector<GroupKey> groups;
		{
			HostTimer t_sel("groups_to_replay");
			groups = groups_to_replay(e, o, sites, refs, crowded, replay_all);
			if (HostTimer::enabled()) fprintf(stderr, "[tnt]   crowded sites %zu, groups to replay %zu\n", crowded.size(), groups.size());
		}
		e->stats.replayed_groups = groups.size();
		if (!groups.empty()) {
			std::vector<tnt_hit> rhits;
			std::vector<HitSites> rrefs;
			HostTimer t_all("replay of culled groups");
			replay_groups(e, o, ao, groups, sites, rhits, rrefs);
			// merge: replayed groups take their hits from the replay, in (fragment, assay) order
			std::vector<tnt_hit> merged;
			std::vector<HitSites> mrefs;
			merged.reserve(e->hits.size());
			mrefs.reserve(e->hits.size());
			size_t i = 0, j = 0, gi = 0;
			auto key_of = [](const tnt_hit &h) { return GroupKey{h.target_id, h.assay_index}; };
			while (i < e->hits.size() || j < rhits.size()) {
				if (i < e->hits.size()) {
					const GroupKey k = key_of(e->hits[i]);
					while (gi < groups.size() && groups[gi] < k) ++gi;
					if (gi < groups.size() && groups[gi] == k) { ++i; continue; } // replayed group: superseded
				}
				if (j >= rhits.size() || (i < e->hits.size() && key_of(e->hits[i]) < key_of(rhits[j]))) {
					merged.push_back(e->hits[i]);
					mrefs.push_back(refs[i]);
					++i;
				}
				else {
					merged.push_back(rhits[j]);
					mrefs.push_back(rrefs[j]);
					++j;
				}
			}
			e->hits.swap(merged);
			refs.swap(mrefs);
		}
	}
	e->stats.hits = e->hits.size();

	// Alignment text only for the sites that made it into a hit: gather their full records
	std::vector<uint32_t> need;
	for (const HitSites &h : refs)
		for (int s : {h.forward, h.reverse, h.probe})
			if (s >= 0) need.push_back(sites[(size_t)s].index);
	std::sort(need.begin(), need.end());
	need.erase(std::unique(need.begin(), need.end()), need.end());
	std::vector<BoundRec> recs(need.size());
	if (!need.empty()) {
		e->d_gather_idx.upload(need, e->stream);
		e->d_gather.reserve(need.size(), 0, e->stream);
		k_gather_recs<<<(unsigned)std::min<size_t>(need.size(), 1024), 128, 0, e->stream>>>(e->d_bound.p, e->d_gather_idx.p, (uint32_t)need.size(), e->d_gather.p);
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches++;
		CUDA_OK(cudaMemcpyAsync(recs.data(), e->d_gather.p, recs.size()*sizeof(BoundRec), cudaMemcpyDeviceToHost, e->stream));
		e->stats.d2h_bytes += recs.size()*sizeof(BoundRec);
		CUDA_OK(cudaStreamSynchronize(e->stream));
	}
	std::vector<uint32_t> text_off(need.size());
	{
		// text of every site (nuc_cruc_output.cpp:74-205); many hits -> a few host threads
		std::vector<std::string> text(need.size());
		const unsigned nthreads = need.size() < 2048 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		auto work = [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) text[i] = render_alignment(recs[i], os_of(recs[i].h.os)); };
		if (nthreads == 1) work(0, need.size());
		else {
			std::vector<std::thread> pool;
			const size_t per = (need.size() + nthreads - 1)/nthreads;
			for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(work, std::min(need.size(), t*per), std::min(need.size(), (t + 1)*per));
			for (std::thread &t : pool) t.join();
		}
		size_t total = e->arena.size();
		for (const std::string &t : text) total += t.size() + 1;
		e->arena.reserve(total);
		for (size_t i = 0; i < need.size(); ++i) {
			text_off[i] = (uint32_t)e->arena.size();
			e->arena.append(text[i]);
			e->arena.push_back('\0');
		}
	}
	auto off_of = [&](int s) -> uint32_t {
		if (s < 0) return 0;
		const uint32_t idx = sites[(size_t)s].index;
		return text_off[(size_t)(std::lower_bound(need.begin(), need.end(), idx) - need.begin())];
	};
	for (size_t i = 0; i < e->hits.size(); ++i) {
		e->hits[i].forward.align_off = off_of(refs[i].forward);
		e->hits[i].reverse.align_off = off_of(refs[i].reverse);
		e->hits[i].probe.align_off = off_of(refs[i].probe);
	}
}

long hit_sequence(tnt_engine *e, const tnt_hit *h, char *out, size_t cap)
{
	CUDA_OK(cudaSetDevice(e->prm.device));
	if (h->target_id >= e->targets.size()) throw std::runtime_error("bad target id");
	const Target &tg = e->targets[h->target_id];
	e->sync_targets();
	int start, stop;
	SeqMode mode;
	hit_sequence_plan(*h, e->last_opt.assay_format, start, stop, mode);
	const int n = stop - start + 1;
	if (n <= 0) throw std::runtime_error("hit with start > stop");
	// fetch the part of the fragment the renderer reads
	int lo, hi;
	hit_sequence_fetch_range(start, stop, mode, (int)tg.len, lo, hi);
	std::vector<uint8_t> codes;
	if (hi >= lo) {
		const uint32_t m = (uint32_t)(hi - lo + 1);
		e->d_extract.reserve(m, 0, e->stream);
		k_extract_codes<<<std::min<uint32_t>((m + 255)/256, 1024u), 256, 0, e->stream>>>(e->view(), h->target_id, (uint32_t)lo, m, e->d_extract.p);
		CUDA_OK(cudaGetLastError());
		codes.resize(m);
		CUDA_OK(cudaMemcpyAsync(codes.data(), e->d_extract.p, m, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
	}
	const std::string s = render_hit_sequence(start, stop, mode, (int)tg.len, lo, codes);
	if (out && cap) {
		const size_t m = std::min(cap - 1, s.size());
		std::memcpy(out, s.data(), m);
		out[m] = '\0';
	}
	return (long)s.size();
}

// ------------------------------------------------------------------------------------------
// Oligo-only structures on the device (k_oligo_jobs)
// ------------------------------------------------------------------------------------------
void ensure_homo_thermo(tnt_engine *e)
{
	if (e->d_thermo_homo.p) return;
	// local_align.dS = param_init_S + param_symmetry_S (nuc_cruc.cpp:1632): one float addition,
	// done here exactly as the reference does it per evaluation
	std::unique_ptr<Thermo> th(new Thermo(e->h_thermo));
	th->init_S = th->init_S + symmetry_S();
	e->d_thermo_homo.reserve(1, 0, e->stream);
	CUDA_OK(cudaMemcpyAsync(e->d_thermo_homo.p, th.get(), sizeof(Thermo), cudaMemcpyHostToDevice, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
}

OligoJob make_job(int kind, const std::string &query, const std::string &target, float ct)
{
	OligoJob j{};
	if (query.empty() || query.size() > (size_t)MAX_OLIGO) throw std::runtime_error("oligo longer than TNT_MAX_OLIGO_LEN (56) bases");
	const std::string &t = kind == JOB_HETERODIMER ? target : query;
	if (t.size() > (size_t)MAX_WINDOW) throw std::runtime_error("second oligo longer than 64 bases");
	j.kind = kind;
	j.qlen = (int)query.size();
	j.tlen = (int)t.size();
	for (size_t i = 0; i < query.size(); ++i) {
		const int b = base_from_ascii(query[i]);
		if (b < 0) throw std::runtime_error(":char_to_nucleic_acid: Illegal base");
		j.q[i] = (uint8_t)b;
	}
	for (size_t i = 0; i < t.size(); ++i) {
		const int b = base_from_ascii(t[i]);
		if (b < 0) throw std::runtime_error(":char_to_nucleic_acid: Illegal base");
		j.t[i] = (uint8_t)b;
	}
	if (kind != JOB_HAIRPIN) {
		if (!(ct > 0.0f)) throw std::runtime_error(":NucCruc::tm_dimer: Invalid strand_concentration");
		j.r_log_ct = r_log_ct(ct);
	}
	return j;
}

// strand(c_a, c_b), nuc_cruc.h:890-910
float duplex_ct(float a, float b) { return a > b ? a - 0.5f*b : b - 0.5f*a; }

std::vector<OligoJobResult> run_oligo_jobs(tnt_engine *e, const std::vector<OligoJob> &jobs)
{
	std::vector<OligoJobResult> res(jobs.size());
	if (jobs.empty()) return res;
	CUDA_OK(cudaSetDevice(e->prm.device));
	ensure_homo_thermo(e);
	e->d_jobs.upload(jobs, e->stream);
	e->d_job_results.reserve(jobs.size(), 0, e->stream);
	e->d_job_trace.reserve(jobs.size()*(size_t)JOB_TRACE_CELLS, 0, e->stream);
	k_oligo_jobs<<<(unsigned)((jobs.size() + 31)/32), 32, 0, e->stream>>>(e->d_jobs.p, (uint32_t)jobs.size(), e->d_thermo.p,
		e->d_thermo_homo.p, e->d_job_trace.p, e->d_job_results.p);
	CUDA_OK(cudaGetLastError());
	e->stats.kernel_launches++;
	CUDA_OK(cudaMemcpyAsync(res.data(), e->d_job_results.p, res.size()*sizeof(OligoJobResult), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	return res;
}

// Text of all hits of the last search: the fragment ranges are read back from the packed database
// with one kernel and one device-to-host copy, the strings are built on a few host threads.
void hit_sequences(tnt_engine *e)
{
	if (e->seq_ready) return;
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->sync_targets();
	const size_t n = e->hits.size();
	struct Plan { int start, stop, lo; SeqMode mode; uint32_t m; uint64_t off; };
	std::vector<Plan> plan(n);
	std::vector<ExtractItem> items;
	items.reserve(n);
	uint64_t total = 0, text_total = 0;
	for (size_t i = 0; i < n; ++i) {
		const tnt_hit &h = e->hits[i];
		if (h.target_id >= e->targets.size()) throw std::runtime_error("bad target id");
		const Target &tg = e->targets[h.target_id];
		Plan &p = plan[i];
		hit_sequence_plan(h, e->last_opt.assay_format, p.start, p.stop, p.mode);
		if (p.stop < p.start) throw std::runtime_error("hit with start > stop");
		int lo, hi;
		hit_sequence_fetch_range(p.start, p.stop, p.mode, (int)tg.len, lo, hi);
		p.lo = lo;
		p.m = hi >= lo ? (uint32_t)(hi - lo + 1) : 0u;
		p.off = total;
		if (p.m) items.push_back(ExtractItem{h.target_id, (uint32_t)lo, p.m, 0u, total});
		total += p.m;
		text_total += (uint64_t)(p.stop - p.start + 1) + 1;
	}
	std::vector<uint8_t> codes(total);
	if (total) {
		e->d_extract_items.upload(items, e->stream);
		e->d_extract.reserve(total, 0, e->stream);
		k_extract_many<<<(unsigned)std::min<size_t>(items.size(), (size_t)e->sm_count*16), 256, 0, e->stream>>>(e->view(),
			e->d_extract_items.p, (uint32_t)items.size(), e->d_extract.p);
		CUDA_OK(cudaGetLastError());
		e->stats.kernel_launches++;
		CUDA_OK(cudaMemcpyAsync(codes.data(), e->d_extract.p, total, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		e->stats.d2h_bytes += total;
	}
	e->seq_off.assign(n + 1, 0);
	for (size_t i = 0; i < n; ++i) e->seq_off[i + 1] = e->seq_off[i] + (uint64_t)(plan[i].stop - plan[i].start + 1) + 1;
	e->seq_text.assign(text_total, '\0');
	auto work = [&](size_t lo, size_t hi) {
		std::vector<uint8_t> tmp;
		for (size_t i = lo; i < hi; ++i) {
			const Plan &p = plan[i];
			tmp.assign(codes.begin() + (ptrdiff_t)p.off, codes.begin() + (ptrdiff_t)(p.off + p.m));
			const std::string s = render_hit_sequence(p.start, p.stop, p.mode, (int)e->targets[e->hits[i].target_id].len, p.lo, tmp);
			std::memcpy(&e->seq_text[e->seq_off[i]], s.data(), s.size());
		}
	};
	const unsigned nthreads = n < 4096 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
	if (nthreads == 1) work(0, n);
	else {
		std::vector<std::thread> pool;
		std::vector<std::string> errors(nthreads);
		const size_t per = (n + nthreads - 1)/nthreads;
		for (unsigned t = 0; t < nthreads; ++t)
			pool.emplace_back([&, t]() {
				try { work(std::min(n, t*per), std::min(n, (t + 1)*per)); }
				catch (const std::exception &ex) { errors[t] = ex.what(); }
			});
		for (std::thread &t : pool) t.join();
		for (const std::string &err : errors) if (!err.empty()) throw std::runtime_error(err);
	}
	e->seq_ready = true;
}

} // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
#define API_BEGIN try {
#define API_END                                                        \
	}                                                                  \
	catch (const std::exception &ex) { g_error = ex.what(); return -1; } \
	catch (const char *msg) { g_error = msg; return -1; }              \
	catch (...) { g_error = "unknown error"; return -1; }              \
	return 0;

extern "C" {

const char *tnt_last_error(void) { return g_error.c_str(); }
int tnt_abi_version(void) { return TNTB200_ABI_VERSION; }

int tnt_engine_create(const tnt_engine_params *p, tnt_engine **out)
{
	API_BEGIN
	if (!p || !out) throw std::runtime_error("null argument");
	if (p->word_size < 3 || p->word_size > 8) throw std::runtime_error(":DNAHash: Unsupported word length");
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev == 0)
		throw std::runtime_error(std::string("no usable CUDA device (the engine has no CPU fallback): ") + cudaGetErrorString(ce));
	if (p->device < 0 || p->device >= ndev) throw std::runtime_error("bad device ordinal");
	CUDA_OK(cudaSetDevice(p->device));
	cudaDeviceProp prop;
	CUDA_OK(cudaGetDeviceProperties(&prop, p->device));
	if (prop.major < 10) throw std::runtime_error("the engine is built for sm_100a (B200) only");

	std::unique_ptr<tnt_engine> e(new tnt_engine);
	e->prm = *p;
	e->sm_count = prop.multiProcessorCount;
	CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
	for (auto &ev : e->ev) CUDA_OK(cudaEventCreate(&ev));
	{
		int prio_lo = 0, prio_hi = 0;
		CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CUDA_OK(cudaStreamCreateWithPriority(&e->up_stream, cudaStreamNonBlocking, prio_hi));
	}
	e->slots.reserve(MAX_SLOTS);
	e->max_slots = MAX_SLOTS;
	if (const char *v = std::getenv("TNT_UPLOAD_SLOTS")) e->max_slots = (size_t)std::min<long>(MAX_SLOTS, std::max<long>(2, std::atol(v)));
	CUDA_OK(cudaEventCreateWithFlags(&e->emit_done, cudaEventDisableTiming));
	CUDA_OK(cudaStreamCreateWithFlags(&e->emit_stream, cudaStreamNonBlocking));
	CUDA_OK(cudaEventCreateWithFlags(&e->pads_ev, cudaEventDisableTiming));
	for (int i = 0; i < 2; ++i) {
		CUDA_OK(cudaMallocHost(&e->h_stage[i], STAGE_BYTES));
		CUDA_OK(cudaEventCreateWithFlags(&e->h_free[i], cudaEventDisableTiming));
	}
	CUDA_OK(cudaMallocHost(&e->h_total, MAX_SLOTS*sizeof(uint64_t)));
	CUDA_OK(cudaMalloc(&e->d_total, MAX_SLOTS*sizeof(uint64_t)));
	build_thermo(e->h_thermo, p->target_T, p->salt, p->dangle5 != 0, p->dangle3 != 0);
	e->h_thermo.dinkelbach = p->dinkelbach ? 1 : 0;
	e->d_thermo.reserve(1, 0, e->stream);
	CUDA_OK(cudaMemcpyAsync(e->d_thermo.p, &e->h_thermo, sizeof(Thermo), cudaMemcpyHostToDevice, e->stream));
	e->d_out_count.reserve(16, 0, e->stream);
	CUDA_OK(cudaMemsetAsync(e->d_out_count.p, 0, 16*sizeof(uint32_t), e->stream));
	e->d_cells.reserve(1, 0, e->stream);
	{
		std::vector<int32_t> p5(20);
		build_p5_table(e->h_thermo, p5.data());
		e->d_p5.upload(p5, e->stream);
	}
	e->exc_pos.reserve(16, 0, e->stream);
	e->exc_code.reserve(16, 0, e->stream);
	CUDA_OK(cudaStreamSynchronize(e->stream));
	*out = e.release();
	API_END
}

void tnt_engine_destroy(tnt_engine *e)
{
	if (!e) return;
	cudaSetDevice(e->prm.device);
	cudaDeviceSynchronize();
	delete e;
}

int tnt_engine_add_target(tnt_engine *e, const uint8_t *codes, uint32_t len, uint32_t *target_id)
{
	API_BEGIN
	if (!e || (!codes && len)) throw std::runtime_error("null argument");
	add_target(e, codes, len, target_id);
	API_END
}

int tnt_engine_add_targets(tnt_engine *e, const uint8_t *const *codes, const uint32_t *lens, uint32_t n, uint32_t *first_target_id)
{
	API_BEGIN
	if (!e || ((!codes || !lens) && n)) throw std::runtime_error("null argument");
	if (first_target_id) *first_target_id = (uint32_t)e->targets.size();
	for (uint32_t i = 0; i < n; ++i) {
		if (!codes[i] && lens[i]) throw std::runtime_error("null argument");
		add_target(e, codes[i], lens[i], nullptr);
	}
	API_END
}

int tnt_engine_add_fasta(tnt_engine *e, const char *text, size_t nbytes, uint32_t fragment_threshold, uint32_t overlap,
	const tnt_fasta_record **records, size_t *n_records, const tnt_fasta_fragment **fragments, size_t *n_fragments)
{
	API_BEGIN
	if (!e || (!text && nbytes)) throw std::runtime_error("null argument");
	add_fasta(e, text, nbytes, fragment_threshold, overlap);
	if (records) *records = e->fa_records.data();
	if (n_records) *n_records = e->fa_records.size();
	if (fragments) *fragments = e->fa_fragments.data();
	if (n_fragments) *n_fragments = e->fa_fragments.size();
	API_END
}

static_assert(sizeof(tnt_packed_target) == sizeof(Target), "tnt_packed_target mirrors tnt::Target");

int tnt_engine_export_packed(tnt_engine *e, tnt_packed_info *info, tnt_packed_target *targets, uint64_t *db2,
	uint32_t *nmask, uint64_t *exc_pos, uint8_t *exc_code)
{
	API_BEGIN
	if (!e || !info) throw std::runtime_error("null argument");
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->finish_upload();
	CUDA_OK(cudaStreamSynchronize(e->up_stream));
	CUDA_OK(cudaStreamSynchronize(e->emit_stream));
	tnt_packed_info pi{};
	pi.format = TNTB200_PACKED_FORMAT;
	pi.word_size = (uint32_t)e->prm.word_size;
	pi.n_targets = e->targets.size();
	pi.n_words = e->packed_words;
	pi.n_exceptions = e->nexc;
	pi.next_base = e->next_base;
	pi.total_bases = e->total_bases;
	*info = pi;
	if (targets && pi.n_targets) std::memcpy(targets, e->targets.data(), pi.n_targets*sizeof(Target));
	if (db2 && pi.n_words) CUDA_OK(cudaMemcpyAsync(db2, e->db2.p, pi.n_words*sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
	if (nmask && pi.n_words) CUDA_OK(cudaMemcpyAsync(nmask, e->nmask.p, pi.n_words*sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
	if (exc_pos && pi.n_exceptions) CUDA_OK(cudaMemcpyAsync(exc_pos, e->exc_pos.p, pi.n_exceptions*sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
	if (exc_code && pi.n_exceptions) CUDA_OK(cudaMemcpyAsync(exc_code, e->exc_code.p, pi.n_exceptions, cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	API_END
}

int tnt_engine_import_packed(tnt_engine *e, const tnt_packed_info *info, const tnt_packed_target *targets,
	const uint64_t *db2, const uint32_t *nmask, const uint64_t *exc_pos, const uint8_t *exc_code)
{
	API_BEGIN
	if (!e || !info) throw std::runtime_error("null argument");
	if (info->format != TNTB200_PACKED_FORMAT) throw std::runtime_error("tnt_engine_import_packed: unknown snapshot format");
	if (!e->targets.empty() || e->batch_open) throw std::runtime_error("tnt_engine_import_packed: the engine already holds fragments (call tnt_engine_clear_targets first)");
	if (info->n_targets >= (1u << 24)) throw std::runtime_error("tnt_engine_import_packed: too many fragments (limit 2^24)");
	if ((info->n_targets && !targets) || (info->n_words && (!db2 || !nmask)) || (info->n_exceptions && (!exc_pos || !exc_code)))
		throw std::runtime_error("null argument");
	// consistency of the fragment table and the non-ACGT list with the arrays (a damaged file must
	// not turn into wild reads)
	uint64_t prev_end = 0, bases = 0;
	for (uint64_t i = 0; i < info->n_targets; ++i) {
		const tnt_packed_target &t = targets[i];
		if ((t.base & 63u) || t.base < prev_end || t.base + t.len > info->next_base || t.exc_begin > t.exc_end || t.exc_end > info->n_exceptions)
			throw std::runtime_error("tnt_engine_import_packed: inconsistent fragment table");
		prev_end = t.base + t.len;
		bases += t.len;
	}
	if ((info->next_base + 31u)/32u > info->n_words || bases != info->total_bases)
		throw std::runtime_error("tnt_engine_import_packed: inconsistent sizes");
	// the sparse list is binary-searched and then walked entry by entry (load_window): it has to be
	// strictly ascending, inside the base space, with codes of the non-ACGT range; the device side
	// additionally never reads past the list, whatever the mask words say
	for (uint64_t i = 0; i < info->n_exceptions; ++i) {
		if (exc_pos[i] >= info->next_base || (i && exc_pos[i] <= exc_pos[i - 1]) || exc_code[i] < 4 || exc_code[i] > 17)
			throw std::runtime_error("tnt_engine_import_packed: inconsistent non-ACGT list");
	}
	CUDA_OK(cudaSetDevice(e->prm.device));
	CUDA_OK(cudaStreamSynchronize(e->stream)); // nothing may still read arrays that reserve() replaces
	e->db2.reserve(info->n_words + 8, 0, e->up_stream);
	e->nmask.reserve(e->db2.cap, 0, e->up_stream);
	// The fragment and tile tables go first: the copy engine serves host-to-device copies in the
	// order they were issued, so a table upload issued by the search would sit behind the whole
	// snapshot (measured: the first scan started 6 ms late).
	e->targets.resize(info->n_targets);
	if (info->n_targets) std::memcpy(e->targets.data(), targets, info->n_targets*sizeof(Target));
	e->targets_dirty = true;
	e->sync_targets(false);
	// the sparse non-ACGT list first (every chunk event then implies it), then the words in chunks of
	// 32 M bases, each followed by an event: stage 1 of a search starts on the chunks that have
	// arrived, like it does with upload batches
	if (info->n_exceptions) {
		e->exc_pos.reserve(info->n_exceptions, 0, e->up_stream);
		e->exc_code.reserve(e->exc_pos.cap, 0, e->up_stream);
		CUDA_OK(cudaMemcpyAsync(e->exc_pos.p, exc_pos, info->n_exceptions*sizeof(uint64_t), cudaMemcpyHostToDevice, e->up_stream));
		CUDA_OK(cudaMemcpyAsync(e->exc_code.p, exc_code, info->n_exceptions, cudaMemcpyHostToDevice, e->up_stream));
	}
	e->import_chunks.clear();
	const uint64_t chunk_words = STAGE_BYTES/32u; // 32 M bases
	for (uint64_t w0 = 0; w0 < info->n_words; w0 += chunk_words) {
		const uint64_t nw = std::min<uint64_t>(chunk_words, info->n_words - w0);
		{
			// the batched form, like the fragment upload: plain cudaMemcpyAsync calls were observed to
			// hold back the small host-to-device copies of a concurrent search until the whole
			// snapshot had crossed
			void *dsts[2] = {e->db2.p + w0, e->nmask.p + w0};
			void *srcs[2] = {const_cast<uint64_t *>(db2 + w0), const_cast<uint32_t *>(nmask + w0)};
			size_t sizes[2] = {nw*sizeof(uint64_t), nw*sizeof(uint32_t)};
			cudaMemcpyAttributes attr{};
			attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
			size_t attr_idx = 0, fail = 0;
			CUDA_OK(cudaMemcpyBatchAsync(dsts, srcs, sizes, 2, &attr, &attr_idx, 1, &fail, e->up_stream));
		}
		const size_t ci = e->import_chunks.size();
		if (ci >= e->import_ev_pool.size()) {
			e->import_ev_pool.emplace_back();
			CUDA_OK(cudaEventCreateWithFlags(&e->import_ev_pool.back(), cudaEventDisableTiming));
		}
		CUDA_OK(cudaEventRecord(e->import_ev_pool[ci], e->up_stream));
		e->import_chunks.push_back(tnt_engine::ImportChunk{(w0 + nw)*32u, e->import_ev_pool[ci]});
	}
	e->packed_words = info->n_words;
	e->next_base = info->next_base;
	e->nexc = info->n_exceptions;
	e->total_bases = info->total_bases;
	e->upload_settled = false;
	API_END
}

int tnt_engine_get_ingest_stats(tnt_engine *e, tnt_ingest_stats *out)
{
	API_BEGIN
	if (!e || !out) throw std::runtime_error("null argument");
	CUDA_OK(cudaSetDevice(e->prm.device));
	double ms = 0.0;
	for (size_t i = 0; i + 1 < e->fa_ev_used; i += 2) {
		float t = 0.0f;
		CUDA_OK(cudaEventSynchronize(e->fa_ev[i + 1]));
		CUDA_OK(cudaEventElapsedTime(&t, e->fa_ev[i], e->fa_ev[i + 1]));
		ms += t;
	}
	e->fa_stats.parse_ms = ms;
	*out = e->fa_stats;
	API_END
}

int tnt_engine_target_codes(tnt_engine *e, uint32_t target_id, uint32_t start, uint32_t n, uint8_t *out)
{
	API_BEGIN
	if (!e || (!out && n)) throw std::runtime_error("null argument");
	if (target_id >= e->targets.size()) throw std::runtime_error("bad target id");
	if ((uint64_t)start + n > e->targets[target_id].len) throw std::runtime_error("range outside the fragment");
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->sync_targets();
	if (n) {
		e->d_extract.reserve(n, 0, e->stream);
		k_extract_codes<<<std::min<uint32_t>((n + 255)/256, 1024u), 256, 0, e->stream>>>(e->view(), target_id, start, n, e->d_extract.p);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaMemcpyAsync(out, e->d_extract.p, n, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
	}
	API_END
}

int tnt_engine_clear_targets(tnt_engine *e)
{
	API_BEGIN
	if (!e) throw std::runtime_error("null argument");
	CUDA_OK(cudaSetDevice(e->prm.device));
	CUDA_OK(cudaStreamSynchronize(e->up_stream));
	CUDA_OK(cudaStreamSynchronize(e->emit_stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	if (e->fa_copy_stream) CUDA_OK(cudaStreamSynchronize(e->fa_copy_stream));
	e->fa_records.clear();
	e->fa_fragments.clear();
	e->import_chunks.clear();
	e->targets.clear();
	e->tiles.clear();
	e->next_base = 0;
	e->packed_words = 0;
	e->nexc = 0;
	e->batch_open = false;
	e->batch_used = 0;
	e->batches.clear();
	e->piece_dst.clear();
	e->piece_src.clear();
	e->piece_size.clear();
	e->piece_uses_mirror = false;
	e->next_emit = 0;
	e->upload_settled = false;
	e->h_mirror_batch[0] = e->h_mirror_batch[1] = ~(size_t)0;
	e->total_bases = 0;
	e->targets_dirty = true;
	e->hits.clear();
	e->seq_ready = false;
	API_END
}

int tnt_engine_set_assays(tnt_engine *e, const tnt_assay *assays, int32_t n)
{
	API_BEGIN
	if (!e || (n && !assays)) throw std::runtime_error("null argument");
	std::vector<AssayHost> v;
	for (int i = 0; i < n; ++i) {
		AssayHost a;
		a.id = assays[i].id;
		a.F = assays[i].forward ? assays[i].forward : "";
		a.R = assays[i].reverse ? assays[i].reverse : "";
		a.P = assays[i].probe ? assays[i].probe : "";
		a.fdeg = assays[i].forward_degen > 0 ? assays[i].forward_degen : 1;
		a.rdeg = assays[i].reverse_degen > 0 ? assays[i].reverse_degen : 1;
		a.pdeg = assays[i].probe_degen > 0 ? assays[i].probe_degen : 1;
		if (a.F.empty() != a.R.empty()) throw std::runtime_error("an assay needs both primers or none");
		if (a.F.empty() && a.P.empty()) throw std::runtime_error("assay without oligos");
		v.push_back(a);
	}
	e->assays.swap(v);
	e->assays_version++;
	API_END
}

int tnt_engine_search(tnt_engine *e, const tnt_search_options *opt)
{
	API_BEGIN
	if (!e || !opt) throw std::runtime_error("null argument");
	search(e, *opt);
	API_END
}

int tnt_engine_get_hits(tnt_engine *e, const tnt_hit **hits, size_t *n, const char **arena, size_t *arena_size)
{
	API_BEGIN
	if (!e) throw std::runtime_error("null argument");
	if (hits) *hits = e->hits.data();
	if (n) *n = e->hits.size();
	if (arena) *arena = e->arena.data();
	if (arena_size) *arena_size = e->arena.size();
	API_END
}

int tnt_engine_get_stats(tnt_engine *e, tnt_stats *out)
{
	API_BEGIN
	if (!e || !out) throw std::runtime_error("null argument");
	*out = e->stats;
	API_END
}

long tnt_engine_hits_near_threshold(tnt_engine *e, float tm_tol, float dg_tol, uint32_t *indices, size_t cap)
{
	try {
		if (!e || (!indices && cap)) throw std::runtime_error("null argument");
		const tnt_search_options &o = e->last_opt;
		const float T = e->prm.target_T;
		size_t n = 0;
		for (size_t i = 0; i < e->hits.size(); ++i) {
			const tnt_hit &h = e->hits[i];
			bool near = false;
			for (const tnt_bound_oligo *b : {&h.forward, &h.reverse, &h.probe}) {
				if (b->oligo == TNT_OLIGO_NONE) continue;
				// the bounds the oligo was filtered with (prepare_sets): primer bounds for the primers of a
				// PCR assay, probe bounds for everything else
				const bool primer = o.assay_format == TNT_ASSAY_PCR && b->oligo != TNT_OLIGO_P;
				const float lo_tm = primer ? o.min_primer_tm : o.min_probe_tm, hi_tm = primer ? o.max_primer_tm : o.max_probe_tm;
				const float lo_dg = primer ? o.min_primer_dg : o.min_probe_dg, hi_dg = primer ? o.max_primer_dg : o.max_probe_dg;
				const float dg = b->dH - T*b->dS;
				near |= std::fabs(b->tm - lo_tm) <= tm_tol || std::fabs(b->tm - hi_tm) <= tm_tol ||
					std::fabs(dg - lo_dg) <= dg_tol || std::fabs(dg - hi_dg) <= dg_tol;
			}
			if (near) {
				if (n < cap) indices[n] = (uint32_t)i;
				++n;
			}
		}
		return (long)n;
	}
	catch (const std::exception &ex) { g_error = ex.what(); return -1; }
	catch (...) { g_error = "unknown error"; return -1; }
}

long tnt_engine_hit_sequence(tnt_engine *e, const tnt_hit *hit, char *out, size_t cap)
{
	try {
		if (!e || !hit) throw std::runtime_error("null argument");
		return hit_sequence(e, hit, out, cap);
	}
	catch (const std::exception &ex) { g_error = ex.what(); return -1; }
	catch (...) { g_error = "unknown error"; return -1; }
}

int tnt_engine_hit_sequences(tnt_engine *e, const char **text, const uint64_t **offsets, size_t *n_hits)
{
	API_BEGIN
	if (!e) throw std::runtime_error("null argument");
	hit_sequences(e);
	if (text) *text = e->seq_text.data();
	if (offsets) *offsets = e->seq_off.data();
	if (n_hits) *n_hits = e->hits.size();
	API_END
}

long tnt_engine_seeds(tnt_engine *e, uint32_t target_id, const char *oligo, int32_t plus_strand,
	uint32_t *query_loc, uint32_t *target_loc, long cap_out)
{
	try {
		if (!e || !oligo) throw std::runtime_error("null argument");
		if (target_id >= e->targets.size()) throw std::runtime_error("bad target id");
		CUDA_OK(cudaSetDevice(e->prm.device));
		e->sync_targets();
		e->stats = tnt_stats{};
		OsSet set;
		tnt_search_options o{};
		o.max_gap = o.max_mismatch = o.max_poly_degen = 999;
		set.os.push_back(make_os(e, 0, TNT_OLIGO_P, plus_strand != 0, oligo, 1.0e-6f, 1.0f, 9999.0f, -9999.0f, 0.0f, 0, 0, o));
		finish_set(e, set);
		// tiles of this fragment only
		const ScanPlan plan = scan_plan(e, set);
		uint32_t t0 = 0, t1 = 0;
		bool found = false;
		for (uint32_t t = 0; t < plan.tiles->size(); ++t) {
			if ((*plan.tiles)[t].target != target_id) continue;
			if (!found) { t0 = t; found = true; }
			t1 = t + 1;
		}
		std::vector<Candidate> cands;
		if (t1 > t0 && set.os[0].nwords > 0) {
			uint32_t cap = (uint32_t)std::min<uint64_t>((uint64_t)e->targets[target_id].len*2 + 64, 1u << 28);
			e->d_cand_count.reserve(COUNT_STRIDE, 0, e->stream);
			e->d_cand.reserve(cap, 0, e->stream);
			CUDA_OK(cudaMemsetAsync(e->d_cand_count.p, 0, sizeof(uint32_t), e->stream));
			launch_scan(e, set, scan_args(e, set, cap), t0, t1);
			uint32_t n = 0;
			CUDA_OK(cudaMemcpyAsync(&n, e->d_cand_count.p, sizeof(n), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			if (n > cap) throw std::runtime_error("seed buffer overflow");
			cands.resize(n);
			if (n) CUDA_OK(cudaMemcpyAsync(cands.data(), e->d_cand.p, n*sizeof(Candidate), cudaMemcpyDeviceToHost, e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
		}
		// the reference list is ordered by diagonal (stable sort by q - t, bind_oligo.cpp:98)
		std::sort(cands.begin(), cands.end(), [](const Candidate &a, const Candidate &b) {
			const int da = (int)(a.target_k >> 24) - (int)a.t, db = (int)(b.target_k >> 24) - (int)b.t;
			return da < db;
		});
		for (long i = 0; i < (long)cands.size() && i < cap_out; ++i) {
			query_loc[i] = cands[i].target_k >> 24;
			target_loc[i] = cands[i].t;
		}
		return (long)cands.size();
	}
	catch (const std::exception &ex) { g_error = ex.what(); return -1; }
	catch (...) { g_error = "unknown error"; return -1; }
}

int tnt_engine_align(tnt_engine *e, uint32_t target_id, const char *oligo, int32_t plus_strand,
	float strand_conc, const uint32_t *query_loc, const uint32_t *target_loc, long n, tnt_align_result *out)
{
	API_BEGIN
	if (!e || !oligo || (n && (!query_loc || !target_loc || !out))) throw std::runtime_error("null argument");
	if (target_id >= e->targets.size()) throw std::runtime_error("bad target id");
	if (n > (1L << 26)) throw std::runtime_error("too many candidates in one call");
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->sync_targets();
	e->stats = tnt_stats{};
	CUDA_OK(cudaMemsetAsync(e->d_cells.p, 0, sizeof(unsigned long long), e->stream));
	OsSet set;
	tnt_search_options o{};
	o.max_gap = o.max_mismatch = o.max_poly_degen = 999;
	set.os.push_back(make_os(e, 0, TNT_OLIGO_P, plus_strand != 0, oligo, strand_conc, 1.0f, 9999.0f, -9999.0f, 0.0f, 0, 0, o));
	finish_set(e, set);
	std::vector<Candidate> cands((size_t)n);
	for (long i = 0; i < n; ++i) {
		if (query_loc[i] > 255) throw std::runtime_error("query_loc out of range");
		cands[i].target_k = target_id | (query_loc[i] << 24);
		cands[i].t = target_loc[i];
	}
	const uint32_t cap = (uint32_t)std::max<long>(n, 1);
	e->d_cand.upload(cands, e->stream);
	e->d_cand_count.reserve(COUNT_STRIDE, 0, e->stream);
	const uint32_t cnt = (uint32_t)n;
	CUDA_OK(cudaMemcpyAsync(e->d_cand_count.p, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, e->stream));
	e->n_bound = 0;
	if (!align_buckets(e, set, cap, 0, true)) throw std::runtime_error("internal: bucket overflow");
	std::vector<BoundRec> recs((size_t)n);
	if (n) CUDA_OK(cudaMemcpyAsync(recs.data(), e->d_bound.p, (size_t)n*sizeof(BoundRec), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	e->n_bound = 0;
	for (long i = 0; i < n; ++i) {
		const BoundRec &b = recs[(size_t)i];
		tnt_align_result &r = out[i];
		std::memset(&r, 0, sizeof(r));
		r.tm = b.h.tm; r.dH = b.h.dH; r.dS = b.h.dS; r.dG = b.dG;
		r.valid = b.valid;
		r.target_start = b.win_start;
		r.target_stop = b.win_stop;
		if (b.h.flags & (F_OOB | F_STACK | F_TRUNC)) r.valid = -1;
		if (b.valid) {
			r.anchor5 = b.h.anchor5; r.anchor3 = b.h.anchor3;
			r.num_mismatch = b.h.num_mm; r.num_gap = b.h.num_gap; r.max_poly_degen = b.poly_degen;
			r.q_first = b.fm_q; r.q_last = b.lm_q;
			r.t_first = b.lm_t; r.t_last = b.fm_t; // alignment_range_target, nuc_cruc_anchor.cpp:386-389
			r.loc_5 = b.h.loc5; r.loc_3 = b.h.loc3;
			std::strncpy(r.alignment, render_alignment(b, set.os[0]).c_str(), sizeof(r.alignment) - 1);
		}
	}
	unsigned long long cells = 0;
	CUDA_OK(cudaMemcpyAsync(&cells, e->d_cells.p, sizeof(cells), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	e->stats.dp_cells = cells;
	API_END
}

int tnt_engine_oligo_dimer(tnt_engine *e, const char *query, const char *target, float conc_a, float conc_b, tnt_align_result *out)
{
	API_BEGIN
	if (!e || !query || !out) throw std::runtime_error("null argument");
	const bool homo = !target || !*target;
	const std::string tseq = homo ? std::string(query) : std::string(target);
	if (tseq.size() > (size_t)MAX_WINDOW) throw std::runtime_error("second oligo longer than 64 bases");
	if (!(conc_a >= 0.0f) || !(conc_b >= 0.0f)) throw std::runtime_error(":strand: negative concentration");
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->stats = tnt_stats{};
	CUDA_OK(cudaMemsetAsync(e->d_cells.p, 0, sizeof(unsigned long long), e->stream));
	// strand(c_a, c_b), nuc_cruc.h:893-910
	const float ct = (conc_a > conc_b) ? conc_a - 0.5f*conc_b : conc_b - 0.5f*conc_a;
	OsSet set;
	tnt_search_options o{};
	o.max_gap = o.max_mismatch = o.max_poly_degen = 999;
	set.os.push_back(make_os(e, 0, TNT_OLIGO_P, true, query, ct, 1.0f, 9999.0f, -9999.0f, 0.0f, 0, 0, o));
	finish_set(e, set);
	set.fast_ok[0] = 0; // the generic kernel takes the explicit target
	std::vector<uint8_t> tcodes(tseq.size());
	for (size_t i = 0; i < tseq.size(); ++i) {
		const int b = base_from_ascii(tseq[i]);
		if (b < 0) throw std::runtime_error(":char_to_nucleic_acid: Illegal base");
		tcodes[i] = (uint8_t)b;
	}
	e->d_explicit.upload(tcodes, e->stream);
	if (homo && !e->d_thermo_homo.p) {
		// local_align.dS = param_init_S + param_symmetry_S (nuc_cruc.cpp:1632): one float addition,
		// done here exactly as the reference does it per evaluation
		std::unique_ptr<Thermo> th(new Thermo(e->h_thermo));
		th->init_S = th->init_S + symmetry_S();
		e->d_thermo_homo.reserve(1, 0, e->stream);
		CUDA_OK(cudaMemcpyAsync(e->d_thermo_homo.p, th.get(), sizeof(Thermo), cudaMemcpyHostToDevice, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
	}
	std::vector<Candidate> cands(1);
	cands[0].target_k = 0;
	cands[0].t = 0;
	e->d_cand.upload(cands, e->stream);
	e->d_cand_count.reserve(COUNT_STRIDE, 0, e->stream);
	const uint32_t cnt = 1;
	CUDA_OK(cudaMemcpyAsync(e->d_cand_count.p, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, e->stream));
	e->n_bound = 0;
	struct Restore {
		tnt_engine *e;
		~Restore() { e->thermo_override = nullptr; e->explicit_tgt = nullptr; e->explicit_len = 0; }
	} restore{e};
	e->thermo_override = homo ? e->d_thermo_homo.p : nullptr;
	e->explicit_tgt = e->d_explicit.p;
	e->explicit_len = (int)tcodes.size();
	if (!align_buckets(e, set, 1, 0, true)) throw std::runtime_error("internal: bucket overflow");
	BoundRec b;
	CUDA_OK(cudaMemcpyAsync(&b, e->d_bound.p, sizeof(BoundRec), cudaMemcpyDeviceToHost, e->stream));
	CUDA_OK(cudaStreamSynchronize(e->stream));
	e->n_bound = 0;
	tnt_align_result &r = *out;
	std::memset(&r, 0, sizeof(r));
	r.tm = b.h.tm; r.dH = b.h.dH; r.dS = b.h.dS; r.dG = b.dG;
	r.valid = b.valid;
	if (b.h.flags & (F_OOB | F_STACK | F_TRUNC)) r.valid = -1;
	if (b.valid) {
		r.num_mismatch = b.h.num_mm; r.num_gap = b.h.num_gap;
		r.q_first = b.fm_q; r.q_last = b.lm_q;
		r.t_first = b.lm_t; r.t_last = b.fm_t;
		std::strncpy(r.alignment, render_alignment(b, set.os[0]).c_str(), sizeof(r.alignment) - 1);
	}
	API_END
}

int tnt_debug_thermo(float T, float na, int32_t *dg, uint8_t *bbp)
{
	API_BEGIN
	std::unique_ptr<Thermo> th(new Thermo);
	build_thermo(*th, T, na, false, false);
	if (dg) std::memcpy(dg, th->dg, sizeof(th->dg));
	if (bbp) std::memcpy(bbp, th->bbp, sizeof(th->bbp));
	API_END
}

int tnt_debug_thermo_at(float T, float na, float T_eval, int32_t *dg)
{
	API_BEGIN
	if (!dg) throw std::runtime_error("null argument");
	std::unique_ptr<Thermo> th(new Thermo);
	build_thermo(*th, T, na, false, false);
	dg_at_temperature(*th, T_eval, dg);
	API_END
}

int tnt_debug_words(const char *oligo, int32_t word_size, int32_t complement, uint16_t *words)
{
	if (!oligo || !words || word_size < 2 || word_size > 8) { g_error = "bad argument"; return -1; }
	if (std::strlen(oligo) > (size_t)MAX_OLIGO) { g_error = "oligo longer than TNT_MAX_OLIGO_LEN"; return -1; }
	return build_words(oligo, word_size, complement != 0, words);
}

int tnt_debug_min_columns(float T, float na, const char *oligo, float strand_concentration, float min_tm)
{
	try {
		if (!oligo || !(strand_concentration > 0.0f)) { g_error = "bad argument"; return -1; }
		const size_t L = std::strlen(oligo);
		if (L == 0 || L > (size_t)MAX_OLIGO) { g_error = "oligo longer than TNT_MAX_OLIGO_LEN"; return -1; }
		Thermo th;
		build_thermo(th, T, na, false, false);
		OligoStrand os{};
		os.len = (int)L;
		for (size_t i = 0; i < L; ++i) {
			const int b = base_from_ascii(oligo[i]);
			if (b < 0) { g_error = ":char_to_nucleic_acid: Illegal base"; return -1; }
			os.seq[i] = (uint8_t)b;
		}
		os.r_log_ct = r_log_ct(strand_concentration);
		return lean_min_columns(th, os, min_tm);
	}
	catch (const std::exception &ex) { g_error = ex.what(); return -1; }
}

int tnt_engine_oligo_hairpin(tnt_engine *e, const char *query, tnt_align_result *out)
{
	API_BEGIN
	if (!e || !query || !out) throw std::runtime_error("null argument");
	const std::vector<OligoJobResult> r = run_oligo_jobs(e, std::vector<OligoJob>(1, make_job(JOB_HAIRPIN, query, "", 0.0f)));
	std::memset(out, 0, sizeof(*out));
	out->tm = r[0].tm; out->dH = r[0].dH; out->dS = r[0].dS;
	out->dG = r[0].dH - e->prm.target_T*r[0].dS;
	out->valid = (r[0].flags & (F_OOB | F_STACK | F_TRUNC)) ? -1 : r[0].valid;
	if (r[0].valid) {
		out->q_first = r[0].fm_q; out->t_first = r[0].fm_t;
		out->q_last = r[0].lm_q; out->t_last = r[0].lm_t;
		out->num_gap = r[0].ncols; // columns of the stem
	}
	API_END
}

int tnt_engine_assay_structures(tnt_engine *e, const tnt_search_options *opt, tnt_assay_structures *out)
{
	API_BEGIN
	if (!e || !opt || (!out && !e->assays.empty())) throw std::runtime_error("null argument");
	const float fps = opt->forward_primer_strand, rps = opt->reverse_primer_strand, ps = opt->probe_strand;
	std::vector<OligoJob> jobs;
	struct Slot { size_t assay; int field; };
	std::vector<Slot> slots;
	auto add = [&](size_t a, int field, const OligoJob &j) { jobs.push_back(j); slots.push_back(Slot{a, field}); };
	for (size_t a = 0; a < e->assays.size(); ++a) {
		const AssayHost &as = e->assays[a];
		tnt_assay_structures &o = out[a];
		for (int k = 0; k < 3; ++k) o.hairpin_tm[k] = o.homodimer_tm[k] = o.heterodimer_tm[k] = -1.0f; // hybrid_sig::init()
		if (!as.F.empty() && !as.R.empty()) {
			// tntblast_local.cpp:659-683: strand(c, c) for the oligo against itself, strand(c_f, c_r) for the pair
			add(a, 0, make_job(JOB_HAIRPIN, as.F, "", 0.0f));
			add(a, 1, make_job(JOB_HAIRPIN, as.R, "", 0.0f));
			add(a, 3, make_job(JOB_HOMODIMER, as.F, "", duplex_ct(fps, fps)));
			add(a, 4, make_job(JOB_HOMODIMER, as.R, "", duplex_ct(rps, rps)));
			add(a, 6, make_job(JOB_HETERODIMER, as.F, as.R, duplex_ct(fps, rps)));
			add(a, 7, make_job(JOB_HETERODIMER, as.F, as.F, duplex_ct(fps, rps)));
			add(a, 8, make_job(JOB_HETERODIMER, as.R, as.R, duplex_ct(fps, rps)));
		}
		if (!as.P.empty()) {
			add(a, 2, make_job(JOB_HAIRPIN, as.P, "", 0.0f));
			add(a, 5, make_job(JOB_HOMODIMER, as.P, "", duplex_ct(ps, ps)));
		}
	}
	const std::vector<OligoJobResult> r = run_oligo_jobs(e, jobs);
	for (size_t i = 0; i < r.size(); ++i) {
		if (r[i].flags & (F_OOB | F_STACK | F_TRUNC))
			throw std::runtime_error("NucCruc traceback left the DP matrix (the reference reads unchecked memory here); unsupported parameters");
		tnt_assay_structures &o = out[slots[i].assay];
		const int f = slots[i].field;
		(f < 3 ? o.hairpin_tm[f] : (f < 6 ? o.homodimer_tm[f - 3] : o.heterodimer_tm[f - 6])) = r[i].tm;
	}
	API_END
}

int tnt_engine_alu_peak(tnt_engine *e, double *tops)
{
	API_BEGIN
	if (!e || !tops) throw std::runtime_error("null argument");
	CUDA_OK(cudaSetDevice(e->prm.device));
	DevBuf<int> sink;
	sink.reserve(4, 0, e->stream);
	const int iters = 4096, grid = e->sm_count*16, seed = (int)(e->total_bases & 1023u) + 3;
	for (int mode = 0; mode < 3; ++mode) {
		double best = 0.0;
		for (int rep = 0; rep < 4; ++rep) {
			CUDA_OK(cudaEventRecord(e->ev[6], e->stream));
			if (mode == 0) k_alu_peak<0><<<grid, 256, 0, e->stream>>>(iters, seed, sink.p);
			else if (mode == 1) k_alu_peak<1><<<grid, 256, 0, e->stream>>>(iters, seed, sink.p);
			else k_alu_peak<2><<<grid, 256, 0, e->stream>>>(iters, seed, sink.p);
			CUDA_OK(cudaGetLastError());
			CUDA_OK(cudaEventRecord(e->ev[7], e->stream));
			CUDA_OK(cudaStreamSynchronize(e->stream));
			float ms = 0;
			CUDA_OK(cudaEventElapsedTime(&ms, e->ev[6], e->ev[7]));
			// instructions: mode 0 / 1 one per chain step; mode 2 counts the add and the max as two operations
			const double ops = (double)grid*256.0*(double)iters*16.0*8.0*(mode == 2 ? 2.0 : 1.0);
			if (rep > 0) best = std::max(best, ops/(ms*1e-3)/1e12);
		}
		tops[mode] = best;
	}
	API_END
}

long tnt_debug_replay_selftest(uint32_t seed, int32_t cases, long *hits)
{
	try { return replay_selftest(seed, cases, hits); }
	catch (const std::exception &ex) { g_error = ex.what(); return -1; }
}

int tnt_engine_scan_only(tnt_engine *e, const tnt_search_options *opt, uint64_t *candidates, double *ms)
{
	API_BEGIN
	if (!e || !opt) throw std::runtime_error("null argument");
	CUDA_OK(cudaSetDevice(e->prm.device));
	e->sync_targets();
	prepare_sets(e, *opt);
	OsSet &set = *e->set1;
	uint64_t total = 0;
	float t_ms = 0;
	if (!set.os.empty() && !e->tiles.empty()) {
		const size_t nos = set.os.size();
		e->d_cand_count.reserve(nos*COUNT_STRIDE, 0, e->stream);
		e->d_cand.reserve(1, 0, e->stream);
		CUDA_OK(cudaMemsetAsync(e->d_cand_count.p, 0, nos*COUNT_STRIDE*sizeof(uint32_t), e->stream));
		const ScanPlan plan = scan_plan(e, set);
		CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
		launch_scan(e, set, scan_args(e, set, 0), 0, (uint32_t)plan.tiles->size()); // capacity 0: count, do not store
		CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
		std::vector<uint32_t> counts(nos);
		CUDA_OK(cudaMemcpy2DAsync(counts.data(), sizeof(uint32_t), e->d_cand_count.p, COUNT_STRIDE*sizeof(uint32_t),
			sizeof(uint32_t), nos, cudaMemcpyDeviceToHost, e->stream));
		CUDA_OK(cudaStreamSynchronize(e->stream));
		CUDA_OK(cudaEventElapsedTime(&t_ms, e->ev[0], e->ev[1]));
		for (uint32_t c : counts) total += c;
	}
	if (candidates) *candidates = total;
	if (ms) *ms = t_ms;
	API_END
}

} // extern "C"
