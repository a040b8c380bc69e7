// sm_100a kernels of the tntb200 engine: target packing, seed scan, NucCruc alignment.
#pragma once

#include <cuda_runtime.h>
#include "tnt_types.h"
#include "align_core.cuh"

namespace tnt {

// ------------------------------------------------------------------------------------------
// Resident database layout (HBM)
//   db2  : 2 bit/base, 32 bases per uint64, base i of a word at bits [2i, 2i+1]  (code & 3 --
//          exactly what DNAHash::hash<SEQPTR> indexes, seq_hash.h:571-573)
//   nmask: 1 bit/base, 32 bases per uint32, set where the seq.h code is > 3
//   exc  : sparse, sorted by global base index: the codes behind the set mask bits
//          (seq.h values 4..17; DB_GAP/DB_UNKNOWN are dropped by the window loader)
// Fragments start at multiples of 64 bases; the tail of the last word is zero.
// ------------------------------------------------------------------------------------------
struct DbView {
	const uint64_t *db2;
	const uint32_t *nmask;
	const uint64_t *exc_pos;
	const uint8_t *exc_code;
	const Target *targets;
	uint64_t nexc;
	uint64_t nwords;   // allocated 2-bit words (read-ahead of whole tiles stays below this)
};

constexpr int PACK_THREADS = 256;
constexpr int PACK_BASES_PER_THREAD = 32;
constexpr int PACK_BASES_PER_BLOCK = PACK_THREADS*PACK_BASES_PER_THREAD;

// Pass 1: bytes -> 2-bit words + mask words, per-block count of non-ACGT bases.
// `first_word` is the index of the first uint64 of this chunk in db2 (== first uint32 in nmask).
__global__ void __launch_bounds__(PACK_THREADS) k_pack(const uint8_t *__restrict__ codes, uint32_t n,
	uint64_t *__restrict__ db2, uint32_t *__restrict__ nmask, uint64_t first_word, uint32_t *__restrict__ block_count)
{
	const uint32_t word = blockIdx.x*PACK_THREADS + threadIdx.x;
	const uint32_t base0 = word*PACK_BASES_PER_THREAD;
	uint64_t bits = 0;
	uint32_t mask = 0;
	if (base0 < n) {
		const uint32_t m = min(32u, n - base0);
		if (m == 32) {
			const uint4 *p = reinterpret_cast<const uint4 *>(codes + base0);
			const uint4 a = __ldg(p), b = __ldg(p + 1);
			const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
			for (int k = 0; k < 8; ++k) {
#pragma unroll
				for (int s = 0; s < 4; ++s) {
					const uint32_t c = (w[k] >> (8*s)) & 0xffu;
					const int i = k*4 + s;
					bits |= (uint64_t)(c & 3u) << (2*i);
					mask |= (c > 3u ? 1u : 0u) << i;
				}
			}
		}
		else {
			for (uint32_t i = 0; i < m; ++i) {
				const uint32_t c = codes[base0 + i];
				bits |= (uint64_t)(c & 3u) << (2*i);
				mask |= (c > 3u ? 1u : 0u) << i;
			}
		}
		db2[first_word + word] = bits;
		nmask[first_word + word] = mask;
	}
	// block reduction of popc(mask)
	unsigned cnt = __popc(mask);
	for (int off = 16; off; off >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, off);
	__shared__ unsigned s_cnt[PACK_THREADS/32];
	if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned t = 0;
		for (int k = 0; k < PACK_THREADS/32; ++k) t += s_cnt[k];
		block_count[blockIdx.x] = t;
	}
}

// Pass 2: exclusive scan of the per-block counts (one block; nblocks <= 64 Ki).
__global__ void __launch_bounds__(1024) k_scan_counts(uint32_t *block_count, uint32_t nblocks, uint64_t *total)
{
	__shared__ uint32_t s_part[1024];
	const uint32_t per = (nblocks + 1023)/1024;
	const uint32_t b0 = threadIdx.x*per;
	uint32_t sum = 0;
	for (uint32_t k = 0; k < per && b0 + k < nblocks; ++k) sum += block_count[b0 + k];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t run = 0;
		for (int k = 0; k < 1024; ++k) { const uint32_t v = s_part[k]; s_part[k] = run; run += v; }
		*total = run;
	}
	__syncthreads();
	uint32_t run = s_part[threadIdx.x];
	for (uint32_t k = 0; k < per && b0 + k < nblocks; ++k) {
		const uint32_t v = block_count[b0 + k];
		block_count[b0 + k] = run;
		run += v;
	}
}

// Pass 3: ordered emission of the non-ACGT codes.  `exc_base` = number of entries already stored.
__global__ void __launch_bounds__(PACK_THREADS) k_emit_exceptions(const uint8_t *__restrict__ codes, uint32_t n,
	const uint32_t *__restrict__ block_offset, uint64_t exc_base, uint64_t global_base0,
	uint64_t *__restrict__ exc_pos, uint8_t *__restrict__ exc_code)
{
	const uint32_t word = blockIdx.x*PACK_THREADS + threadIdx.x;
	const uint32_t base0 = word*PACK_BASES_PER_THREAD;
	unsigned cnt = 0;
	if (base0 < n) {
		const uint32_t m = min(32u, n - base0);
		for (uint32_t i = 0; i < m; ++i) cnt += codes[base0 + i] > 3u;
	}
	// exclusive scan over the block
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned incl = cnt;
	for (int off = 1; off < 32; off <<= 1) {
		const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= (unsigned)off) incl += v;
	}
	__shared__ unsigned s_warp[PACK_THREADS/32];
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	unsigned warp_off = 0;
	for (unsigned k = 0; k < warp; ++k) warp_off += s_warp[k];
	uint64_t dst = exc_base + block_offset[blockIdx.x] + warp_off + (incl - cnt);
	if (cnt) {
		const uint32_t m = min(32u, n - base0);
		for (uint32_t i = 0; i < m; ++i) {
			const uint32_t c = codes[base0 + i];
			if (c > 3u) {
				exc_pos[dst] = global_base0 + base0 + i;
				exc_code[dst] = (uint8_t)c;
				++dst;
			}
		}
	}
}

// Small host -> device parameter blocks travel as kernel arguments instead of copy-engine
// transfers: the copy engine serves host-to-device copies in issue order, so a 16-byte
// descriptor upload issued while fragments are streaming in waits behind a 32 MB batch.
struct PokePayload { uint32_t w[1000]; };
__global__ void k_poke(uint32_t *__restrict__ dst, uint32_t nwords, PokePayload p)
{
	for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = p.w[i];
}

// ------------------------------------------------------------------------------------------
// Base access helpers
// ------------------------------------------------------------------------------------------
// W consecutive 2-bit bases starting at global base index g, first base in the low bits.
__device__ __forceinline__ uint32_t kmer_at(const uint64_t *__restrict__ db2, uint64_t g, uint32_t kmask)
{
	const uint64_t wi = g >> 5;
	const unsigned sh = (unsigned)(g & 31u)*2u;
	uint64_t x = __ldg(db2 + wi) >> sh;
	if (sh > 48u) x |= __ldg(db2 + wi + 1) << (64u - sh);
	return (uint32_t)x & kmask;
}

// Where the 2-bit words of the database come from: HBM / L2 ...
struct GlobalWords {
	const uint64_t *__restrict__ db2;
	__device__ __forceinline__ uint64_t operator()(uint64_t wi) const { return __ldg(db2 + wi); }
};
// ... or the copy of the current tile (with a halo on both sides) a scan CTA keeps in shared memory
struct TileWords {
	const uint64_t *s;   // s[0] holds word wi0
	uint64_t wi0;
	__device__ __forceinline__ uint64_t operator()(uint64_t wi) const { return s[(uint32_t)(wi - wi0)]; }
};

template <class WORDS>
__device__ __forceinline__ uint32_t kmer_from(const WORDS &words, uint64_t g, uint32_t kmask)
{
	const uint64_t wi = g >> 5;
	const unsigned sh = (unsigned)(g & 31u)*2u;
	uint64_t x = words(wi) >> sh;
	if (sh > 48u) x |= words(wi + 1) << (64u - sh);
	return (uint32_t)x & kmask;
}

__device__ inline int exception_code(const DbView &db, const Target &tg, uint64_t g)
{
	(void)tg;
	uint64_t lo = 0, hi = db.nexc; // one sorted list for the whole database
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		const uint64_t v = __ldg(db.exc_pos + mid);
		if (v < g) lo = mid + 1;
		else hi = mid;
	}
	// a mask bit without a list entry (damaged snapshot) reads as N instead of past the array
	return lo < db.nexc ? (int)__ldg(db.exc_code + lo) : 15;
}

// Load the NucCruc target for the window [start, stop) of a fragment
// (bind_oligo.cpp:521-592 minus strand: complement + push_front; :1224-1295 plus strand).
__device__ inline int load_window(const DbView &db, const Target &tg, uint32_t start, uint32_t stop, bool plus, uint8_t *out)
{
	int n = 0;
	// the non-ACGT codes of a window are consecutive entries of the sorted exception list: one
	// binary search for the first masked base, then a walk
	uint64_t exc_at = ~0ull;
	for (uint32_t p = start; p < stop; ++p) {
		const uint64_t g = tg.base + p;
		int code = (int)((__ldg(db.db2 + (g >> 5)) >> ((g & 31u)*2u)) & 3u);
		if ((__ldg(db.nmask + (g >> 5)) >> (g & 31u)) & 1u) {
			if (exc_at == ~0ull) {
				uint64_t lo = 0, hi = db.nexc;
				while (lo < hi) {
					const uint64_t mid = (lo + hi) >> 1;
					if (__ldg(db.exc_pos + mid) < g) lo = mid + 1;
					else hi = mid;
				}
				exc_at = lo;
			}
			code = exc_at < db.nexc ? (int)__ldg(db.exc_code + exc_at) : 15; // never past the list (damaged snapshot)
			++exc_at;
			if (code > 15) continue; // DB_GAP / DB_UNKNOWN are skipped silently
		}
		int b;
		if (plus) b = code <= 4 ? code : code + 2;
		else {
			// complement: A<->T C<->G I M<->K R<->Y S V<->B W H<->D N
			const uint8_t COMP[16] = {bT, bG, bC, bA, bI, bK, bY, bS, bB, bW, bR, bD, bM, bH, bV, bN};
			b = COMP[code];
		}
		out[n++] = (uint8_t)b;
	}
	if (!plus) for (int i = 0, j = n - 1; i < j; ++i, --j) { const uint8_t x = out[i]; out[i] = out[j]; out[j] = x; }
	return n;
}

// ------------------------------------------------------------------------------------------
// Seed scan (replaces DNAHash::hash + find/find_complement + the sort/unique by diagonal of
// match_oligo_to_*_strand, seq_hash.h:524-779, bind_oligo.cpp:84-122)
// ------------------------------------------------------------------------------------------
struct WordTable {
	const uint32_t *present;   // bitmap over the 4^W little-endian k-mer keys
	const uint32_t *offset;    // [4^W + 1]
	const uint32_t *entry;     // os << 8 | word index
	uint32_t nkeys;            // 4^W
};

struct ScanTile { uint32_t target; uint32_t start; };

constexpr int COUNT_STRIDE = 32; // bucket counters live in separate L2 lines (same-line atomics serialise)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_THREADS*32;

struct ScanArgs {
	DbView db;
	WordTable wt;
	const OligoStrand *os;
	const uint16_t *os_keys;   // [nos][MAX_OLIGO] little-endian keys of the compacted word list
	const uint64_t *os_packed; // [nos][2] seed-orientation oligo, 2 bit/base; bit 127 set: words are contiguous (no degenerate letter)
	const uint32_t *assay_present; // [n_assays][nkeys/32] k-mer bitmap per assay (region scan: a region belongs to one assay)
	const ScanTile *tiles;
	uint32_t tile_begin, tile_end;
	int W;
	Candidate *cand;           // [nos][cap]
	uint32_t *cand_count;      // [nos][COUNT_STRIDE]: one 128-byte line per bucket counter
	uint32_t cap;
	uint32_t *tile_counter;    // dense scan: tiles beyond the first one of a CTA are drawn from this counter (zero at launch)
};

// One seed survives per (oligo strand, diagonal): the one with the smallest word index.  A hit
// (k, t) is therefore dropped when an earlier word k' < k of the same list matches the target
// at t - (k - k') (same q - t).  `lo` bounds how far back that test may look (0 for a whole
// fragment, the region start in a region scan).
template <class WORDS>
__device__ __forceinline__ bool first_on_diagonal_from(const WORDS &words, uint64_t tbase, uint32_t t,
	uint32_t k, uint32_t lo, const uint16_t *__restrict__ keys, uint32_t kmask, int W, uint64_t plo, uint64_t phi)
{
	if (k == 0) return true;
	if ((phi >> 63) && t >= lo + k) {
		// Oligo without degenerate letters: word kk is oligo[kk, kk+W), so "an earlier word matches
		// on this diagonal" is a run of W equal bases starting before base k when the oligo is laid
		// against the target at t - k.  One XOR of the 2-bit strings, runs by shift-and-AND.
		// Two cheap answers first.  (a) The base in front of the word: if target[t-1] pairs with oligo
		// base k-1, word k-1 matches at t-1 -- not the first.  (b) Otherwise every earlier word kk >= k-W
		// contains that mismatching base and cannot match; for k <= W those are all of them.  Only
		// k > W is left with the words 0 .. k-W-1 (three quarters of the entries never get here).
		{
			const uint64_t g1 = tbase + t - 1;
			const unsigned tb = (unsigned)(words(g1 >> 5) >> ((unsigned)(g1 & 31u)*2u)) & 3u;
			const unsigned ob = (unsigned)((k - 1 < 32u ? plo >> (2u*(k - 1)) : phi >> (2u*(k - 33u)))) & 3u;
			if (tb == ob) return false;
			if (k <= (uint32_t)W) return true;
		}
		const uint64_t g0 = tbase + t - k;
		const uint64_t wi = g0 >> 5;
		const unsigned sh = (unsigned)(g0 & 31u)*2u;
		const uint64_t w0 = words(wi), w1 = words(wi + 1), w2 = words(wi + 2);
		const uint64_t tl = sh ? ((w0 >> sh) | (w1 << (64u - sh))) : w0;
		const uint64_t th = sh ? ((w1 >> sh) | (w2 << (64u - sh))) : w1;
		const uint64_t xl = tl ^ plo, xh = th ^ (phi & ~(1ull << 63));
		const uint64_t even = 0x5555555555555555ull;
		uint64_t rl = ~(xl | (xl >> 1)) & even, rh = ~(xh | (xh >> 1)) & even; // bit 2i: base i equal
		int len = 1;
		while (2*len <= W) {
			const unsigned s2 = 2u*(unsigned)len;
			rl &= (rl >> s2) | (rh << (64u - s2));
			rh &= rh >> s2;
			len *= 2;
		}
		if (len < W) {
			const unsigned s2 = 2u*(unsigned)(W - len);
			rl &= (rl >> s2) | (rh << (64u - s2));
			rh &= rh >> s2;
		}
		// words 0 .. k-1
		const uint64_t ml = k >= 32 ? ~0ull : ((1ull << (2*k)) - 1ull);
		const uint64_t mh = k > 32 ? ((1ull << (2*(k - 32))) - 1ull) : 0ull;
		return ((rl & ml) | (rh & mh)) == 0;
	}
	for (uint32_t kk = 0; kk < k; ++kk) {
		const uint32_t back = k - kk;
		if (t < lo + back) continue;
		if (kmer_from(words, tbase + t - back, kmask) == keys[kk]) return false;
	}
	return true;
}

__device__ __forceinline__ bool first_on_diagonal(const uint64_t *__restrict__ db2, uint64_t tbase, uint32_t t,
	uint32_t k, uint32_t lo, const uint16_t *__restrict__ keys, uint32_t kmask, int W, const uint64_t *__restrict__ packed)
{
	if (k == 0) return true;
	return first_on_diagonal_from(GlobalWords{db2}, tbase, t, k, lo, keys, kmask, W, __ldg(packed), __ldg(packed + 1));
}

// Append to the bucket of `os`.  Lanes of the warp that append to the same bucket in the same step
// share one atomic (a scan with one or two oligo strands would otherwise serialise on a single
// counter).
__device__ __forceinline__ void emit_candidate(const ScanArgs &a, uint32_t os, uint32_t target, uint32_t k, uint32_t t)
{
	const unsigned active = __activemask();
	const unsigned peers = __match_any_sync(active, os);
	const unsigned lane = threadIdx.x & 31u;
	const int leader = __ffs(peers) - 1;
	uint32_t base = 0;
	if ((int)lane == leader) base = atomicAdd(a.cand_count + (size_t)os*COUNT_STRIDE, (uint32_t)__popc(peers));
	base = __shfl_sync(peers, base, leader);
	const uint32_t slot = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
	if (slot < a.cap)
		__stcs(reinterpret_cast<unsigned long long *>(a.cand + (size_t)os*a.cap + slot),
			(unsigned long long)(target | (k << 24)) | ((unsigned long long)t << 32));
}

// Candidates a warp has found but not yet appended to the buckets
struct StagedCand { uint32_t os, target_k, t; };

// Append the first n (<= 32) staged candidates of a warp: lane i takes entry i; lanes that target
// the same bucket share one atomic.
__device__ __forceinline__ void staged_flush(const ScanArgs &a, const StagedCand *cbuf, uint32_t n)
{
	const unsigned lane = threadIdx.x & 31u;
	__syncwarp();
	if (lane < n) {
		const StagedCand c = cbuf[lane];
		const unsigned active = __activemask();
		const unsigned peers = __match_any_sync(active, c.os);
		const int leader = __ffs(peers) - 1;
		uint32_t base = 0;
		if ((int)lane == leader) base = atomicAdd(a.cand_count + (size_t)c.os*COUNT_STRIDE, (uint32_t)__popc(peers));
		base = __shfl_sync(peers, base, leader);
		const uint32_t slot = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
		if (slot < a.cap) {
			// one 8-byte streaming store: the candidate is read once, by another kernel, much later
			__stcs(reinterpret_cast<unsigned long long *>(a.cand + (size_t)c.os*a.cap + slot),
				(unsigned long long)c.target_k | ((unsigned long long)c.t << 32));
		}
	}
	__syncwarp();
}

// The same append split in two, so that the round trip of the atomics overlaps the work on the next 32
// candidates: `begin` takes the first 32 staged entries into registers and issues one atomic per distinct
// bucket; `end` (called before the next `begin`, and at the end of the kernel) reads the returned bases and
// stores.  Timing experiments (tools/scan_timing.py) put the atomics at 1 of 3.9 ms per Gbp x 200 strands
// when the warp waits for them on the spot.
struct PendingAppend {
	uint32_t os, target_k, t, base, rank;
	int leader;
	bool mine;    // this lane holds an entry
	bool open;    // warp-uniform: an append is in flight
};

__device__ __forceinline__ void staged_append_end(const ScanArgs &a, PendingAppend &pa)
{
	if (!pa.open) return;
	const uint32_t base = __shfl_sync(0xffffffffu, pa.base, pa.leader);
	if (pa.mine) {
		const uint32_t slot = base + pa.rank;
		if (slot < a.cap)
			__stcs(reinterpret_cast<unsigned long long *>(a.cand + (size_t)pa.os*a.cap + slot),
				(unsigned long long)pa.target_k | ((unsigned long long)pa.t << 32));
	}
	pa.open = false;
}

__device__ __forceinline__ void staged_append_begin(const ScanArgs &a, const StagedCand *cbuf, uint32_t n, PendingAppend &pa)
{
	const unsigned lane = threadIdx.x & 31u;
	__syncwarp();
	pa.mine = lane < n;
	pa.leader = 0;
	pa.base = 0;
	pa.rank = 0;
	if (pa.mine) {
		const StagedCand c = cbuf[lane];
		pa.os = c.os; pa.target_k = c.target_k; pa.t = c.t;
		const unsigned active = __activemask();
		const unsigned peers = __match_any_sync(active, c.os);
		pa.leader = __ffs(peers) - 1;
		pa.rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
		if ((int)lane == pa.leader) pa.base = atomicAdd(a.cand_count + (size_t)c.os*COUNT_STRIDE, (uint32_t)__popc(peers));
	}
	pa.open = true;
	__syncwarp();
}

// Hit handling of the scan kernels, one table hit per lane: the table entries of the 32 hits are
// walked in lock step; survivors of the diagonal test are compacted into the warp's staging
// buffer (64 entries), and a full buffer is appended to the buckets by all 32 lanes at once (one
// atomic per distinct bucket, issued in parallel: one L2 round trip per 32 candidates instead of
// one per divergent append).
__device__ __forceinline__ void warp_process_hits(const ScanArgs &a, const Target &tg, uint32_t target, bool hit, uint32_t p,
	uint32_t key, uint32_t kmask, StagedCand *cbuf, uint32_t &cn)
{
	const unsigned lane = threadIdx.x & 31u;
	uint32_t e0 = 0, e1 = 0;
	if (hit) {
		e0 = __ldg(a.wt.offset + key);
		e1 = __ldg(a.wt.offset + key + 1);
	}
	uint32_t more = __ballot_sync(0xffffffffu, e0 < e1);
	while (more) {
		bool keep = false;
		uint32_t os = 0, k = 0;
		if (e0 < e1) {
			const uint32_t ent = __ldg(a.wt.entry + e0);
			os = ent >> 8;
			k = ent & 0xffu;
			keep = first_on_diagonal(a.db.db2, tg.base, p, k, 0u, a.os_keys + (size_t)os*MAX_OLIGO, kmask, a.W, a.os_packed + 2*(size_t)os);
			++e0;
		}
		const uint32_t kept = __ballot_sync(0xffffffffu, keep);
		if (keep) {
			StagedCand &c = cbuf[cn + (uint32_t)__popc(kept & ((1u << lane) - 1u))];
			c.os = os;
			c.target_k = target | (k << 24);
			c.t = p;
		}
		cn += (uint32_t)__popc(kept);
		if (cn >= 32) {
			staged_flush(a, cbuf, 32u);
			cn -= 32;
			if (lane < cn) cbuf[lane] = cbuf[32 + lane]; // the overhang moves to the front
			__syncwarp();
		}
		more = __ballot_sync(0xffffffffu, e0 < e1);
	}
}

// Two phases per 8192-base tile so that warps stay converged: (1) every thread tests its 32
// positions against the k-mer bitmap in shared memory and the block compacts the hit positions
// into a shared queue; (2) the queue is processed one hit per thread (CSR walk, diagonal test,
// bucket append).
__global__ void __launch_bounds__(SCAN_THREADS) k_seed_scan(ScanArgs a)
{
	extern __shared__ uint32_t s_present[];
	__shared__ uint16_t s_queue[SCAN_TILE];
	__shared__ uint32_t s_warp[SCAN_THREADS/32];
	__shared__ uint32_t s_total;
	__shared__ StagedCand s_cbuf[SCAN_THREADS/32][64];

	for (uint32_t i = threadIdx.x; i < (a.wt.nkeys + 31)/32; i += SCAN_THREADS) s_present[i] = a.wt.present[i];
	__syncthreads();

	const uint32_t kmask = a.wt.nkeys - 1;
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	StagedCand *cbuf = s_cbuf[warp];
	uint32_t cn = 0; // staged candidates of this warp (warp-uniform)

	// Tiles differ in cost (hits per tile, queue rounds), so a CTA takes its first tile by index and
	// draws the following ones from a counter: one wave of CTAs, balanced to the tile (a static
	// stride was 7 % slower, and sensitive to the grid size).
	__shared__ uint32_t s_next;
	uint32_t tile = a.tile_begin + blockIdx.x;
	while (tile < a.tile_end) {
		const ScanTile tl = a.tiles[tile];
		const Target tg = a.db.targets[tl.target];
		const uint32_t p0 = tl.start + threadIdx.x*32u;
		if (threadIdx.x == 0) s_next = a.tile_begin + gridDim.x + atomicAdd(a.tile_counter, 1u); // read after the barriers below

		// phase 1: 32 positions per thread -> bit mask of positions whose W-mer is in the table
		uint32_t hitmask = 0;
		if (p0 < tg.len) {
			const uint64_t wi = (tg.base + p0) >> 5;
			const uint64_t lo = __ldg(a.db.db2 + wi);
			const uint64_t hi = __ldg(a.db.db2 + wi + 1); // the allocation is padded
			const uint32_t nvalid = (p0 + (uint32_t)a.W <= tg.len) ? min(32u, tg.len - (uint32_t)a.W + 1u - p0) : 0u;
#pragma unroll
			for (uint32_t b = 0; b < 32; ++b) {
				const uint64_t x = b ? ((lo >> (2*b)) | (hi << (64 - 2*b))) : lo;
				const uint32_t key = (uint32_t)x & kmask;
				hitmask |= ((s_present[key >> 5] >> (key & 31u)) & 1u) << b;
			}
			if (nvalid < 32) hitmask &= (nvalid ? ((1u << nvalid) - 1u) : 0u);
		}

		// block-wide exclusive scan of the hit counts
		const uint32_t cnt = __popc(hitmask);
		uint32_t incl = cnt;
		for (int off = 1; off < 32; off <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= (unsigned)off) incl += v;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		uint32_t base = 0;
		for (unsigned w = 0; w < warp; ++w) base += s_warp[w];
		if (threadIdx.x == SCAN_THREADS - 1) s_total = base + incl;
		uint32_t slot = base + incl - cnt;
		uint32_t m = hitmask;
		while (m) {
			const uint32_t b = __ffs(m) - 1;
			m &= m - 1;
			s_queue[slot++] = (uint16_t)(threadIdx.x*32u + b);
		}
		__syncthreads();

		// phase 2: every warp takes 32 queued positions at a time, one per lane
		const uint32_t total = s_total;
		for (uint32_t q0 = warp*32u; q0 < total; q0 += SCAN_THREADS) {
			const uint32_t q = q0 + lane;
			uint32_t p = 0, key = 0;
			if (q < total) {
				p = tl.start + s_queue[q];
				key = kmer_at(a.db.db2, tg.base + p, kmask);
			}
			warp_process_hits(a, tg, tl.target, q < total, p, key, kmask, cbuf, cn);
		}
		tile = s_next;
		__syncthreads();
	}
	staged_flush(a, cbuf, cn);
}

// ------------------------------------------------------------------------------------------
// Dense scan with the whole hit path in shared memory (k_seed_scan_smem).
//
// ncu on k_seed_scan (profiles/ncu_k_seed_scan_r02_v7_raw.csv): LSU wavefronts 66 %, issue 47 %.  The hit
// path issued about ten scattered global loads per table hit (two CSR offsets, the entry, two words
// of the packed oligo, three database words for the diagonal test, the k-mer of the queued position),
// and a warp-wide load of 32 unrelated addresses costs up to 32 wavefronts where a shared-memory
// load with random banks costs three or four.  Here a CTA keeps
//   - the 2-bit words of its tile plus a halo (the diagonal test looks back at most MAX_OLIGO bases
//     and forward 64),
//   - the k-mer table in a rank-compressed form: the presence bitmap, the number of set bits in
//     front of each bitmap word, and offsets only for the keys that occur (rank = prefix + popcount),
//   - the entries and the packed oligos
// in shared memory, so that only the candidate append leaves the SM.  Chosen by the host when the
// table fits (SCAN_SMEM_TABLE_MAX bytes; ~25 KB for 100 TaqMan assays); larger tables (thousands of
// oligo strands) keep k_seed_scan.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_HALO_BEFORE = 2;   // 64 bases >= MAX_OLIGO
constexpr int SCAN_HALO_AFTER = 4;
constexpr size_t SCAN_SMEM_TABLE_MAX = 72u << 10;
constexpr int SCAN_ITEMS = 128;       // expanded (position, entry) items per warp and round

struct SmemScanArgs {
	ScanArgs s;
	const uint16_t *prefix;    // [nkeys/32] set bits of the presence bitmap in front of each word
	const uint16_t *doff;      // [distinct + 1] entry offsets of the keys that occur, in key order
	uint32_t distinct;
	uint32_t nentries;
	uint32_t nos;
};

__global__ void __launch_bounds__(SCAN_THREADS) k_seed_scan_smem(SmemScanArgs sa)
{
	const ScanArgs &a = sa.s;
	extern __shared__ __align__(16) unsigned char s_dyn_scan[];
	// layout: packed oligos (8-byte aligned) | entries | bitmap | prefix | offsets
	uint64_t *s_packed = reinterpret_cast<uint64_t *>(s_dyn_scan);
	uint32_t *s_entry = reinterpret_cast<uint32_t *>(s_packed + 2*(size_t)sa.nos);
	uint32_t *s_present = s_entry + sa.nentries;
	const uint32_t bm_words = (a.wt.nkeys + 31)/32;
	uint16_t *s_prefix = reinterpret_cast<uint16_t *>(s_present + bm_words);
	uint16_t *s_doff = s_prefix + bm_words;
	__shared__ uint64_t s_tile[SCAN_HALO_BEFORE + SCAN_THREADS + SCAN_HALO_AFTER];
	__shared__ uint16_t s_queue[SCAN_TILE];
	__shared__ uint32_t s_warp[SCAN_THREADS/32];
	__shared__ uint32_t s_total;
	__shared__ uint32_t s_next;
	__shared__ StagedCand s_cbuf[SCAN_THREADS/32][64];
	__shared__ uint32_t s_items[SCAN_THREADS/32][SCAN_ITEMS];   // entry index << 16 | position in the tile
	__shared__ uint32_t s_full[SCAN_THREADS/32][64];

	for (uint32_t i = threadIdx.x; i < 2*sa.nos; i += SCAN_THREADS) s_packed[i] = a.os_packed[i];
	for (uint32_t i = threadIdx.x; i < sa.nentries; i += SCAN_THREADS) s_entry[i] = a.wt.entry[i];
	for (uint32_t i = threadIdx.x; i < bm_words; i += SCAN_THREADS) { s_present[i] = a.wt.present[i]; s_prefix[i] = sa.prefix[i]; }
	for (uint32_t i = threadIdx.x; i <= sa.distinct; i += SCAN_THREADS) s_doff[i] = sa.doff[i];
	__syncthreads();

	const uint32_t kmask = a.wt.nkeys - 1;
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	StagedCand *cbuf = s_cbuf[warp];
	uint32_t cn = 0;
	PendingAppend pend;
	pend.open = false;
	pend.mine = false;
	pend.os = pend.target_k = pend.t = pend.base = pend.rank = 0;
	pend.leader = 0;

	uint32_t tile = a.tile_begin + blockIdx.x;
	while (tile < a.tile_end) {
		const ScanTile tl = a.tiles[tile];
		const Target tg = a.db.targets[tl.target];
		const uint32_t p0 = tl.start + threadIdx.x*32u;
		if (threadIdx.x == 0) s_next = a.tile_begin + gridDim.x + atomicAdd(a.tile_counter, 1u);

		// the tile's words (one per thread) and the halo; words outside the fragment are never looked
		// at by the tests below (positions are bounded by tg.len, look-backs by the fragment start),
		// they only have to be readable: the allocation is padded behind, and nothing lies in front of word 0
		const uint64_t wi0 = (tg.base + tl.start) >> 5;
		const uint64_t lo = wi0 + threadIdx.x < a.db.nwords ? __ldg(a.db.db2 + wi0 + threadIdx.x) : 0ull;
		s_tile[SCAN_HALO_BEFORE + threadIdx.x] = lo;
		if (threadIdx.x < SCAN_HALO_BEFORE)
			s_tile[threadIdx.x] = wi0 + threadIdx.x >= (uint64_t)SCAN_HALO_BEFORE ? __ldg(a.db.db2 + wi0 + threadIdx.x - SCAN_HALO_BEFORE) : 0ull;
		if (threadIdx.x >= SCAN_THREADS - SCAN_HALO_AFTER) {
			const uint64_t w = wi0 + SCAN_HALO_AFTER + threadIdx.x;
			s_tile[SCAN_HALO_BEFORE + SCAN_HALO_AFTER + threadIdx.x] = w < a.db.nwords ? __ldg(a.db.db2 + w) : 0ull;
		}
		__syncthreads();
		const TileWords words{s_tile, wi0 - SCAN_HALO_BEFORE};

		// phase 1: 32 positions per thread -> bit mask of positions whose W-mer is in the table
		uint32_t hitmask = 0;
		if (p0 < tg.len) {
			const uint64_t hi = s_tile[SCAN_HALO_BEFORE + threadIdx.x + 1];
			const uint32_t nvalid = (p0 + (uint32_t)a.W <= tg.len) ? min(32u, tg.len - (uint32_t)a.W + 1u - p0) : 0u;
#pragma unroll
			for (uint32_t b = 0; b < 32; ++b) {
				const uint64_t x = b ? ((lo >> (2*b)) | (hi << (64 - 2*b))) : lo;
				const uint32_t key = (uint32_t)x & kmask;
				hitmask |= ((s_present[key >> 5] >> (key & 31u)) & 1u) << b;
			}
			if (nvalid < 32) hitmask &= (nvalid ? ((1u << nvalid) - 1u) : 0u);
		}

		// block-wide exclusive scan of the hit counts
		const uint32_t cnt = __popc(hitmask);
		uint32_t incl = cnt;
		for (int off = 1; off < 32; off <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= (unsigned)off) incl += v;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		uint32_t base = 0;
		for (unsigned w = 0; w < warp; ++w) base += s_warp[w];
		if (threadIdx.x == SCAN_THREADS - 1) s_total = base + incl;
		uint32_t slot = base + incl - cnt;
		uint32_t m = hitmask;
		while (m) {
			const uint32_t b = __ffs(m) - 1;
			m &= m - 1;
			s_queue[slot++] = (uint16_t)(threadIdx.x*32u + b);
		}
		__syncthreads();

		// phase 2, dense at every step (the lock-step walk of k_seed_scan spends 185 warp instructions per
		// round whatever the number of busy lanes, and runs the full diagonal test for the whole warp
		// when one lane needs it -- ncu source page of k_seed_scan, profiles/README.md):
		//  (a) 32 queued positions per warp: key -> rank -> entry range, expanded into (position, entry)
		//      items in a per-warp buffer (most keys carry one entry, a few carry several);
		//  (b) one item per lane: the cheap answers of the diagonal test (word 0, the base in front of
		//      the word, k <= W); items that need the bit-parallel run test go to a second buffer;
		//  (c) that buffer, one item per lane.
		// Survivors are staged and appended like before.
		const uint32_t total = s_total;
		uint32_t *ibuf = s_items[warp], *fbuf = s_full[warp];
		uint32_t fn = 0; // queued full tests of this warp (warp-uniform)
		auto stage_kept = [&](bool keep, uint32_t os, uint32_t k, uint32_t p) {
			const uint32_t kept = __ballot_sync(0xffffffffu, keep);
			if (keep) {
				StagedCand &c = cbuf[cn + (uint32_t)__popc(kept & ((1u << lane) - 1u))];
				c.os = os;
				c.target_k = tl.target | (k << 24);
				c.t = p;
			}
			cn += (uint32_t)__popc(kept);
			if (cn >= 32) {
#if !defined(TNT_SCAN_EXP) || TNT_SCAN_EXP != 1
				staged_append_end(a, pend);          // the previous append: its atomics have long returned
				staged_append_begin(a, cbuf, 32u, pend);
#endif
				cn -= 32;
				if (lane < cn) cbuf[lane] = cbuf[32 + lane];
				__syncwarp();
			}
		};
		auto run_full = [&](uint32_t n) {
			// (c) n <= 32 items from the front of fbuf
			bool keep = false;
			uint32_t os = 0, k = 0, p = 0;
			if (lane < n) {
				const uint32_t it = fbuf[lane];
				const uint32_t ent = s_entry[it >> 16];
				os = ent >> 8;
				k = ent & 0xffu;
				p = tl.start + (it & 0xffffu);
				keep = first_on_diagonal_from(words, tg.base, p, k, 0u, a.os_keys + (size_t)os*MAX_OLIGO, kmask, a.W,
					s_packed[2*os], s_packed[2*os + 1]);
			}
			stage_kept(keep, os, k, p);
		};
		auto run_items = [&](uint32_t i0, uint32_t n) {
			// (b) n <= 32 items from ibuf[i0 ...]
			bool keep = false, full = false;
			uint32_t os = 0, k = 0, p = 0, it = 0;
			if (lane < n) {
				it = ibuf[i0 + lane];
				const uint32_t ent = s_entry[it >> 16];
				os = ent >> 8;
				k = ent & 0xffu;
				p = tl.start + (it & 0xffffu);
				const uint64_t phi = s_packed[2*os + 1];
				if (k == 0) keep = true;
				else if (!(phi >> 63) || p < k) full = true;  // degenerate oligo or fragment start: general test
				else {
					const uint64_t g1 = tg.base + p - 1;
					const unsigned tb = (unsigned)(words(g1 >> 5) >> ((unsigned)(g1 & 31u)*2u)) & 3u;
					const unsigned ob = (unsigned)((k - 1 < 32u ? s_packed[2*os] >> (2u*(k - 1)) : phi >> (2u*(k - 33u)))) & 3u;
					if (tb != ob) { if (k <= (uint32_t)a.W) keep = true; else full = true; }
				}
			}
			const uint32_t fm = __ballot_sync(0xffffffffu, full);
			if (full) fbuf[fn + (uint32_t)__popc(fm & ((1u << lane) - 1u))] = it;
			fn += (uint32_t)__popc(fm);
			__syncwarp();
			stage_kept(keep, os, k, p);
			if (fn >= 32) {
				run_full(32u);
				fn -= 32;
				if (lane < fn) fbuf[lane] = fbuf[32 + lane];
				__syncwarp();
			}
		};
		uint32_t in = 0; // items waiting in ibuf (warp-uniform, < 32 between rounds)
#if defined(TNT_SCAN_EXP) && TNT_SCAN_EXP == 2
		if (total == 0xffffffffu) // timing experiment: phase 1 only
#endif
		for (uint32_t q0 = warp*32u; q0 < total; q0 += SCAN_THREADS) {
			const uint32_t q = q0 + lane;
			uint32_t prel = 0, e0 = 0, e1 = 0;
			if (q < total) {
				prel = s_queue[q];
				const uint32_t key = kmer_from(words, tg.base + tl.start + prel, kmask);
				const uint32_t w = key >> 5, bit = key & 31u;
				const uint32_t rank = (uint32_t)s_prefix[w] + (uint32_t)__popc(s_present[w] & ((1u << bit) - 1u));
				e0 = s_doff[rank];
				e1 = s_doff[rank + 1];
			}
			// (a) exclusive scan of the entry counts; the items join what the previous round left over
			// (fewer than 32), so that every pass of (b) but the last has 32 busy lanes
			const uint32_t ne = e1 - e0;
			uint32_t incl2 = ne;
			for (int off = 1; off < 32; off <<= 1) {
				const uint32_t v = __shfl_up_sync(0xffffffffu, incl2, off);
				if (lane >= (unsigned)off) incl2 += v;
			}
			const uint32_t nitems = __shfl_sync(0xffffffffu, incl2, 31);
			const uint32_t first = incl2 - ne;
			for (uint32_t cb = 0; cb < nitems; cb += SCAN_ITEMS - 32) {
				for (uint32_t j = 0; j < ne; ++j) {
					const uint32_t idx = first + j;
					if (idx >= cb && idx < cb + (SCAN_ITEMS - 32)) ibuf[in + idx - cb] = ((e0 + j) << 16) | prel;
				}
				__syncwarp();
				in += min((uint32_t)(SCAN_ITEMS - 32), nitems - cb);
				uint32_t i0 = 0;
#if defined(TNT_SCAN_EXP) && TNT_SCAN_EXP == 3
				if (in > 96) { in = 0; } // timing experiment: expansion only
#else
				for (; i0 + 32 <= in; i0 += 32) run_items(i0, 32u);
#endif
				// the remainder moves to the front
				const uint32_t rest = in - i0;
				uint32_t carry = 0;
				if (lane < rest) carry = ibuf[i0 + lane];
				__syncwarp();
				if (lane < rest) ibuf[lane] = carry;
				__syncwarp();
				in = rest;
			}
		}
		if (in) run_items(0u, in);
		if (fn) run_full(fn);
		tile = s_next;
		__syncthreads();
	}
	staged_append_end(a, pend);
	staged_flush(a, cbuf, cn);
}

// ------------------------------------------------------------------------------------------
// Sparse-table variant of the seed scan (few oligo strands: most positions miss).
//
// The work per base has to shrink to a couple of instructions for the scan to approach the HBM
// roofline (0.375 B/base), so positions are tested G at a time: a second bitmap, indexed by the
// (W+G-1)-mer that covers G consecutive W-mers, says whether any of them is in the table.  Only
// groups that pass are looked at base by base.  Each thread streams 4 x 64 bases with all loads
// issued up front (coalesced 16-byte loads, the word that follows a segment comes from the
// neighbouring lane by shuffle), hit positions go through a small shared-memory queue so that the
// rare hit path runs converged (warp_process_hits: one hit per lane, staged emission).
// ------------------------------------------------------------------------------------------
constexpr int SPARSE_THREADS = 1024;
constexpr int SPARSE_SEGS = 4;                       // 64-base segments per lane: one warp covers one SCAN_TILE
constexpr int SPARSE_QUEUE = 256;                    // queued four-base groups per warp
static_assert(32*SPARSE_SEGS*64 == SCAN_TILE, "a warp scans exactly one tile");

struct SparseScanArgs {
	ScanArgs s;
	const uint32_t *group_present;   // bitmap over the 4^(W+G-1) group keys
	int G;                           // positions per group (1..4)
};

// one bit per (W+G-1)-mer: does any of its G W-mers occur in the table?
__global__ void k_build_group_bitmap(const uint32_t *__restrict__ present, uint32_t kmask, int G, uint32_t nkeys2,
	uint32_t *__restrict__ out)
{
	for (uint32_t key = blockIdx.x*blockDim.x + threadIdx.x; key < nkeys2; key += gridDim.x*blockDim.x) {
		bool any = false;
		for (int i = 0; i < G; ++i) {
			const uint32_t k = (key >> (2*i)) & kmask;
			any |= ((present[k >> 5] >> (k & 31u)) & 1u) != 0;
		}
		const uint32_t word = __ballot_sync(0xffffffffu, any);
		if ((threadIdx.x & 31) == 0) out[key >> 5] = word;
	}
}

// Warps work independently (own tiles, own queue, no block-wide barrier after the bitmaps are
// staged), so loads, filtering and the rare hit processing of different warps overlap.
__global__ void __launch_bounds__(SPARSE_THREADS) k_seed_scan_sparse(SparseScanArgs sa)
{
	extern __shared__ uint32_t s_dyn[];       // [group bitmap | W-mer bitmap | per-warp queues | per-warp staging]
	const ScanArgs &a = sa.s;
	const uint32_t kmask = a.wt.nkeys - 1;
	const int gbits = 2*(a.W + sa.G - 1);
	const uint32_t nkeys2 = 1u << gbits;
	uint32_t *s_group = s_dyn;
	uint32_t *s_present = s_dyn + nkeys2/32;
	uint32_t *s_queue_all = s_present + (a.wt.nkeys + 31)/32;
	StagedCand *s_cbuf_all = reinterpret_cast<StagedCand *>(s_queue_all + (SPARSE_THREADS/32)*SPARSE_QUEUE);
	for (uint32_t i = threadIdx.x; i < nkeys2/32; i += SPARSE_THREADS) s_group[i] = sa.group_present[i];
	for (uint32_t i = threadIdx.x; i < (a.wt.nkeys + 31)/32; i += SPARSE_THREADS) s_present[i] = a.wt.present[i];
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	__syncthreads();

	const uint32_t gmask = nkeys2 - 1;
	const int G = sa.G;
	uint32_t *queue = s_queue_all + warp*SPARSE_QUEUE;
	StagedCand *cbuf = s_cbuf_all + warp*64;
	uint32_t cn = 0; // staged candidates (warp-uniform)
	const uint32_t nwarps = gridDim.x*(SPARSE_THREADS/32);

	for (uint32_t tile = a.tile_begin + blockIdx.x*(SPARSE_THREADS/32) + warp; tile < a.tile_end; tile += nwarps) {
		const ScanTile tl = a.tiles[tile];
		const Target tg = a.db.targets[tl.target];
		const uint64_t w0 = (tg.base + tl.start) >> 5;         // first 32-base word of the tile (64-base aligned)
		const bool any_valid = tg.len >= (uint32_t)a.W;
		const uint32_t last_valid = any_valid ? tg.len - (uint32_t)a.W : 0u; // last position with a whole W-mer

		uint32_t qn = 0; // queued groups of this tile (warp-uniform)
		// all loads first: segment sgi of this lane = bases [ (sgi*32 + lane)*64, +64 ) of the tile
		uint4 seg[SPARSE_SEGS];
#pragma unroll
		for (int sgi = 0; sgi < SPARSE_SEGS; ++sgi) {
			const uint32_t segbase = tl.start + (uint32_t)(sgi*32 + lane)*64u;
			if (segbase < tg.len) seg[sgi] = __ldg(reinterpret_cast<const uint4 *>(a.db.db2 + w0) + (sgi*32 + lane));
			else seg[sgi] = make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int sgi = 0; sgi < SPARSE_SEGS; ++sgi) {
			const uint32_t segbase = tl.start + (uint32_t)(sgi*32 + lane)*64u;
			const uint64_t lo = (uint64_t)seg[sgi].x | ((uint64_t)seg[sgi].y << 32);
			const uint64_t hi = (uint64_t)seg[sgi].z | ((uint64_t)seg[sgi].w << 32);
			// the 32 bases after the segment: first word of the next lane's segment
			uint64_t nx = __shfl_down_sync(0xffffffffu, lo, 1);
			if (lane == 31) nx = (segbase < tg.len) ? __ldg(a.db.db2 + w0 + (uint64_t)(sgi*32 + lane)*2u + 2u) : 0ull;
			// branch-free pass over the 16 four-base groups: one bit per group that may hold a hit.
			// The segment is a stream of 32-bit words; a group key starts at bit 8g, i.e. inside
			// word g/4 (it spills into the next word for two of the four phases).
			const uint32_t w[5] = {seg[sgi].x, seg[sgi].y, seg[sgi].z, seg[sgi].w, (uint32_t)nx};
			uint32_t gm = 0; // G == 4: bit g <-> group g; otherwise two bits per group
			if (G == 4) {
#pragma unroll
				for (int g = 0; g < 16; ++g) {
					const int sh = (8*g) & 31;
					const uint32_t x = sh <= 12 ? (w[g >> 2] >> sh) : __funnelshift_r(w[g >> 2], w[(g >> 2) + 1], sh);
					const uint32_t gkey = x & gmask;
					const uint32_t word = s_group[gkey >> 5];
					// shift the key's bit down to bit 0 and push it into gm from the top
					gm = __funnelshift_r(gm, __funnelshift_r(word, 0u, gkey), 1);
				}
				gm >>= 16; // group 0 arrives at bit 16, group 15 at bit 31
			}
			else {
#pragma unroll
				for (int g = 0; g < 16; ++g) {
					const int bit = 8*g;
					uint64_t x;
					if (bit < 64) x = bit ? ((lo >> bit) | (hi << (64 - bit))) : lo;
					else x = (bit - 64) ? ((hi >> (bit - 64)) | (nx << (128 - bit))) : hi;
					// two sub-groups per four bases (W = 8: G = 3 -> positions 0..2 and 3; smaller G: conservative)
					const uint32_t k0 = (uint32_t)x & gmask, k1 = (uint32_t)(x >> (2*G)) & gmask;
					const uint32_t any = ((s_group[k0 >> 5] >> (k0 & 31u)) | (s_group[k1 >> 5] >> (k1 & 31u))) & 1u;
					gm |= any << g;
				}
			}
			if (segbase >= tg.len || !any_valid) gm = 0;
			// rare: queue the four-base groups that passed (phase 2 looks at the bases one by one);
			// slots come from a warp-wide ballot, the queue length lives in a register
			uint32_t pend = gm;
			while (__any_sync(0xffffffffu, pend != 0u)) {
				const bool has = pend != 0u;
				const int grp = __ffs(pend) - 1;
				pend &= pend - 1u;
				const uint32_t bal = __ballot_sync(0xffffffffu, has);
				const uint32_t slot = qn + (uint32_t)__popc(bal & ((1u << lane) - 1u));
				if (has && slot < SPARSE_QUEUE) queue[slot] = segbase + (uint32_t)(4*grp);
				qn += (uint32_t)__popc(bal);
			}
		}
		const bool overflow = qn > SPARSE_QUEUE;
		__syncwarp();
		if (overflow) {
			// more groups passed the pre-filter than the queue holds (dense table / repeats): redo
			// this tile base by base, one position per lane
			// (the trip count is the same for all lanes: warp_process_hits synchronises the whole warp)
			const uint32_t p_end = min(tl.start + (uint32_t)SCAN_TILE, tg.len);
			for (uint32_t p0 = tl.start; p0 < p_end; p0 += 32) {
				const uint32_t p = p0 + lane;
				bool hit = false;
				uint32_t key = 0;
				if (any_valid && p <= last_valid) {
					key = kmer_at(a.db.db2, tg.base + p, kmask);
					hit = ((s_present[key >> 5] >> (key & 31u)) & 1u) != 0;
				}
				warp_process_hits(a, tg, tl.target, hit, p, key, kmask, cbuf, cn);
			}
		}
		else {
			// phase 2: eight queued groups at a time, one position per lane
			const uint32_t total = qn;
			for (uint32_t base = 0; base < total; base += 8) {
				const uint32_t q = base + (lane >> 2);
				bool hit = false;
				uint32_t p = 0, key = 0;
				if (q < total) {
					p = queue[q] + (lane & 3u);
					if (p <= last_valid) {
						key = kmer_at(a.db.db2, tg.base + p, kmask);
						hit = ((s_present[key >> 5] >> (key & 31u)) & 1u) != 0;
					}
				}
				warp_process_hits(a, tg, tl.target, hit, p, key, kmask, cbuf, cn);
			}
		}
		__syncwarp();
	}
	staged_flush(a, cbuf, cn);
}

// Stage-2 scan: only the oligo strands of the region's assay, only inside the region
// (the neighbourhood of a bound minus-strand primer site where a partner / probe may sit).
struct RegionScanArgs {
	ScanArgs s;
	const Region *regions;
	uint32_t nregions;
	uint32_t whole_fragment;   // 1: the regions are pieces of whole fragments (replay): the one-seed-per-diagonal
	                           // rule looks back to the start of the fragment, not to the start of the piece
};

__global__ void __launch_bounds__(SCAN_THREADS) k_region_scan(RegionScanArgs ra)
{
	const ScanArgs &a = ra.s;
	const uint32_t kmask = a.wt.nkeys - 1;
	for (uint32_t r = blockIdx.x; r < ra.nregions; r += gridDim.x) {
		const Region rg = ra.regions[r];
		const Target tg = a.db.targets[rg.target];
		// only the words of this region's assay matter: its own bitmap rejects the rest of the table
		const uint32_t *__restrict__ present = a.assay_present ? a.assay_present + (size_t)rg.assay*(a.wt.nkeys/32u) : a.wt.present;
		for (uint32_t p = rg.start + threadIdx.x; p < rg.stop; p += SCAN_THREADS) {
			if (p + (uint32_t)a.W > tg.len) break;
			const uint32_t key = kmer_at(a.db.db2, tg.base + p, kmask);
			if (!((__ldg(present + (key >> 5)) >> (key & 31u)) & 1u)) continue;
			const uint32_t e0 = __ldg(a.wt.offset + key), e1 = __ldg(a.wt.offset + key + 1);
			for (uint32_t e = e0; e < e1; ++e) {
				const uint32_t ent = __ldg(a.wt.entry + e);
				const uint32_t os = ent >> 8, k = ent & 0xffu;
				if (a.os[os].assay != rg.assay) continue;
				if (first_on_diagonal(a.db.db2, tg.base, p, k, ra.whole_fragment ? 0u : rg.start, a.os_keys + (size_t)os*MAX_OLIGO, kmask, a.W, a.os_packed + 2*(size_t)os))
					emit_candidate(a, os, rg.target, k, p);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// NucCruc alignment of candidate windows + per-oligo filters
// (bind_oligo_to_{minus,plus}_strand, bind_oligo.cpp:456-827 / :1159-1530, one seed per thread)
// ------------------------------------------------------------------------------------------
#ifndef TNT_ALIGN_THREADS
#define TNT_ALIGN_THREADS 128
#endif
constexpr int ALIGN_THREADS = TNT_ALIGN_THREADS;

struct AlignUnit { uint32_t os; uint32_t begin; uint32_t count; }; // `begin` indexes cand[os*cap + ...]

// Work of one launch = a few groups (one per oligo strand), each cut into units of ALIGN_THREADS
// candidates.  Only the per-group arrays travel to the device; a CTA finds the group of unit u by
// binary search in the prefix of unit counts.
struct AlignGroup { uint32_t os; uint32_t first; uint32_t count; uint32_t unit_prefix; }; // first: index of the group's first candidate in cand[os*cap + ...]

// `hint` carries the group of the CTA's previous unit (units only move forward): the first call
// (hint == ~0u) searches, later calls step.
__device__ __forceinline__ AlignUnit unit_of(const AlignGroup *__restrict__ groups, uint32_t ngroups, uint32_t u, uint32_t &hint)
{
	uint32_t lo = hint;
	if (lo == 0xffffffffu) {
		lo = 0;
		uint32_t hi = ngroups; // last group with unit_prefix <= u
		while (hi - lo > 1) {
			const uint32_t mid = (lo + hi) >> 1;
			if (groups[mid].unit_prefix <= u) lo = mid;
			else hi = mid;
		}
	}
	else while (lo + 1 < ngroups && groups[lo + 1].unit_prefix <= u) ++lo;
	hint = lo;
	const AlignGroup g = groups[lo];
	const uint32_t local = (u - g.unit_prefix)*ALIGN_THREADS;
	AlignUnit r;
	r.os = g.os;
	r.begin = g.first + local;
	r.count = min((uint32_t)ALIGN_THREADS, g.count - local);
	return r;
}

// A candidate the fast kernel hands to the generic one (window with non-ACGT target bases)
struct SlowItem { uint32_t os; uint32_t slot; Candidate c; };

struct AlignArgs {
	DbView db;
	const Thermo *thermo;
	const OligoStrand *os;
	const Candidate *cand;
	uint32_t cap;
	const AlignGroup *groups;
	uint32_t ngroups;
	uint32_t nunits;
	int max_lt;                // generic kernel: row stride of the shared DP rows (columns 0..max_lt)
	uint16_t *trace;           // [gridDim.x][cells][ALIGN_THREADS]
	uint32_t trace_cells;      // generic: 16-bit cells per CTA and thread; fast tiers: 32-bit words
	BoundRec *out;             // filtered mode: appended; all mode: out[slot]
	uint32_t *out_count;
	uint32_t out_cap;
	uint32_t os_base;          // added to the oligo-strand index stored in the records
	int emit_all;              // 1: write every result at out[units[u].begin + tid] (or slot_map[...])
	const uint32_t *slot_map;  // optional, emit_all only: output slot per candidate index
	unsigned long long *cells; // sum of Lq*Lt
	// fast kernel only
	const int32_t *row_tab;    // per oligo strand: len rows x ROW_WORDS (full-trace tier)
	const int32_t *lean_tab;   // per oligo strand: len rows x LEAN_WORDS (lean tier)
	const uint32_t *row_off;   // first row of each oligo strand in both tables
	const int32_t *p5_tab;     // [20]
	SlowItem *slow;            // -> generic kernel
	uint32_t *slow_count;
	// -> full-trace fast kernel: handed over in per-oligo-strand segments (no regrouping pass)
	Candidate *retry_cand;      // segment of strand s: [retry_base[s], retry_base[s] + retry_cap[s])
	uint32_t *retry_slot;       // output slot of each handed-over candidate (emit_all)
	const uint32_t *retry_base;
	const uint32_t *retry_cap;
	uint32_t *retry_fill;       // per strand, COUNT_STRIDE apart; keeps counting past the capacity
	uint32_t slow_cap;         // capacity of both lists
	// generic kernel only: align against this sequence (NucCruc codes, 5'->3') instead of a database
	// window -- the oligo-only duplexes of tntblast_local.cpp:657-686
	const uint8_t *explicit_tgt;
	int explicit_len;
};

// Thresholds in the reference's order (bind_oligo.cpp:598-714), target coordinates
// (:721-731 minus, :1424-1434 plus) and the output record.
template <class TG>
__device__ inline void finish_alignment(const AlignArgs &a, const DpShared &sh, const OligoStrand &os, uint32_t os_index,
	uint32_t target, uint32_t k, uint32_t t, uint32_t start, uint32_t stop, const TG &tgt, int Lt,
	const Best &best, const AlnState &best_aln, unsigned flags, uint32_t all_slot)
{
	const float tm = best.tm;
	const float dG = __fsub_rn(best.dH, __fmul_rn(a.thermo->T, best.dS));
	bool pass = !(tm < os.min_tm || tm > os.max_tm);
	if (pass) pass = !(dG < os.min_dg || dG > os.max_dg);
	unsigned anchor5 = 0, anchor3 = 0, mm = 0, gaps = 0, poly = 0;
	const bool want = a.emit_all || pass;
	if (want && best.valid) {
		anchor5 = nc_anchor5(sh, tgt, Lt, best_aln);
		anchor3 = nc_anchor3(sh, tgt, Lt, best_aln);
		nc_counts(sh, best_aln, mm, gaps, poly);
	}
	else if (want) mm = (unsigned)os.len;
	// A window without any alignment has Tm = 0 and dG = 0 in the reference; bounds that accept those
	// values let it through there with whatever coordinates the thread's previous valid alignment
	// left behind (alignment::clear() keeps first_match / last_match, nuc_cruc.h:360-371).  Such a
	// window is not a binding site: it is dropped and counted (tnt_stats::nonbinding_dropped).
	if (!best.valid && pass && !a.emit_all) {
		atomicAdd(a.out_count + 3, 1u);
		pass = false;
	}
	if (pass) pass = anchor5 >= os.clamp5;
	if (pass) pass = anchor3 >= os.clamp3;
	if (pass) pass = mm <= os.max_mismatch;
	if (pass) pass = gaps <= os.max_gap;
	if (pass) pass = poly <= os.max_poly_degen;
	// The traceback left the matrix where the reference reads unchecked ring-buffer memory
	// (nuc_cruc.cpp:1497-1541, only when a terminal penalty clamps to zero): no defined answer
	// exists for this window; it is dropped and counted (tnt_stats::undefined_dropped), the rest of
	// the search is unaffected.
	if (flags & (F_OOB | F_STACK | F_TRUNC)) {
		if (!a.emit_all) atomicAdd(a.out_count + 4, 1u);
		pass = false;
	}

	if (!(a.emit_all || pass)) return;

	uint32_t slot;
	if (a.emit_all) slot = all_slot;
	else {
		slot = atomicAdd(a.out_count, 1u);
		if (slot >= a.out_cap) return;
	}
	BoundRec &r = a.out[slot];
	r.h.os = a.os_base + os_index;
	r.h.target = target;
	r.h.tm = tm; r.h.dH = best.dH; r.h.dS = best.dS;
	r.dG = dG;
	r.h.anchor5 = (int16_t)anchor5; r.h.anchor3 = (int16_t)anchor3;
	r.h.num_mm = (int16_t)mm; r.h.num_gap = (int16_t)gaps;
	r.poly_degen = (int16_t)poly;
	r.valid = best.valid ? 1 : 0;
	r.h.k = (uint8_t)k; r.h.t = t;
	r.h.flags = (uint8_t)flags;
	r.h.pad = 0;
	r.win_start = (int32_t)start; r.win_stop = (int32_t)stop;
	r.fm_q = (int16_t)best_aln.fm_q; r.fm_t = (int16_t)best_aln.fm_t;
	r.lm_q = (int16_t)best_aln.lm_q; r.lm_t = (int16_t)best_aln.lm_t;
	r.Lt = (uint8_t)Lt;
	r.pad0 = r.pad1 = 0;
	const int ncols = best.valid ? best_aln.e - best_aln.b : 0;
	r.ncols = (uint8_t)ncols;
	for (int i = 0; i < ncols; ++i) { r.cols_q[i] = best_aln.q[best_aln.b + i]; r.cols_t[i] = best_aln.t[best_aln.b + i]; }
	for (int i = 0; i < Lt; ++i) r.win[i] = (uint8_t)tgt[i];
	{
		// length of the text nuc_cruc_output.cpp:87-204 renders: unaligned prefix + columns + suffix
		const int prefix = max(0, min(best_aln.fm_q, Lt - 1 - best_aln.fm_t));
		const int suffix = max(0, min(os.len - 1 - best_aln.lm_q, best_aln.lm_t));
		r.h.align_len = (uint16_t)(best.valid ? 3*(prefix + ncols + suffix) + 17 : 0);
	}

	const int q_first = best_aln.fm_q, q_last = best_aln.lm_q, t_first = best_aln.lm_t, t_last = best_aln.fm_t;
	int t5 = (int)start, t3 = (int)start;
	if (os.plus) {
		t5 += t_first;
		t3 += t_last;
		t3 += q_first;
		t5 -= (os.len - 1) - q_last;
	}
	else {
		t5 += (int)(stop - start) - 1 - t_last;
		t3 += (int)(stop - start) - 1 - t_first;
		t5 -= q_first;
		t3 += (os.len - 1) - q_last;
	}
	r.h.loc5 = t5;
	r.h.loc3 = t3;
}

// Hand-over lists (SlowItem) -> candidate arrays grouped by oligo strand, on the device:
// histogram, exclusive scan (one block), scatter.
__global__ void k_regroup_hist(const SlowItem *__restrict__ items, uint32_t n, uint32_t *__restrict__ hist)
{
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) atomicAdd(hist + items[i].os, 1u);
}

// start[0..nos] = exclusive prefix of hist; fill[] = copy of start[0..nos)
__global__ void __launch_bounds__(1024) k_regroup_scan(const uint32_t *__restrict__ hist, uint32_t nos, uint32_t *__restrict__ start, uint32_t *__restrict__ fill)
{
	__shared__ uint32_t s_part[1024];
	const uint32_t per = (nos + 1023)/1024;
	const uint32_t b0 = threadIdx.x*per;
	uint32_t sum = 0;
	for (uint32_t k = 0; k < per && b0 + k < nos; ++k) sum += hist[b0 + k];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t run = 0;
		for (int k = 0; k < 1024; ++k) { const uint32_t v = s_part[k]; s_part[k] = run; run += v; }
		start[nos] = run;
	}
	__syncthreads();
	uint32_t run = s_part[threadIdx.x];
	for (uint32_t k = 0; k < per && b0 + k < nos; ++k) {
		const uint32_t v = hist[b0 + k];
		start[b0 + k] = run;
		fill[b0 + k] = run;
		run += v;
	}
}

__global__ void k_regroup_scatter(const SlowItem *__restrict__ items, uint32_t n, uint32_t *__restrict__ fill,
	Candidate *__restrict__ cand, uint32_t *__restrict__ slots)
{
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		const SlowItem it = items[i];
		const uint32_t d = atomicAdd(fill + it.os, 1u);
		cand[d] = it.c;
		slots[d] = it.slot;
	}
}

// ------------------------------------------------------------------------------------------
// PCR staging on the device (the reference's bind + cull sequence, amplicon_search.cpp:58-441)
// ------------------------------------------------------------------------------------------
// Stage-2 search region of every bound stage-1 (minus strand) primer site: partner primers and
// probes can only matter downstream of it, amplicon <= max_len (cull_oligo_match
// amplicon_search.cpp:679-765 uses max_len + 50 on seed positions).  One region per record (empty
// when the record carries an error flag); per-assay totals size the stage-2 candidate buckets.
__global__ void k_make_regions(const BoundRec *__restrict__ recs, uint32_t n, const OligoStrand *__restrict__ os,
	const Target *__restrict__ targets, int max_len, Region *__restrict__ out,
	unsigned long long *__restrict__ per_assay, uint32_t *__restrict__ err_flags)
{
	// slack: seed positions and loc_5 of one site lie within an oligo length (+ flanks) of each other;
	// everything that can interleave with this site in the reference's sorted list is inside the region
	const int64_t slack = 2*MAX_OLIGO + 16;
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		const BoundHead b = recs[i].h;
		Region r;
		r.target = b.target;
		r.assay = os[b.os].assay;
		r.start = r.stop = 0;
		if (b.flags & (F_OOB | F_STACK | F_TRUNC)) atomicOr(err_flags, (uint32_t)b.flags);
		else {
			const int64_t t = (int64_t)b.t, l5 = (int64_t)b.loc5;
			const int64_t lo = (t < l5 ? t : l5) - slack;
			const int64_t hi = (t > l5 ? t : l5) + (int64_t)max_len + 50 + slack;
			const int64_t len = (int64_t)targets[b.target].len;
			const int64_t start = lo > 0 ? lo : 0;
			const int64_t stop = hi < len ? (hi > 0 ? hi : 0) : len;
			if (stop > start) {
				r.start = (uint32_t)start;
				r.stop = (uint32_t)stop;
				atomicAdd(per_assay + r.assay, (unsigned long long)(stop - start));
			}
		}
		out[i] = r;
	}
}

// Which bound sites can take part in an amplicon at all?  A hit needs a minus-strand primer site
// f and a plus-strand primer site r of the same (fragment, assay) with
// f.loc_3 < r.loc_5 and r.loc_3 - f.loc_5 + 1 <= max_len (amplicon_search.cpp:383-390), and a
// probe inside [f.loc_5, r.loc_3] (:399-406).  With position buckets at least max_len wide,
// r.loc_5 (and a probe's loc_5) falls into the bucket of f.loc_5 or the next one.  Three passes
// over the site heads: plus primers mark `live_r`; minus primers that see a mark survive and
// mark `live_f`; plus primers and probes that see a live minus primer survive.  Only survivors
// (heads + record indices) leave the device.
struct LiveArgs {
	const BoundRec *recs;
	const OligoStrand *os1, *os2;
	uint32_t nos1, nassay;
	uint32_t nbucket, shift;      // buckets per fragment, log2(bucket width)
	uint32_t *live_r, *live_f;
	BoundHead *out_heads;
	uint32_t *out_index;
	uint32_t *count;
	uint32_t *err_flags;
};

__device__ __forceinline__ uint64_t live_key(const LiveArgs &a, uint32_t target, int assay, int32_t loc5)
{
	const uint32_t b = min((uint32_t)max(loc5, 0) >> a.shift, a.nbucket - 2u);
	return ((uint64_t)target*a.nassay + (uint32_t)assay)*a.nbucket + b;
}
__device__ __forceinline__ bool live_test(const uint32_t *bits, uint64_t key) { return (bits[key >> 5] >> (key & 31u)) & 1u; }

// pass 0: stage-2 records [from, to), plus-strand primers mark live_r
__global__ void k_live_mark_plus(LiveArgs a, uint32_t from, uint32_t to)
{
	for (uint32_t i = from + blockIdx.x*blockDim.x + threadIdx.x; i < to; i += gridDim.x*blockDim.x) {
		const BoundHead b = a.recs[i].h;
		if (b.flags & (F_OOB | F_STACK | F_TRUNC)) atomicOr(a.err_flags, (uint32_t)b.flags);
		const OligoStrand &o = a.os2[b.os - a.nos1];
		if (o.role == 2 /* TNT_OLIGO_P */ || !o.plus) continue;
		const uint64_t key = live_key(a, b.target, o.assay, b.loc5);
		atomicOr(a.live_r + (key >> 5), 1u << (key & 31u));
	}
}

// pass 1: stage-1 records [0, n1) (minus-strand primers); pass 2: stage-2 records [from, to)
__global__ void k_live_compact(LiveArgs a, uint32_t from, uint32_t to, int pass)
{
	for (uint32_t i = from + blockIdx.x*blockDim.x + threadIdx.x; i < to; i += gridDim.x*blockDim.x) {
		const BoundHead b = a.recs[i].h;
		bool keep;
		if (pass == 1) {
			if (b.flags & (F_OOB | F_STACK | F_TRUNC)) atomicOr(a.err_flags, (uint32_t)b.flags);
			const OligoStrand &o1 = a.os1[b.os];
			if (o1.role == 2) keep = true; // site of a probe-only assay in a PCR run (tntblast_local.cpp:612-625): a hit by itself
			else {
				const uint64_t key = live_key(a, b.target, o1.assay, b.loc5);
				keep = live_test(a.live_r, key) || live_test(a.live_r, key + 1); // bucket nbucket-1 is never marked
				if (keep) atomicOr(a.live_f + (key >> 5), 1u << (key & 31u));
			}
		}
		else {
			const uint64_t key = live_key(a, b.target, a.os2[b.os - a.nos1].assay, b.loc5);
			// a mark in the bucket before belongs to this group only if this is not the group's first bucket
			const bool first = (key % a.nbucket) == 0;
			keep = live_test(a.live_f, key) || (!first && live_test(a.live_f, key - 1));
		}
		if (!keep) continue;
		const uint32_t slot = atomicAdd(a.count, 1u);
		a.out_heads[slot] = b;
		a.out_index[slot] = i;
	}
}

// The same idea for padlock / MIPS assays (padlock_search.cpp:62-361): a ligation site needs an
// upstream probe site u (role R) and a downstream probe site d (role F) of the same (fragment,
// assay) on the same strand, 0 <= gap <= max_len bases apart.  All sites come from stage 1.  Keys
// carry the strand; with buckets at least max_len + 128 wide the partner's loc_5 lies in the same
// bucket or a neighbouring one.  pass 0: d marks live_r; pass 1: u with a mark nearby survives and
// marks live_f; pass 2: d with a live u nearby survives.
__global__ void k_live_padlock(LiveArgs a, uint32_t n, int pass)
{
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		const BoundHead b = a.recs[i].h;
		const OligoStrand &o = a.os1[b.os];
		if (pass == 0 && (b.flags & (F_OOB | F_STACK | F_TRUNC))) atomicOr(a.err_flags, (uint32_t)b.flags);
		const bool downstream = o.role == 0; /* TNT_OLIGO_F */
		if ((pass == 1) == downstream) continue;
		const uint32_t bk = min((uint32_t)max(b.loc5, 0) >> a.shift, a.nbucket - 3u) + 1u; // buckets 0 and nbucket-1 stay empty
		const uint64_t key = (((uint64_t)b.target*a.nassay + (uint32_t)o.assay)*2u + (o.plus ? 1u : 0u))*a.nbucket + bk;
		if (pass == 0) { atomicOr(a.live_r + (key >> 5), 1u << (key & 31u)); continue; }
		const uint32_t *bits = pass == 1 ? a.live_r : a.live_f;
		if (!(live_test(bits, key - 1) || live_test(bits, key) || live_test(bits, key + 1))) continue;
		if (pass == 1) atomicOr(a.live_f + (key >> 5), 1u << (key & 31u));
		const uint32_t slot = atomicAdd(a.count, 1u);
		a.out_heads[slot] = b;
		a.out_index[slot] = i;
	}
}

// ------------------------------------------------------------------------------------------
// Which bound sites have another bound site of their (fragment, assay) group close by?
//
// The reference's cull_oligo_match (amplicon_search.cpp:679-765) sorts a list that mixes bound sites
// (ordered by loc_5 / loc_3) with not-yet-bound seeds (ordered by seed position) and stops its partner
// scan on an unsigned difference of seed positions (:709).  That is only the pure optimisation it is
// meant to be while the order of the bound sites by loc_5 agrees with their order by seed position;
// the two can disagree only for sites less than an oligo length apart (a seed lies inside its
// site).  When they do, the scan breaks early and a real site is culled.  The host decides that
// pair by pair (engine.cu: groups_to_replay); this pair of kernels hands it the few sites that have
// a neighbour at all: every bound site sets a bit for (fragment, assay, loc_5 >> CROWD_SHIFT) in a
// hashed bitmap (a second bitmap records buckets hit twice); a site whose own bucket was hit twice
// or whose neighbouring bucket is occupied is reported.  Hash collisions only add reports.
// ------------------------------------------------------------------------------------------
constexpr int CROWD_SHIFT = 7;                      // buckets of 128 bases >= CROWD_REACH
constexpr int CROWD_REACH = 2*MAX_OLIGO + 16;       // farthest two sites can be with their orders disagreeing
static_assert((1 << CROWD_SHIFT) >= CROWD_REACH, "sites within reach share a bucket or sit in adjacent ones");

struct CrowdRec { uint32_t target, assay, os; int32_t loc5, loc3; uint32_t t, rec, pad; };

struct CrowdArgs {
	const BoundRec *recs;
	uint32_t n;
	const OligoStrand *os1, *os2;
	uint32_t nos1;
	uint32_t *bits;               // [2][2^log2_bits bits]: occupied | hit twice; zero before pass 0
	uint32_t log2_bits;
	CrowdRec *out;
	uint32_t *out_count;
	uint32_t out_cap;
};

__device__ __forceinline__ uint64_t crowd_hash(uint32_t target, uint32_t assay, uint32_t bucket)
{
	uint64_t x = ((uint64_t)target << 40) ^ ((uint64_t)assay << 20) ^ (uint64_t)bucket;
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
	x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
	x ^= x >> 33;
	return x;
}

__global__ void k_crowd(CrowdArgs a, int pass)
{
	const uint64_t mask = ((uint64_t)1 << a.log2_bits) - 1u;
	uint32_t *occupied = a.bits, *twice = a.bits + ((size_t)1 << (a.log2_bits - 5));
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < a.n; i += gridDim.x*blockDim.x) {
		const BoundHead b = a.recs[i].h;
		const OligoStrand &o = b.os < a.nos1 ? a.os1[b.os] : a.os2[b.os - a.nos1];
		const uint32_t bucket = ((uint32_t)max(b.loc5, 0) >> CROWD_SHIFT) + 1u;
		const uint64_t h = crowd_hash(b.target, (uint32_t)o.assay, bucket) & mask;
		const uint32_t bit = 1u << (h & 31u);
		if (pass == 0) {
			if (atomicOr(occupied + (h >> 5), bit) & bit) atomicOr(twice + (h >> 5), bit);
			continue;
		}
		const uint64_t h0 = crowd_hash(b.target, (uint32_t)o.assay, bucket - 1u) & mask;
		const uint64_t h1 = crowd_hash(b.target, (uint32_t)o.assay, bucket + 1u) & mask;
		if (!(((twice[h >> 5] >> (h & 31u)) | (occupied[h0 >> 5] >> (h0 & 31u)) | (occupied[h1 >> 5] >> (h1 & 31u))) & 1u)) continue;
		const uint32_t slot = atomicAdd(a.out_count, 1u);
		if (slot < a.out_cap) a.out[slot] = CrowdRec{b.target, (uint32_t)o.assay, b.os, b.loc5, b.loc3, b.t, i, 0u};
	}
}

// Seeds of the replayed groups -> one dense array (oligo strand, word index, position), bucket by bucket
struct CandSpan { uint32_t os, count; uint64_t out_off; };
struct ReplaySeedRec { uint32_t os, target_k, t; };

__global__ void k_compact_cands(const Candidate *__restrict__ cand, uint32_t cap, const CandSpan *__restrict__ spans, uint32_t nspans,
	ReplaySeedRec *__restrict__ out)
{
	for (uint32_t s = blockIdx.x; s < nspans; s += gridDim.x) {
		const CandSpan sp = spans[s];
		for (uint32_t i = threadIdx.x; i < sp.count; i += blockDim.x) {
			const Candidate c = cand[(size_t)sp.os*cap + i];
			ReplaySeedRec r;
			r.os = sp.os;
			r.target_k = c.target_k;
			r.t = c.t;
			out[sp.out_off + i] = r;
		}
	}
}

// ------------------------------------------------------------------------------------------
// Amplicon pairing on the device (the final loops of amplicon(), amplicon_search.cpp:355-441)
//
// The bound sites that can be part of an amplicon (k_live_*) are put into the order of the
// reference's final list -- by (fragment, assay), then (loc_5, loc_3) like sort_by_oligo_loc sorts a
// list of bound sites (:12-26), sites of equal range in the order of their categories -- by two radix
// sorts over 64-bit keys (k_pair_keys).  k_pair_unique keeps one site per category and range, the one
// bind_oligo's own uniqueness step keeps (bind_oligo.cpp:49-82, :810-826: highest Tm, then most
// mismatches, then the longest alignment text, then the later seed).  k_pair_join runs the F x R (x P)
// loops of one forward site per thread: orientation, f.loc_3 < r.loc_5, amplicon length <= max_len,
// the single-primer rule, min-max 3' clamp (:383-397), and for assays with a probe every probe site
// strictly between the two in the list, inside the amplicon and clear of the primer that binds
// its own strand (:399-441).  What leaves the device are index triples.
// ------------------------------------------------------------------------------------------
struct PairArgs {
	const BoundHead *heads;        // live sites (compacted by k_live_compact)
	uint32_t n;
	const OligoStrand *os1, *os2;
	uint32_t nos1;
	const uint8_t *assay_has_probe;
	uint64_t *key_group, *key_loc; // [n]
	uint32_t *order;               // [n] site index, becomes the sorted order
	uint8_t *alive;                // [n] by sorted position
	PairRec *pairs;
	uint32_t *pair_count;
	uint32_t pair_cap;
	int32_t max_len, single_primer_pcr, min_max_primer_clamp;
};

__device__ __forceinline__ const OligoStrand &pair_os(const PairArgs &a, uint32_t os) { return os < a.nos1 ? a.os1[os] : a.os2[os - a.nos1]; }
__device__ __forceinline__ uint32_t pair_category(const OligoStrand &o) { return (uint32_t)(o.plus ? 3 : 0) + (uint32_t)o.role; }

__global__ void k_pair_keys(PairArgs a)
{
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < a.n; i += gridDim.x*blockDim.x) {
		const BoundHead b = a.heads[i];
		const OligoStrand &o = pair_os(a, b.os);
		a.key_group[i] = ((uint64_t)b.target << 32) | (uint32_t)o.assay;
		const uint32_t span = (uint32_t)min(max(b.loc3 - b.loc5 + 32768, 0), 0xffffff);
		a.key_loc[i] = ((uint64_t)((uint32_t)b.loc5 ^ 0x80000000u) << 32) | ((uint64_t)span << 8) | pair_category(o);
		a.order[i] = i;
	}
}

// key_out[j] = key_in[order[j]]
__global__ void k_pair_gather(const uint64_t *__restrict__ key_in, const uint32_t *__restrict__ order, uint32_t n, uint64_t *__restrict__ key_out)
{
	for (uint32_t j = blockIdx.x*blockDim.x + threadIdx.x; j < n; j += gridDim.x*blockDim.x) key_out[j] = key_in[order[j]];
}

__device__ __forceinline__ bool pair_same_site(const PairArgs &a, const BoundHead &x, const BoundHead &y)
{
	if (x.target != y.target || x.loc5 != y.loc5 || x.loc3 != y.loc3) return false;
	const OligoStrand &ox = pair_os(a, x.os), &oy = pair_os(a, y.os);
	return ox.assay == oy.assay && pair_category(ox) == pair_category(oy);
}

// true: x is kept in preference to y (sort_by_bound_match, bind_oligo.cpp:49-82; equal keys: the later seed)
__device__ __forceinline__ bool pair_better(const BoundHead &x, const BoundHead &y)
{
	if (x.tm != y.tm) return x.tm > y.tm;
	if (x.num_mm != y.num_mm) return x.num_mm > y.num_mm;
	if (x.align_len != y.align_len) return x.align_len > y.align_len;
	return x.t > y.t;
}

__global__ void k_pair_unique(PairArgs a)
{
	for (uint32_t j = blockIdx.x*blockDim.x + threadIdx.x; j < a.n; j += gridDim.x*blockDim.x) {
		const BoundHead me = a.heads[a.order[j]];
		if (j > 0 && pair_same_site(a, me, a.heads[a.order[j - 1]])) continue; // not the first of its run
		uint32_t best = j;
		BoundHead bh = me;
		uint32_t k = j + 1;
		for (; k < a.n; ++k) {
			const BoundHead x = a.heads[a.order[k]];
			if (!pair_same_site(a, me, x)) break;
			if (pair_better(x, bh)) { best = k; bh = x; }
		}
		for (uint32_t m = j; m < k; ++m) a.alive[m] = m == best ? 1 : 0;
	}
}

__global__ void k_pair_join(PairArgs a)
{
	const bool apply_mmc = a.min_max_primer_clamp >= 0;
	for (uint32_t fi = blockIdx.x*blockDim.x + threadIdx.x; fi < a.n; fi += gridDim.x*blockDim.x) {
		if (!a.alive[fi]) continue;
		const BoundHead f = a.heads[a.order[fi]];
		const OligoStrand &of = pair_os(a, f.os);
		if (of.plus || of.role == 2) continue;
		const bool has_probe = a.assay_has_probe[of.assay] != 0;
		for (uint32_t ri = fi + 1; ri < a.n; ++ri) {
			const BoundHead r = a.heads[a.order[ri]];
			if (r.target != f.target) break;
			const OligoStrand &orr = pair_os(a, r.os);
			if (orr.assay != of.assay) break;
			if ((int64_t)r.loc5 > (int64_t)f.loc5 + a.max_len) break; // sorted by loc_5: no later site can close an amplicon
			if (!a.alive[ri] || !orr.plus || orr.role == 2) continue;
			if (!a.single_primer_pcr && of.role == orr.role) continue;
			if (f.loc3 >= r.loc5) continue;
			if (r.loc3 - f.loc5 + 1 > a.max_len) continue;
			if (apply_mmc && max((int)f.anchor3, (int)r.anchor3) <= a.min_max_primer_clamp) continue;
			if (!has_probe) {
				const uint32_t slot = atomicAdd(a.pair_count, 1u);
				if (slot < a.pair_cap) a.pairs[slot] = PairRec{(int32_t)fi, (int32_t)ri, -1};
				continue;
			}
			for (uint32_t pi = fi + 1; pi < ri; ++pi) {
				if (!a.alive[pi]) continue;
				const BoundHead p = a.heads[a.order[pi]];
				const OligoStrand &op = pair_os(a, p.os);
				if (op.role != 2) continue;
				if (!(p.loc5 >= f.loc5 && p.loc3 <= r.loc3)) continue;
				if (op.plus == of.plus) { if (p.loc5 <= f.loc3) continue; }
				else if (p.loc3 >= r.loc5) continue;
				const uint32_t slot = atomicAdd(a.pair_count, 1u);
				if (slot < a.pair_cap) a.pairs[slot] = PairRec{(int32_t)fi, (int32_t)ri, (int32_t)pi};
			}
		}
	}
}

// Copy selected records into a dense array (the hits' oligo sites, for text rendering)
__global__ void k_gather_recs(const BoundRec *__restrict__ src, const uint32_t *__restrict__ index, uint32_t n, BoundRec *__restrict__ dst)
{
	const uint32_t words = sizeof(BoundRec)/4;
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
		const uint32_t *s = reinterpret_cast<const uint32_t *>(src + index[i]);
		uint32_t *d = reinterpret_cast<uint32_t *>(dst + i);
		for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) d[w] = s[w];
	}
}

constexpr int DINKELBACH_MAX_ITER = 256; // the reference loop has no bound; q rises strictly, a handful of passes in practice

// Generic kernel: any IUPAC / inosine content, DP rows in shared memory.
__global__ void __launch_bounds__(ALIGN_THREADS) k_align(AlignArgs a)
{
	extern __shared__ __align__(16) unsigned char s_raw[];
	int32_t *s_dg = reinterpret_cast<int32_t *>(s_raw);
	uint8_t *s_bbp = reinterpret_cast<uint8_t *>(s_dg + TABLE);
	uint8_t *s_wc = s_bbp + NB*NB;
	uint8_t *s_q = s_wc + 52;
	int32_t *s_rows = reinterpret_cast<int32_t *>(s_raw + ((TABLE*4 + NB*NB + 52 + MAX_OLIGO + 15) & ~15));
	const int row_stride = (a.max_lt + 1)*ALIGN_THREADS;

	const int tid = threadIdx.x;
	for (int i = tid; i < TABLE; i += ALIGN_THREADS) s_dg[i] = a.thermo->dg[i];
	for (int i = tid; i < NB*NB; i += ALIGN_THREADS) s_bbp[i] = a.thermo->bbp[i];
	for (int i = tid; i < NPAIR; i += ALIGN_THREADS) s_wc[i] = a.thermo->wc[i];

	uint16_t *trace = a.trace + (size_t)blockIdx.x*a.trace_cells*ALIGN_THREADS + tid;
	int32_t *rowM = s_rows + tid, *rowIq = rowM + row_stride, *rowIt = rowIq + row_stride;
	unsigned long long my_cells = 0;

	uint32_t group_hint = 0xffffffffu;
	uint32_t cur_os = 0xffffffffu;
	// consecutive units per CTA, like the fast kernels: the oligo changes rarely
	const uint32_t units_per_cta = (a.nunits + gridDim.x - 1)/gridDim.x;
	const uint32_t u_end = min(a.nunits, (blockIdx.x + 1u)*units_per_cta);
	for (uint32_t u = blockIdx.x*units_per_cta; u < u_end; ++u) {
		const AlignUnit unit = unit_of(a.groups, a.ngroups, u, group_hint);
		const OligoStrand &os = a.os[unit.os];
		if (unit.os != cur_os) { // uniform across the block
			__syncthreads();
			for (int i = tid; i < os.len; i += ALIGN_THREADS) s_q[i] = os.seq[i];
			cur_os = unit.os;
			__syncthreads();
		}

		if ((uint32_t)tid >= unit.count) continue;

		DpShared sh;
		sh.dg = s_dg; sh.bbp = s_bbp; sh.wc = s_wc; sh.q = s_q; sh.Lq = os.len; sh.T = a.thermo->T;

		const Candidate c = a.cand[(size_t)unit.os*a.cap + unit.begin + tid];
		const uint32_t target = c.target_k & 0xffffffu, k = c.target_k >> 24;
		uint8_t tgt[MAX_WINDOW];
		uint32_t start = 0, stop = 0;
		int Lt;
		if (a.explicit_tgt) {
			Lt = a.explicit_len;
			for (int i = 0; i < Lt; ++i) tgt[i] = a.explicit_tgt[i];
			stop = (uint32_t)Lt;
		}
		else {
			const Target tg = a.db.targets[target];
			// window (bind_oligo.cpp:502-505)
			const int s0 = (int)c.t - (int)(k + NUM_FLANK);
			start = s0 > 0 ? (uint32_t)s0 : 0u;
			stop = min(start + (uint32_t)os.len + 2u*NUM_FLANK, tg.len);
			Lt = load_window(a.db, tg, start, stop, os.plus != 0, tgt);
		}
		my_cells += (unsigned long long)(os.len*Lt);

		unsigned flags = 0;
		AlnState work, best_aln;
		Best best;
		best.valid = false;
		best.dH = best.dS = best.tm = 0.0f;
		best_aln.b = best_aln.e = 2;
		best_aln.fm_q = best_aln.fm_t = best_aln.lm_q = best_aln.lm_t = 0;

		if (Lt > 0 && !a.thermo->dinkelbach) {
			const DpResult dp = nc_fill<ALIGN_THREADS>(sh, tgt, Lt, rowM, rowIq, rowIt, trace);
			RowMajorTrace<ALIGN_THREADS> tv;
			tv.trace = trace;
			tv.Lt = Lt;
			uint16_t cells[MAX_MAXCELLS];
			int cursor = dp.last_raise < 0 ? 0 : dp.last_raise, remaining = dp.nmax;
			bool fresh = true;
			do {
				const int ncells = collect_max_cells<ALIGN_THREADS>(tv, os.len, Lt, cursor, remaining, cells);
				nc_enumerate(sh, a.thermo, os.r_log_ct, tgt, Lt, tv, cells, ncells, work, best_aln, best, flags, fresh);
				fresh = false;
			} while (remaining > 0 && !(flags & (F_OOB | F_STACK)));
		}
		else if (Lt > 0) {
			// approximate_tm_heterodimer with use_dinkelbach (nuc_cruc.cpp:2399-2440): align at 0 degC, then
			// again and again at the melting temperature of the previous pass while dH - T*dS of the
			// pass (taken at the temperature it was aligned at) is negative and still rising.  The
			// record keeps the alignment of the last pass; dG is taken at the engine's temperature.
			DgAtT dgt;
			float T = 273.15f, q = -999999.9f, last_q;
			int iter = 0;
			do {
				dgt.set(a.thermo, T);
				sh.T = T;
				const DpResult dp = nc_fill<ALIGN_THREADS, DgAtT>(sh, dgt, tgt, Lt, rowM, rowIq, rowIt, trace);
				RowMajorTrace<ALIGN_THREADS> tv;
				tv.trace = trace;
				tv.Lt = Lt;
				uint16_t cells[MAX_MAXCELLS];
				int cursor = dp.last_raise < 0 ? 0 : dp.last_raise, remaining = dp.nmax;
				bool fresh = true;
				do {
					const int ncells = collect_max_cells<ALIGN_THREADS>(tv, os.len, Lt, cursor, remaining, cells);
					nc_enumerate(sh, a.thermo, os.r_log_ct, tgt, Lt, tv, cells, ncells, work, best_aln, best, flags, fresh);
					fresh = false;
				} while (remaining > 0 && !(flags & (F_OOB | F_STACK)));
				if (flags & (F_OOB | F_STACK)) break;
				if (fresh) { best.valid = false; best.dH = best.dS = best.tm = 0.0f; }
				last_q = q;
				q = __fsub_rn(best.dH, __fmul_rn(T, best.dS));
				T = __fadd_rn(273.15f, best.tm);
				if (++iter >= DINKELBACH_MAX_ITER) { flags |= F_STACK; break; } // q rises strictly: never expected
			} while (q < 0.0f && q > last_q);
			sh.T = a.thermo->T;
		}
		const uint32_t idx = unit.begin + tid;
		finish_alignment(a, sh, os, unit.os, target, k, c.t, start, stop, tgt, Lt, best, best_aln, flags,
			a.slot_map ? a.slot_map[idx] : idx);
	}

	// one atomic per warp for the cell counter
	for (int off = 16; off; off >>= 1) my_cells += __shfl_down_sync(0xffffffffu, my_cells, off);
	if ((tid & 31) == 0 && my_cells) atomicAdd(a.cells, my_cells);
}

// ------------------------------------------------------------------------------------------
// Oligo-only structures (tntblast_local.cpp:657-686): hairpin of one oligo, homodimer of one oligo,
// heterodimer of two oligos.  A handful of jobs per assay, one thread each; the DP rows live in
// local memory, the trace in a global scratch slab (cleared by the thread: the hairpin fill only
// writes a triangle).
// ------------------------------------------------------------------------------------------
enum { JOB_HETERODIMER = 0, JOB_HOMODIMER = 1, JOB_HAIRPIN = 2 };

struct OligoJob {
	int32_t kind, qlen, tlen;
	float r_log_ct;                // NC_R*log(Ct) of the duplex (unused for hairpins)
	uint8_t q[MAX_OLIGO];          // NucCruc codes, 5'->3'
	uint8_t t[MAX_WINDOW];         // second strand, 5'->3' (the query itself for homodimers / hairpins)
};

struct OligoJobResult { float tm, dH, dS; int32_t valid, flags; int16_t fm_q, fm_t, lm_q, lm_t; int32_t ncols; };

constexpr int JOB_TRACE_CELLS = MAX_OLIGO*MAX_WINDOW;

__global__ void __launch_bounds__(32) k_oligo_jobs(const OligoJob *__restrict__ jobs, uint32_t njobs, const Thermo *__restrict__ thermo,
	const Thermo *__restrict__ thermo_homo, uint16_t *__restrict__ trace_all, OligoJobResult *__restrict__ out)
{
	const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
	if (j >= njobs) return;
	const OligoJob &job = jobs[j];
	// homodimers carry the symmetry entropy in the initiation term (nuc_cruc.cpp:1632): own table copy
	const Thermo *th = job.kind == JOB_HOMODIMER ? thermo_homo : thermo;
	DpShared sh;
	sh.dg = th->dg; sh.bbp = th->bbp; sh.wc = th->wc; sh.q = job.q; sh.Lq = job.qlen; sh.T = th->T;
	const int Lt = job.tlen;
	uint16_t *trace = trace_all + (size_t)j*JOB_TRACE_CELLS;
	int32_t rowM[MAX_WINDOW + 1], rowIq[MAX_WINDOW + 1], rowIt[MAX_WINDOW + 1];
	unsigned flags = 0;
	AlnState work, best_aln;
	Best best;
	best.valid = false;
	best.dH = best.dS = best.tm = 0.0f;
	best_aln.b = best_aln.e = 2;
	best_aln.fm_q = best_aln.fm_t = best_aln.lm_q = best_aln.lm_t = 0;
	const bool hairpin = job.kind == JOB_HAIRPIN;
	const int tri = hairpin ? job.qlen - 4 : 0; // steric limit: three loop bases + one anchor (nuc_cruc.cpp:781-790)
	if (Lt > 0 && job.qlen > 0 && !(hairpin && tri <= 0)) {
		// with use_dinkelbach the three structure temperatures iterate like the heterodimer of a window
		// (nuc_cruc.cpp:2399-2440, :2459-2500, :2548-2588)
		const bool dink = th->dinkelbach != 0;
		DgAtT dgt;
		float T = dink ? 273.15f : th->T, q = -999999.9f, last_q;
		int iter = 0;
		do {
			for (int c = 0; c < job.qlen*Lt; ++c) trace[c] = 0;
			DpResult dp;
			if (dink) {
				dgt.set(th, T);
				sh.T = T;
				dp = nc_fill<1, DgAtT>(sh, dgt, job.t, Lt, rowM, rowIq, rowIt, trace, tri);
			}
			else dp = nc_fill<1>(sh, job.t, Lt, rowM, rowIq, rowIt, trace, tri);
			RowMajorTrace<1> tv;
			tv.trace = trace;
			tv.Lt = Lt;
			uint16_t cells[MAX_MAXCELLS];
			int cursor = dp.last_raise < 0 ? 0 : dp.last_raise, remaining = dp.nmax;
			bool fresh = true;
			do {
				const int ncells = collect_max_cells<1>(tv, job.qlen, Lt, cursor, remaining, cells);
				if (hairpin) nc_enumerate_hairpin(sh, th, job.t, Lt, tv, cells, ncells, work, best_aln, best, flags, fresh);
				else nc_enumerate(sh, th, job.r_log_ct, job.t, Lt, tv, cells, ncells, work, best_aln, best, flags, fresh);
				fresh = false;
			} while (remaining > 0 && !(flags & (F_OOB | F_STACK)));
			if (!dink || (flags & (F_OOB | F_STACK))) break;
			if (fresh) { best.valid = false; best.dH = best.dS = best.tm = 0.0f; }
			last_q = q;
			q = __fsub_rn(best.dH, __fmul_rn(T, best.dS));
			T = __fadd_rn(273.15f, best.tm);
			if (++iter >= DINKELBACH_MAX_ITER) { flags |= F_STACK; break; }
		} while (q < 0.0f && q > last_q);
	}
	OligoJobResult r;
	r.tm = best.tm; r.dH = best.dH; r.dS = best.dS;
	r.valid = best.valid ? 1 : 0;
	r.flags = (int32_t)flags;
	r.fm_q = (int16_t)best_aln.fm_q; r.fm_t = (int16_t)best_aln.fm_t;
	r.lm_q = (int16_t)best_aln.lm_q; r.lm_t = (int16_t)best_aln.lm_t;
	r.ncols = best.valid ? best_aln.e - best_aln.b : 0;
	out[j] = r;
}

// ------------------------------------------------------------------------------------------
// Measured integer-ALU peak (the denominator of the NucCruc roofline): every thread runs eight
// independent chains of one 32-bit instruction; MODE 0 = add (IADD3), 1 = max (IMNMX), 2 = the
// add-then-max pair of the DP recurrence (VIADDMNMX where the compiler fuses it).
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_alu_peak(int iters, int seed, int *out)
{
	int a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const int b = seed*3 + 1, c = seed - 7;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
			// operands come from the neighbouring chain, so that nothing folds into constants
			if (MODE == 0) { a0 += a1; a1 += a2; a2 += a3; a3 += a4; a4 += a5; a5 += a6; a6 += a7; a7 += a0; }
			else if (MODE == 1) {
				a0 = max(a0, a1); a1 = min(a1, a2); a2 = max(a2, a3); a3 = min(a3, a4);
				a4 = max(a4, a5); a5 = min(a5, a6); a6 = max(a6, a7); a7 = min(a7, a0 ^ u);
			}
			else {
				a0 = max(a0 - b, c); a1 = max(a1 - c, b); a2 = max(a2 - b, c); a3 = max(a3 - c, b);
				a4 = max(a4 - b, c); a5 = max(a5 - c, b); a6 = max(a6 - b, c); a7 = max(a7 - c, b);
			}
		}
	}
	if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x7fffffff) out[0] = a0; // keeps the chains alive
}

// Fast kernel: windows made of A/C/G/T only (the oligo may hold any code).  All units of one
// launch belong to oligo strands of at most LQ bases.
// Resident CTAs per SM the register allocation has to allow (other modes: experiments with
// -DTNT_OCC_MODE=n).  Mode 6 = four CTAs for rows <= 24 (128 registers, no spills in the lean tier,
// a few bytes in the full-trace tier), three up to 32, two up to 40.  Measured on B200: 100 TaqMan
// assays, lean tier 91.0 ms vs 98.2 ms without a bound, full-trace tier 9.2 ms vs 10.2 ms; 30-mer
// probes (class 32), lean tier 112 ms vs 126 ms with two CTAs.
#ifndef TNT_OCC_MODE
#define TNT_OCC_MODE 6
#endif
#if TNT_OCC_MODE == 1
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) ((FULL) ? 1 : ((LQ) <= 22 ? 5 : ((LQ) <= 28 ? 4 : 1))*(128/ALIGN_THREADS))
#elif TNT_OCC_MODE == 2
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) ((FULL) ? 1 : ((LQ) <= 28 ? 4 : 1)*(128/ALIGN_THREADS))
#elif TNT_OCC_MODE == 3
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) ((FULL) ? 1 : ((LQ) <= 24 ? 4 : ((LQ) <= 28 ? 3 : 1))*(128/ALIGN_THREADS))
#elif TNT_OCC_MODE == 4
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) ((FULL) ? 1 : ((LQ) <= 20 ? 5 : ((LQ) <= 24 ? 4 : ((LQ) <= 28 ? 3 : 1)))*(128/ALIGN_THREADS))
#elif TNT_OCC_MODE == 5
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) (((LQ) <= 24 ? 4 : ((LQ) <= 28 ? 3 : 1))*(128/ALIGN_THREADS))
#elif TNT_OCC_MODE == 6
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) (((LQ) <= 24 ? 4 : ((LQ) <= 32 ? 3 : ((LQ) <= 40 ? 2 : 1)))*(128/ALIGN_THREADS))
#else
#define TNT_FAST_MIN_BLOCKS(LQ, FULL) 1
#endif
template <int LQ, bool FULL>
__global__ void __launch_bounds__(ALIGN_THREADS, TNT_FAST_MIN_BLOCKS(LQ, FULL)) k_align_fast(AlignArgs a)
{
	constexpr int TAB_WORDS = FULL ? ROW_WORDS : LEAN_WORDS;
	__shared__ __align__(16) int32_t s_tab[LQ*TAB_WORDS];
	__shared__ int32_t s_p5[20];
	__shared__ uint8_t s_bbp[NB*NB];
	__shared__ uint8_t s_wc[52];
	__shared__ uint8_t s_q[MAX_OLIGO];

	const int tid = threadIdx.x;
	for (int i = tid; i < NB*NB; i += ALIGN_THREADS) s_bbp[i] = a.thermo->bbp[i];
	for (int i = tid; i < NPAIR; i += ALIGN_THREADS) s_wc[i] = a.thermo->wc[i];
	for (int i = tid; i < 20; i += ALIGN_THREADS) s_p5[i] = FULL ? a.p5_tab[i] : a.p5_tab[i]*LEAN_SCALE;

	// trace scratch: the lean tier's 2 bit/cell fit shared memory ([word][thread]: every lane its own
	// bank, for the column stores as well as for the per-lane reads of the traceback); the
	// full-trace tier keeps its 16 bit/cell in a per-CTA slab of global memory
	extern __shared__ __align__(16) uint32_t s_trace[];
	uint32_t *trace32 = FULL ? reinterpret_cast<uint32_t *>(a.trace) + (size_t)blockIdx.x*a.trace_cells*ALIGN_THREADS + tid
	                         : s_trace + tid;
	unsigned long long my_cells = 0;
	uint32_t cur_os = 0xffffffffu;

	uint32_t group_hint = 0xffffffffu;
	// A CTA takes consecutive units: they belong to the same oligo strand nearly always, so the strand's
	// table is loaded once per CTA instead of once per unit (with a stride of gridDim.x every unit of a
	// CTA was another strand: 5 KB of table and two block-wide barriers per unit; barrier stalls were
	// 10 % of the lean tier's stall reasons and 17 % of the full-trace tier's, ncu r02_v8 / r02_v9).
	const uint32_t units_per_cta = (a.nunits + gridDim.x - 1)/gridDim.x;
	const uint32_t u_end = min(a.nunits, (blockIdx.x + 1u)*units_per_cta);
	for (uint32_t u = blockIdx.x*units_per_cta; u < u_end; ++u) {
		const AlignUnit unit = unit_of(a.groups, a.ngroups, u, group_hint);
		const OligoStrand &os = a.os[unit.os];
		if (unit.os != cur_os) { // uniform across the block
			__syncthreads();
			const int32_t *src = (FULL ? a.row_tab : a.lean_tab) + (size_t)a.row_off[unit.os]*TAB_WORDS;
			const int nreal = os.len*TAB_WORDS;
			for (int i = tid; i < LQ*TAB_WORDS; i += ALIGN_THREADS) s_tab[i] = i < nreal ? src[i] : ROW_PAD_PENALTY;
			for (int i = tid; i < os.len; i += ALIGN_THREADS) s_q[i] = os.seq[i];
			cur_os = unit.os;
			__syncthreads();
		}
		if ((uint32_t)tid >= unit.count) continue;

		DpShared sh;
		sh.dg = nullptr; sh.bbp = s_bbp; sh.wc = s_wc; sh.q = s_q; sh.Lq = os.len; sh.T = a.thermo->T;

		const uint32_t idx = unit.begin + tid;
		const Candidate c = a.cand[(size_t)unit.os*a.cap + idx];
		const uint32_t target = c.target_k & 0xffffffu, k = c.target_k >> 24;
		const Target tg = a.db.targets[target];
		const int s0 = (int)c.t - (int)(k + NUM_FLANK);
		const uint32_t start = s0 > 0 ? (uint32_t)s0 : 0u;
		const uint32_t stop = min(start + (uint32_t)os.len + 2u*NUM_FLANK, tg.len);
		const int Lt = (int)(stop - start);

		// 2-bit window straight from the packed database (<= 64 bases -> three words)
		const uint64_t g0 = tg.base + start;
		const uint64_t wi = g0 >> 5;
		const unsigned sh2 = (unsigned)(g0 & 31u)*2u;
		const uint64_t w0 = __ldg(a.db.db2 + wi), w1 = __ldg(a.db.db2 + wi + 1), w2 = __ldg(a.db.db2 + wi + 2);
		uint64_t lo = sh2 ? ((w0 >> sh2) | (w1 << (64u - sh2))) : w0;
		uint64_t hi = sh2 ? ((w1 >> sh2) | (w2 << (64u - sh2))) : w1;
		// non-ACGT bases anywhere in the window? -> generic kernel
		const unsigned shm = (unsigned)(g0 & 31u);
		const uint64_t m01 = (uint64_t)__ldg(a.db.nmask + wi) | ((uint64_t)__ldg(a.db.nmask + wi + 1) << 32);
		const uint64_t m2 = (uint64_t)__ldg(a.db.nmask + wi + 2);
		uint64_t mwin = shm ? ((m01 >> shm) | (m2 << (64u - shm))) : m01;
		if (Lt < 64) mwin &= (1ull << Lt) - 1ull;
		if (mwin != 0 || Lt <= 0) {
			if (Lt > 0) {
				const uint32_t sl = atomicAdd(a.slow_count, 1u);
				if (sl < a.slow_cap) {
					SlowItem it;
					it.os = unit.os;
					it.slot = a.slot_map ? a.slot_map[idx] : idx;
					it.c = c;
					a.slow[sl] = it;
				}
				continue;
			}
		}
		my_cells += (unsigned long long)(os.len*Lt);

		// NucCruc target 5'->3': the window as is (plus strand) or its reverse complement, 2 bit/base
		PackedTgt tgt;
		tgt.lo = lo;
		tgt.hi = hi;
		if (!os.plus && Lt > 0) {
			// reverse the 2-bit groups of the 128-bit window, drop the unused tail, complement (3 - b == b ^ 3)
			const uint64_t odd = 0xaaaaaaaaaaaaaaaaull;
			uint64_t rl = __brevll(hi), rh = __brevll(lo);
			rl = ((rl & odd) >> 1) | ((rl & ~odd) << 1);
			rh = ((rh & odd) >> 1) | ((rh & ~odd) << 1);
			const unsigned drop = 128u - 2u*(unsigned)Lt; // Lt >= 1
			if (drop >= 64u) { rl = rh >> (drop - 64u); rh = 0; }
			else if (drop) { rl = (rl >> drop) | (rh << (64u - drop)); rh >>= drop; }
			tgt.lo = ~rl;
			tgt.hi = ~rh;
		}
		if (Lt <= 0) tgt.lo = tgt.hi = 0;
		else if (Lt < 32) { tgt.lo &= (1ull << (2*Lt)) - 1ull; tgt.hi = 0; }
		else if (Lt < 64) tgt.hi &= (1ull << (2*(Lt - 32))) - 1ull;
		const uint64_t tlo = tgt.lo, thi = tgt.hi;

		unsigned flags = 0;
		AlnState work, best_aln;
		Best best;
		best.valid = false;
		best.dH = best.dS = best.tm = 0.0f;
		best_aln.b = best_aln.e = 2;
		best_aln.fm_q = best_aln.fm_t = best_aln.lm_q = best_aln.lm_t = 0;

		// 0: done here, 1: retry with the full-trace fill, 2: generic kernel
		int handoff = 0;
		if (Lt > 0) {
			uint16_t cells[MAX_MAXCELLS];
			if constexpr (FULL) {
				const FastDpFull dp = nc_fill_fast_full<LQ, ALIGN_THREADS>(s_tab, s_p5, tlo, thi, Lt, trace32);
				ColMajorTraceFull<LQ, ALIGN_THREADS> tv;
				tv.trace32 = trace32;
				if (dp.runmax <= 0) handoff = 2; // the reference's ">= -1" rule decides: not recorded here
				else if (dp.nmax > MAX_MAXCELLS) handoff = 2; // the generic kernel walks any number of tied cells
				else {
					const int ncells = collect_max_cells_full<LQ, ALIGN_THREADS>(tv, dp, os.len, Lt, cells, flags);
					nc_enumerate(sh, a.thermo, os.r_log_ct, tgt, Lt, tv, cells, ncells, work, best_aln, best, flags);
				}
			}
			else {
				const FastDp dp = nc_fill_lean<LQ, ALIGN_THREADS>(s_tab, s_p5, tlo, thi, Lt, trace32);
				ColMajorLean<LQ, ALIGN_THREADS> tv;
				tv.trace32 = trace32;
				tv.tab = s_tab;
				tv.tgt = tgt;
				tv.maxscore = (int)(dp.runkey >> 12)*LEAN_SCALE;
				const int ncells = lean_max_cell(dp, Lt, cells);
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 7
				tnt_dbg_skip = a.out_count + 12;
#endif
#if defined(TNT_EXPERIMENT) && TNT_EXPERIMENT == 6
				if (tid == 0 && blockIdx.x == 0) tnt_dbg_why = a.out_count + 8;
				if (ncells == -2) atomicAdd(a.out_count + 8, 1u);
#endif
				if (ncells < 0) handoff = ncells == -1 ? 2 : 1;
				else if (!lean_finish(sh, a.thermo, os.r_log_ct, tv, Lt, cells[0], a.emit_all != 0, os.min_tm, os.max_tm, os.lean_min_cols, best_aln, best, flags)) handoff = 1;
			}
		}
		if (handoff) {
			my_cells -= (unsigned long long)(os.len*Lt); // counted again by the kernel that takes over
			const uint32_t out_slot = a.slot_map ? a.slot_map[idx] : idx;
			if (handoff == 1 && !FULL) {
				const uint32_t sl = atomicAdd(a.retry_fill + (size_t)unit.os*COUNT_STRIDE, 1u);
				if (sl < a.retry_cap[unit.os]) {
					const uint32_t at = a.retry_base[unit.os] + sl;
					a.retry_cand[at] = c;
					a.retry_slot[at] = out_slot;
				}
			}
			else {
				const uint32_t sl = atomicAdd(a.slow_count, 1u);
				if (sl < a.slow_cap) {
					SlowItem it;
					it.os = unit.os;
					it.slot = out_slot;
					it.c = c;
					a.slow[sl] = it;
				}
			}
			continue;
		}
		finish_alignment(a, sh, os, unit.os, target, k, c.t, start, stop, tgt, Lt, best, best_aln, flags,
			a.slot_map ? a.slot_map[idx] : idx);
	}

	for (int off = 16; off; off >>= 1) my_cells += __shfl_down_sync(0xffffffffu, my_cells, off);
	if ((tid & 31) == 0 && my_cells) atomicAdd(a.cells, my_cells);
}

// seq.h codes of [start, start+n) of one fragment (used to rebuild amplicon text for a hit)
__global__ void k_extract_codes(DbView db, uint32_t target, uint32_t start, uint32_t n, uint8_t *out)
{
	const Target tg = db.targets[target];
	for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		const uint64_t g = tg.base + start + i;
		int code = (int)((__ldg(db.db2 + (g >> 5)) >> ((g & 31u)*2u)) & 3u);
		if ((__ldg(db.nmask + (g >> 5)) >> (g & 31u)) & 1u) code = exception_code(db, tg, g);
		out[i] = (uint8_t)code;
	}
}

// The same for many ranges at once (tnt_engine_hit_sequences: the text of every hit of a search
// with one launch and one device-to-host copy): one CTA per range, grid-stride.
struct ExtractItem { uint32_t target, start, n, pad; uint64_t out_off; };

__global__ void __launch_bounds__(256) k_extract_many(DbView db, const ExtractItem *__restrict__ items, uint32_t nitems, uint8_t *__restrict__ out)
{
	for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x) {
		const ExtractItem x = items[it];
		const Target tg = db.targets[x.target];
		for (uint32_t i = threadIdx.x; i < x.n; i += blockDim.x) {
			const uint64_t g = tg.base + x.start + i;
			int code = (int)((__ldg(db.db2 + (g >> 5)) >> ((g & 31u)*2u)) & 3u);
			if ((__ldg(db.nmask + (g >> 5)) >> (g & 31u)) & 1u) code = exception_code(db, tg, g);
			out[x.out_off + i] = (uint8_t)code;
		}
	}
}

} // namespace tnt
