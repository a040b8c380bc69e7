// FASTA text -> seq.h base codes on the device (sm_100a).
//
// Replaces, for targets that arrive as FASTA text, the host reader of the reference:
//   sequence_data::load_fasta               sequence_data_fastx.cpp:13-79    record index
//   sequence_data::read_bio_seq_fasta_slow  sequence_data_fastx.cpp:190-382  defline + sequence characters
//   ascii_to_hash_base                      seq.h:148-189                    character -> code
// The reader is a four-state machine over the bytes of the file:
//   SEQ       sequence characters; the first '>' of a line opens a record (load_fasta :41-52)
//   LEAD      after that '>': white space in front of the defline is skipped (:254-257)
//   DEFLINE   up to the first '\n' or '\r' (:267-269)
//   SAMELINE  a defline that ended with '\r': sequence characters, but a '>' before the next
//             '\n' does not open a record (read_fasta is only reset by '\n', :47-52)
// Every byte range has an effect "entry state -> (exit state, bases, records, error)"; effects
// compose associatively, so the file is parsed by a three-kernel scan: per-block effects
// (k_fa_summary), one block that turns them into concrete entry states and offsets
// (k_fa_scan), and the emission pass (k_fa_emit) that writes the codes back to back (1 B/base,
// coalesced through shared memory) and one (byte offset, base offset) pair per record.
// HBM-bound byte work: the text is read twice, the codes written once (3 B per text byte).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tnt {

enum : uint32_t { FS_SEQ = 0, FS_DEFLINE = 1, FS_SAMELINE = 2, FS_LEAD = 3 };

// Effect of a byte range for one entry state, packed: exit state (2 bits) | error (1) | records (14) | bases (15)
constexpr uint32_t FA_ERR = 4u;
constexpr uint32_t FA_REC_ONE = 1u << 3;
constexpr uint32_t FA_BASE_ONE = 1u << 17;
constexpr int FA_THREADS = 256;
constexpr int FA_BYTES_PER_THREAD = 64;
constexpr int FA_BLOCK_BYTES = FA_THREADS*FA_BYTES_PER_THREAD; // 16 KB: 14-bit record and 15-bit base counts suffice
constexpr size_t FA_SLAB_BYTES = (size_t)64 << 20;            // text bytes parsed per pass (32-bit offsets inside a slab)
constexpr uint32_t FA_SCAN_THREADS = 1024;

struct FaMap { uint32_t v[4]; };

// state of the parse between slabs / at the entry of a block
struct FaCarry {
	uint64_t bases;
	uint64_t recs;
	uint32_t state;
	uint32_t err;
};

// lut[state][byte]: packed effect of one byte.  `code`: seq.h code of a sequence character.
struct FaTables {
	uint32_t lut[4][256];
	uint8_t code[256];
};

__device__ __forceinline__ uint32_t fa_step(uint32_t v, uint32_t e) { return ((v & ~3u) + (e & ~7u)) | (e & 7u); }

__device__ __forceinline__ uint32_t fa_pick(const FaMap &m, uint32_t s)
{
	return s == 0 ? m.v[0] : s == 1 ? m.v[1] : s == 2 ? m.v[2] : m.v[3];
}

// a followed by b
__device__ __forceinline__ FaMap fa_compose(const FaMap &a, const FaMap &b)
{
	FaMap r;
#pragma unroll
	for (int s = 0; s < 4; ++s) r.v[s] = fa_step(a.v[s], fa_pick(b, a.v[s] & 3u));
	return r;
}

__device__ __forceinline__ FaMap fa_identity()
{
	FaMap r;
	r.v[0] = 0; r.v[1] = 1; r.v[2] = 2; r.v[3] = 3;
	return r;
}

__device__ __forceinline__ void fa_load_tables(const FaTables *__restrict__ g, uint32_t (*s_lut)[256], uint8_t *s_code)
{
	for (uint32_t i = threadIdx.x; i < 4*256; i += blockDim.x) s_lut[i >> 8][i & 255u] = g->lut[i >> 8][i & 255u];
	if (s_code) for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = g->code[i];
}

// All four bytes >= 0x41: letters (and bytes >= 0x80).  The line ends, '>', the blanks, '*' and
// '-' all lie below 'A', so such a word holds four sequence characters whatever the state.
__device__ __forceinline__ bool fa_plain_word(uint32_t x)
{
	return ((((x & 0x7f7f7f7fu) + 0x3f3f3f3fu) | x) & 0x80808080u) == 0x80808080u;
}

// Effect of this thread's (up to) 64 bytes for all four entry states.  `w` receives the bytes.
__device__ __forceinline__ FaMap fa_thread_map(const uint8_t *__restrict__ text, uint32_t n, uint32_t off,
	const uint32_t (*s_lut)[256], uint32_t w[16], uint32_t &m)
{
	FaMap r = fa_identity();
	m = off < n ? min((uint32_t)FA_BYTES_PER_THREAD, n - off) : 0u;
	if (m == FA_BYTES_PER_THREAD) {
		const uint4 *p = reinterpret_cast<const uint4 *>(text + off);
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint4 x = __ldg(p + k);
			w[4*k] = x.x; w[4*k + 1] = x.y; w[4*k + 2] = x.z; w[4*k + 3] = x.w;
		}
	}
	else {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			uint32_t x = 0;
			for (int s = 0; s < 4; ++s) {
				const uint32_t i = 4u*k + s;
				if (i < m) x |= (uint32_t)text[off + i] << (8*s);
			}
			w[k] = x;
		}
	}
	// Words of four ordinary sequence characters need no table walk: +4 bases in the sequence
	// states (even), a defline stays a defline, LEAD has met its first defline character.  Which
	// words are ordinary differs from lane to lane, so they are found first (branch-free) and only
	// the others are walked, one per loop iteration (an 80-column file has about one per thread).
	uint32_t special = 0;
#pragma unroll
	for (int k = 0; k < 16; ++k)
		if (!(4u*k + 4u <= m && fa_plain_word(w[k])) && 4u*k < m) special |= 1u << k;
	const uint32_t nwords = (m + 3u)/4u;
	uint32_t done = 0; // words handled so far
	while (true) {
		const uint32_t k = special ? (uint32_t)__ffs((int)special) - 1u : nwords;
		const uint32_t plain = k - done;
		if (plain) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t st = r.v[q] & 3u;
				r.v[q] += (st & 1u) ? 0u : 4u*plain*FA_BASE_ONE;
				if (st == FS_LEAD) r.v[q] ^= (FS_LEAD ^ FS_DEFLINE);
			}
		}
		if (!special) break;
		special &= special - 1u;
		// the word again, from L1 (a dynamic index into w[] would put the array into local memory)
		uint32_t x;
		if (4u*k + 4u <= m) x = __ldg(reinterpret_cast<const uint32_t *>(text + off) + k);
		else {
			x = 0;
			for (uint32_t s = 0; 4u*k + s < m; ++s) x |= (uint32_t)text[off + 4u*k + s] << (8u*s);
		}
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			if (4u*k + s < m) {
				const uint32_t c = (x >> (8*s)) & 0xffu;
#pragma unroll
				for (int q = 0; q < 4; ++q) r.v[q] = fa_step(r.v[q], s_lut[r.v[q] & 3u][c]);
			}
		}
		done = k + 1u;
	}
	return r;
}

__device__ __forceinline__ FaMap fa_shfl_up(const FaMap &m, int d)
{
	FaMap r;
#pragma unroll
	for (int s = 0; s < 4; ++s) r.v[s] = __shfl_up_sync(0xffffffffu, m.v[s], d);
	return r;
}

// Block-wide scan in thread order.  Returns the effect of all earlier threads of the block
// (exclusive); `total` = effect of the whole block (valid in every thread).
__device__ __forceinline__ FaMap fa_block_exclusive(const FaMap &mine, FaMap *s_warp, FaMap &total)
{
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	FaMap incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const FaMap prev = fa_shfl_up(incl, d);
		if (lane >= (unsigned)d) incl = fa_compose(prev, incl);
	}
	if (lane == 31) s_warp[warp] = incl;
	FaMap excl = fa_shfl_up(incl, 1);
	if (lane == 0) excl = fa_identity();
	__syncthreads();
	FaMap before = fa_identity();
	total = fa_identity();
	for (unsigned k = 0; k < FA_THREADS/32; ++k) {
		if (k == warp) before = total;
		total = fa_compose(total, s_warp[k]);
	}
	return fa_compose(before, excl);
}

// Pass 1: effect of every 16 KB block.
__global__ void __launch_bounds__(FA_THREADS) k_fa_summary(const uint8_t *__restrict__ text, uint32_t n,
	const FaTables *__restrict__ tables, uint4 *__restrict__ block_map)
{
	__shared__ uint32_t s_lut[4][256];
	__shared__ FaMap s_warp[FA_THREADS/32];
	fa_load_tables(tables, s_lut, nullptr);
	__syncthreads();
	uint32_t w[16], m;
	const uint32_t off = blockIdx.x*(uint32_t)FA_BLOCK_BYTES + threadIdx.x*(uint32_t)FA_BYTES_PER_THREAD;
	const FaMap mine = fa_thread_map(text, n, off, s_lut, w, m);
	FaMap total;
	(void)fa_block_exclusive(mine, s_warp, total);
	if (threadIdx.x == 0) block_map[blockIdx.x] = make_uint4(total.v[0], total.v[1], total.v[2], total.v[3]);
}

// Pass 2 (one block): concrete entry state and offsets of every block, updated carry.
// Wide effects: exit state (2) | error (1) | records << 3 (29 bits) | bases << 32.
__device__ __forceinline__ uint64_t fa_widen(uint32_t v)
{
	return (uint64_t)(v & 7u) | ((uint64_t)((v >> 3) & 0x3fffu) << 3) | ((uint64_t)(v >> 17) << 32);
}
__device__ __forceinline__ uint64_t fa_step64(uint64_t v, uint64_t e) { return ((v & ~3ull) + (e & ~7ull)) | (e & 7ull); }

__global__ void __launch_bounds__(FA_SCAN_THREADS) k_fa_scan(const uint4 *__restrict__ block_map, uint32_t nblocks,
	FaCarry *__restrict__ carry, FaCarry *__restrict__ block_entry)
{
	__shared__ uint64_t s_map[FA_SCAN_THREADS][4];
	__shared__ uint64_t s_entry[FA_SCAN_THREADS];
	// at least 16 blocks per thread: the serial pass of thread 0 below is the latency of this kernel
	const uint32_t per = max(16u, (nblocks + FA_SCAN_THREADS - 1)/FA_SCAN_THREADS);
	const uint32_t nact = (nblocks + per - 1)/per;
	const uint32_t b0 = threadIdx.x*per;
	uint64_t acc[4] = {0, 1, 2, 3};
	for (uint32_t k = 0; k < per && b0 + k < nblocks; ++k) {
		const uint4 bm = __ldg(block_map + b0 + k);
		const uint32_t v[4] = {bm.x, bm.y, bm.z, bm.w};
#pragma unroll
		for (int s = 0; s < 4; ++s) acc[s] = fa_step64(acc[s], fa_widen(v[acc[s] & 3u]));
	}
#pragma unroll
	for (int s = 0; s < 4; ++s) s_map[threadIdx.x][s] = acc[s];
	__syncthreads();
	if (threadIdx.x == 0) {
		// counts relative to the slab; the carry's own counts are added when the entries are written
		uint64_t cur = (uint64_t)(carry->state & 3u) | (carry->err ? FA_ERR : 0u);
		for (uint32_t k = 0; k < nact; ++k) {
			s_entry[k] = cur;
			cur = fa_step64(cur, s_map[k][cur & 3u]);
		}
		// s_map[0] is no longer needed by thread 0: keep the slab total there
		s_map[0][0] = cur;
	}
	__syncthreads();
	const uint64_t base0 = carry->bases, rec0 = carry->recs;
	uint64_t cur = s_entry[threadIdx.x];
	for (uint32_t k = 0; k < per && b0 + k < nblocks; ++k) {
		FaCarry en;
		en.state = (uint32_t)(cur & 3u);
		en.err = (uint32_t)((cur >> 2) & 1u);
		en.recs = rec0 + ((cur >> 3) & 0x1fffffffull);
		en.bases = base0 + (cur >> 32);
		block_entry[b0 + k] = en;
		const uint4 bm = __ldg(block_map + b0 + k);
		const uint32_t v[4] = {bm.x, bm.y, bm.z, bm.w};
		cur = fa_step64(cur, fa_widen(v[cur & 3u]));
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint64_t tot = s_map[0][0];
		carry->state = (uint32_t)(tot & 3u);
		carry->err = (uint32_t)((tot >> 2) & 1u);
		carry->recs = rec0 + ((tot >> 3) & 0x1fffffffull);
		carry->bases = base0 + (tot >> 32);
	}
}

__device__ __forceinline__ uint32_t fa_pad(uint32_t i) { return i + 8u*(i >> 7); }

// Pass 3: codes and record table.  `text_pos0` = offset of the slab in the caller's text,
// `codes` = the output array for the whole text (base index counted from the first record).
__global__ void __launch_bounds__(FA_THREADS) k_fa_emit(const uint8_t *__restrict__ text, uint32_t n, uint64_t text_pos0,
	const FaTables *__restrict__ tables, const FaCarry *__restrict__ block_entry,
	uint8_t *__restrict__ codes, uint64_t *__restrict__ rec_pos, uint64_t *__restrict__ rec_base)
{
	__shared__ uint32_t s_lut[4][256];
	__shared__ uint8_t s_code[256];
	__shared__ FaMap s_warp[FA_THREADS/32];
	// Codes of the block, compacted.  A lane's output starts ~63 bytes after its neighbour's, i.e.
	// every second lane would hit the same bank (ncu: half of the shared-memory wavefronts of the
	// first version were bank conflicts): two pad words per 128 bytes spread the lanes over the banks.
	__shared__ __align__(16) uint8_t s_out[FA_BLOCK_BYTES + 8*(FA_BLOCK_BYTES/128)];
	fa_load_tables(tables, s_lut, s_code);
	__syncthreads();
	uint32_t w[16], m;
	const uint32_t off = blockIdx.x*(uint32_t)FA_BLOCK_BYTES + threadIdx.x*(uint32_t)FA_BYTES_PER_THREAD;
	const FaMap mine = fa_thread_map(text, n, off, s_lut, w, m);
	FaMap total;
	const FaMap before = fa_block_exclusive(mine, s_warp, total);
	const FaCarry en = block_entry[blockIdx.x];
	const uint32_t pv = fa_pick(before, en.state);
	uint32_t st = pv & 3u;
	uint32_t nb = pv >> 17;                    // bases of the block in front of this thread
	uint64_t rec = en.recs + ((pv >> 3) & 0x3fffu);
#pragma unroll
	for (int k = 0; k < 16; ++k) {
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			if (4u*k + s < m) {
				const uint32_t c = (w[k] >> (8*s)) & 0xffu;
				const uint32_t e = s_lut[st][c];
				if (e & FA_BASE_ONE) { s_out[fa_pad(nb)] = s_code[c]; ++nb; }
				if (e & FA_REC_ONE) {
					rec_pos[rec] = text_pos0 + off + 4u*k + s;
					rec_base[rec] = en.bases + nb;
					++rec;
				}
				st = e & 3u;
			}
		}
	}
	__syncthreads();
	const uint32_t block_bases = fa_pick(total, en.state) >> 17;
	uint8_t *dst = codes + en.bases;
	// head up to a 4-byte boundary of the destination, body, tail
	const uint32_t head = min(block_bases, (uint32_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u));
	if (threadIdx.x < head) dst[threadIdx.x] = s_out[fa_pad(threadIdx.x)];
	// body: one aligned 32-bit store per lane (consecutive lanes read consecutive words of s_out:
	// no bank conflicts; a warp writes 128 contiguous bytes)
	const uint32_t nvec = (block_bases - head)/4u;
	for (uint32_t k = threadIdx.x; k < nvec; k += FA_THREADS) {
		const uint32_t at = head + 4u*k;
		const uint32_t x = (uint32_t)s_out[fa_pad(at)] | ((uint32_t)s_out[fa_pad(at + 1)] << 8) |
			((uint32_t)s_out[fa_pad(at + 2)] << 16) | ((uint32_t)s_out[fa_pad(at + 3)] << 24);
		*reinterpret_cast<uint32_t *>(dst + at) = x;
	}
	const uint32_t done = head + 4u*nvec;
	if (threadIdx.x < block_bases - done) dst[done + threadIdx.x] = s_out[fa_pad(done + threadIdx.x)];
}

} // namespace tnt
