// Micro-benchmark: shared-memory wavefronts per LDS instruction for the access patterns the
// NucCruc fill could use (lanes differ only in a small table index).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_wavefronts lds_wavefronts.cu
// Run under: ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(const int *idx, int *out, int iters)
{
	__shared__ __align__(16) int tab[4096];
	for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i*7 + 1;
	__syncthreads();
	const int t = idx[threadIdx.x];   // lane-dependent table index
	int acc = 0;
	for (int it = 0; it < iters; ++it) {
		const int row = (it & 31)*72;
		if (MODE == 0) acc += tab[row + (t & 3)];                                   // 32-bit, 4 distinct
		if (MODE == 1) acc += tab[row + (t % 20)];                                  // 32-bit, 20 distinct
		if (MODE == 2) { const int2 v = *reinterpret_cast<const int2 *>(tab + row + 2*(t & 3)); acc += v.x ^ v.y; }   // 64-bit, 4 distinct
		if (MODE == 3) { const int2 v = *reinterpret_cast<const int2 *>(tab + row + 2*(t % 20)); acc += v.x ^ v.y; }  // 64-bit, 20 distinct
		if (MODE == 4) { const int4 v = *reinterpret_cast<const int4 *>(tab + row + 4*(t & 3)); acc += v.x ^ v.y ^ v.z ^ v.w; }  // 128-bit, 4 distinct
		if (MODE == 5) { const int4 v = *reinterpret_cast<const int4 *>(tab + row + 4*(t & 7)); acc += v.x ^ v.y ^ v.z ^ v.w; }  // 128-bit, 8 distinct
		if (MODE == 6) { const int4 v = *reinterpret_cast<const int4 *>(tab + row + 4*(t % 20)); acc += v.x ^ v.y ^ v.z ^ v.w; } // 128-bit, 20 distinct (conflicts)
		if (MODE == 7) { const int2 v = *reinterpret_cast<const int2 *>(tab + row + 2*(t & 15)); acc += v.x ^ v.y; }  // 64-bit, 16 distinct
	}
	out[blockIdx.x*blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const int *d_idx, int *d_out, const char *name)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	const int iters = 1 << 16;
	k<MODE><<<148*4, 128>>>(d_idx, d_out, 16);
	cudaEventRecord(a);
	k<MODE><<<148*4, 128>>>(d_idx, d_out, iters);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms = 0;
	cudaEventElapsedTime(&ms, a, b);
	// warp-level LDS per SM per microsecond
	const double lds = (double)iters*16.0;   // warp instructions per SM (4 CTAs x 4 warps)
	printf("%-28s %8.3f ms  %.3f ns per warp-LDS per SM\n", name, ms, ms*1e6/lds);
}

int main()
{
	int h[128];
	uint32_t s = 12345;
	for (int i = 0; i < 128; ++i) { s = s*1664525u + 1013904223u; h[i] = (int)(s >> 8) & 0xffff; }
	int *d_idx, *d_out;
	cudaMalloc(&d_idx, sizeof(h));
	cudaMalloc(&d_out, 148*4*128*sizeof(int));
	cudaMemcpy(d_idx, h, sizeof(h), cudaMemcpyHostToDevice);
	run<0>(d_idx, d_out, "lds32 4 distinct");
	run<1>(d_idx, d_out, "lds32 20 distinct");
	run<2>(d_idx, d_out, "lds64 4 distinct");
	run<3>(d_idx, d_out, "lds64 20 distinct");
	run<7>(d_idx, d_out, "lds64 16 distinct");
	run<4>(d_idx, d_out, "lds128 4 distinct");
	run<5>(d_idx, d_out, "lds128 8 distinct");
	run<6>(d_idx, d_out, "lds128 20 distinct");
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
