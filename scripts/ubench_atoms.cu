// Micro-benchmark: cost of shared-memory histogram updates on sm_100a (decides the van Hove update scheme).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_atoms scripts/ubench_atoms.cu
// Patterns of the 32 bins a warp hits per update: 0 = all lanes one bin, 1 = 4 distinct bins, 2 = 32 consecutive
// bins (conflict-free), 3 = pseudo-random over 500 bins, 4 = pseudo-random over 16 bins.
// Variants: red.shared (no return), match.any + leader add, 4 samples per thread combined when equal.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 256;
constexpr int kIters = 4096;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ int pattern_bin(int pattern, int lane, uint32_t &seed, int it) {
    switch (pattern) {
        case 0: return 7;
        case 1: return 5 + (lane & 3);
        case 2: return lane + (it & 63);
        case 3: return (int)(lcg(seed) % 500u);
        default: return (int)(lcg(seed) % 16u);
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(kThreads, 2) hist_kernel(int pattern, unsigned int *out, long long *cycles) {
    __shared__ unsigned int s_hist[8 * 512];
    for (int i = threadIdx.x; i < 8 * 512; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        const int bin = pattern_bin(pattern, lane, seed, it);
        unsigned int *h = s_hist + (it & 7) * 512;
        if (VARIANT == 0) {
            atomicAdd(&h[bin], 1u);
        } else if (VARIANT == 1) {
            const unsigned m = __match_any_sync(0xffffffffu, bin);
            if ((m & ((1u << lane) - 1u)) == 0) atomicAdd(&h[bin], (unsigned)__popc(m));
        } else if (VARIANT == 2) {
            // uniformity test first (one redux-style compare), match only when it pays
            const int b0 = __shfl_sync(0xffffffffu, bin, 0);
            const unsigned same = __ballot_sync(0xffffffffu, bin == b0);
            if (same == 0xffffffffu) {
                if (lane == 0) atomicAdd(&h[bin], 32u);
            } else {
                atomicAdd(&h[bin], 1u);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned int acc = 0;
    for (int i = threadIdx.x; i < 8 * 512; i += kThreads) acc += s_hist[i];
    atomicAdd(out, acc);
    if (threadIdx.x == 0) atomicMax((unsigned long long *)cycles, (unsigned long long)(t1 - t0));
}

int main() {
    unsigned int *out;
    long long *cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8);
    const char *pn[] = {"one bin", "4 bins", "32 consecutive", "random/500", "random/16"};
    const char *vn[] = {"red.shared per lane", "match.any + leader", "ballot-uniform else red"};
    for (int v = 0; v < 3; ++v)
        for (int p = 0; p < 5; ++p) {
            cudaMemset(out, 0, 4);
            cudaMemset(cyc, 0, 8);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            const int grid = 148 * 2;
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(cyc, 0, 8);
                cudaEventRecord(e0);
                if (v == 0) hist_kernel<0><<<grid, kThreads>>>(p, out, cyc);
                if (v == 1) hist_kernel<1><<<grid, kThreads>>>(p, out, cyc);
                if (v == 2) hist_kernel<2><<<grid, kThreads>>>(p, out, cyc);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            long long c;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            // per SM: 2 CTAs x 8 warps x kIters warp-updates in c cycles
            printf("%-26s %-16s %8.3f ms  %6.2f cycles per warp-update per SM (16 warps resident)  err=%s\n", vn[v], pn[p], ms,
                   (double)c / (16.0 * kIters), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
