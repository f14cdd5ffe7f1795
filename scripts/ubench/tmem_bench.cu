// Micro-benchmark: tcgen05.ld latency / throughput for thread-private scratch use of tensor memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../concrete-numpy_b200/csrc/tmem.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int NW, int INFLIGHT>
__global__ void __launch_bounds__(256) k_tmem(long long* out, int iters, double* sink) {
    __shared__ unsigned s_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&s_base, 128);
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const unsigned mine = s_base + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)((warp >> 2) * 64);
    unsigned r[64];
    for (int i = 0; i < 64; i++) r[i] = threadIdx.x * 64 + i;
    TmemIo<64>::st(mine, r);
    tmem_wait_st();
    __syncthreads();
    long long t0 = clock64();
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned v[INFLIGHT][NW];
#pragma unroll
        for (int f = 0; f < INFLIGHT; f++) TmemIo<NW>::ld(mine + (f * NW) % (64 - NW + 1), v[f]);
        tmem_wait_ld();
#pragma unroll
        for (int f = 0; f < INFLIGHT; f++)
#pragma unroll
            for (int i = 0; i < NW; i++) acc += v[f][i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345u) *sink = 1.0;
    tmem_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 128);
}

template <int NW, int INFLIGHT>
int run(int nthreads, int ctas_per_sm) {
    long long* d; double* sink; CK(cudaMalloc(&d, 8 * 1024)); CK(cudaMalloc(&sink, 8));
    int iters = 2000;
    int grid = 148 * ctas_per_sm;
    k_tmem<NW, INFLIGHT><<<grid, nthreads>>>(d, 10, sink);
    k_tmem<NW, INFLIGHT><<<grid, nthreads>>>(d, iters, sink);
    CK(cudaDeviceSynchronize());
    long long h[1024]; CK(cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
    double bytes = (double)nthreads * NW * 4 * INFLIGHT * ctas_per_sm;
    printf("tcgen05.ld.32x32b.x%-2d inflight=%d threads=%d ctas/sm=%d: %.1f cycles/iter, %.1f B/clk/SM\n", NW, INFLIGHT,
           nthreads, ctas_per_sm, mx / iters, bytes * iters / mx);
    fflush(stdout);
    cudaFree(d); cudaFree(sink);
    return 0;
}

int main() {
    run<32, 1>(32, 1);
    run<32, 1>(128, 1);
    run<32, 1>(256, 1);
    run<64, 1>(256, 1);
    run<32, 2>(256, 1);
    run<16, 4>(256, 1);
    run<32, 1>(128, 3);
    run<64, 1>(128, 3);
    return 0;
}
