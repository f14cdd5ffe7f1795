// Micro-benchmark: L2 -> SM read bandwidth per SM as a function of active SMs and loads in flight.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(256, 1) k_read(const double2* __restrict__ src, size_t elems, int iters, double* sink) {
    double2 acc = make_double2(0, 0);
    const double2* p = src;
    for (int it = 0; it < iters; it++) {
        for (size_t i = threadIdx.x; i + (U - 1) * 256 < elems; i += 256 * U) {
            double2 v[U];
#pragma unroll
            for (int r = 0; r < U; r++) v[r] = __ldg(&p[i + r * 256]);
#pragma unroll
            for (int r = 0; r < U; r++) { acc.x += v[r].x; acc.y += v[r].y; }
        }
    }
    if (acc.x == 123.456) *sink = acc.y;
}

template <int U>
int run(const double2* src, double* sink, int grid, size_t elems) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 200;
    k_read<U><<<grid, 256>>>(src, elems, 2, sink);
    cudaEventRecord(e0);
    k_read<U><<<grid, 256>>>(src, elems, iters, sink);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tot = (double)grid * elems * 16 * iters;
    printf("grid=%3d loads in flight/thread=%2d region=%zu KiB: %.1f GB/s aggregate, %.1f GB/s per SM (%.1f B/clk @1.96GHz)\n",
           grid, U, elems * 16 / 1024, tot / ms / 1e6, tot / ms / 1e6 / grid, tot / ms / 1e6 / grid / 1.96);
    fflush(stdout);
    return 0;
}

int main() {
    size_t elems = (2u << 20) / 16;   // the same 2 MiB region read by every CTA (GGSW_i pattern)
    double2* src; double* sink; CK(cudaMalloc(&src, elems * 16)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(src, 0, elems * 16));
    for (int grid : {1, 8, 16, 37, 74, 148}) {
        run<8>(src, sink, grid, elems);
        run<16>(src, sink, grid, elems);
        run<32>(src, sink, grid, elems);
    }
    return 0;
}
