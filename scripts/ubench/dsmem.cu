// Development micro-benchmarks that size the PBS kernel design (DSMEM, cluster sync, L2 broadcast reads).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_dsmem(long long* out, int iters, int C) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* buf = reinterpret_cast<double2*>(smem);     // 64 KiB
    auto cl = cg::this_cluster();
    const int rank = cl.block_rank();
    for (int i = threadIdx.x; i < 4096; i += 256) buf[i] = make_double2(i, rank);
    cl.sync();
    long long t0 = clock64();
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {          // remote 16-byte stores, all-to-all
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int peer = (rank + 1 + ((r + it) % (C - 1))) % C;
                double2* p = cl.map_shared_rank(buf, peer);
                p[((r * 256 + threadIdx.x) + it * 7) & 4095] = make_double2(it, r);
            }
        } else if (MODE == 1) {   // remote 16-byte loads, 8 in flight
            double2 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int peer = (rank + 1 + ((r + it) % (C - 1))) % C;
                const double2* p = cl.map_shared_rank(buf, peer);
                v[r] = p[((r * 256 + threadIdx.x) + it * 7) & 4095];
            }
#pragma unroll
            for (int r = 0; r < 8; r++) { acc.x += v[r].x; acc.y += v[r].y; }
        } else if (MODE == 2) {   // local 16-byte loads for comparison
            double2 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = buf[((r * 256 + threadIdx.x) + it * 7) & 4095];
#pragma unroll
            for (int r = 0; r < 8; r++) { acc.x += v[r].x; acc.y += v[r].y; }
        } else if (MODE == 3) {   // cluster.sync cost
            cl.sync();
        } else if (MODE == 4) {   // remote 8-byte loads (rotation-like)
            unsigned long long v[16];
            const unsigned long long* b8 = reinterpret_cast<const unsigned long long*>(buf);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                int peer = (rank + 1 + ((r + it) % (C - 1))) % C;
                const unsigned long long* p = cl.map_shared_rank(b8, peer);
                v[r] = p[((r * 256 + threadIdx.x) + it * 7) & 8191];
            }
#pragma unroll
            for (int r = 0; r < 16; r++) acc.x += (double)v[r];
        }
    }
    if (MODE != 3) cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc.x == 123.456) out[0] = 0;
    cl.sync();
}

// every CTA streams the same `bytes` region (GGSW broadcast pattern) or its own region
__global__ void __launch_bounds__(256, 1) k_l2read(const double2* __restrict__ src, size_t elems_per_cta, int same,
                                                  int iters, double* sink) {
    const double2* p = src + (same ? 0 : (size_t)blockIdx.x * elems_per_cta);
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; it++) {
        for (size_t i = threadIdx.x; i < elems_per_cta; i += 256 * 8) {
            double2 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = __ldg(&p[(i + r * 256) % elems_per_cta]);
#pragma unroll
            for (int r = 0; r < 8; r++) { acc.x += v[r].x; acc.y += v[r].y; }
        }
    }
    if (acc.x == 123.456) *sink = acc.y;
}

template <int MODE>
int run_dsmem(const char* name, int C, int iters, double bytes_per_iter_per_cta) {
    long long* d; CK(cudaMalloc(&d, 8 * 256));
    CK(cudaFuncSetAttribute(k_dsmem<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 65536;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(C * (148 / C));
    for (int rep = 0; rep < 2; rep++) { CK(cudaLaunchKernelEx(&cfg, k_dsmem<MODE>, d, iters, C)); CK(cudaDeviceSynchronize()); }
    long long h[256]; CK(cudaMemcpy(h, d, 8 * cfg.gridDim.x, cudaMemcpyDeviceToHost));
    double mx = 0; for (unsigned i = 0; i < cfg.gridDim.x; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-28s C=%d  cycles/iter=%.1f  B/clk/SM=%.2f\n", name, C, mx / iters, bytes_per_iter_per_cta * iters / mx);
    cudaFree(d);
    return 0;
}

int main() {
    for (int C : {2, 4, 8}) {
        run_dsmem<0>("dsmem st.v2.f64 all-to-all", C, 2000, 256 * 8 * 16.0);
        run_dsmem<1>("dsmem ld.v2.f64 x8 inflight", C, 2000, 256 * 8 * 16.0);
        run_dsmem<4>("dsmem ld.u64 x16 inflight", C, 2000, 256 * 16 * 8.0);
        run_dsmem<3>("cluster.sync", C, 2000, 0);
    }
    run_dsmem<2>("local ld.v2.f64", 8, 2000, 256 * 8 * 16.0);
    // L2 broadcast read
    size_t elems = (2u << 20) / 16;   // 2 MiB per CTA-iteration region
    double2* src; double* sink; CK(cudaMalloc(&src, 148 * elems * 16)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(src, 0, 148 * elems * 16));
    for (int same = 1; same >= 0; same--)
        for (size_t per : {elems / 8, elems}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            int iters = 50;
            k_l2read<<<148, 256>>>(src, per, same, 2, sink);
            cudaEventRecord(e0);
            k_l2read<<<148, 256>>>(src, per, same, iters, sink);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double tot = 148.0 * per * 16 * iters;
            printf("L2 read %s region %zu KiB/CTA: %.1f GB/s aggregate, %.1f GB/s per SM\n", same ? "same" : "distinct",
                   per * 16 / 1024, tot / ms / 1e6, tot / ms / 1e6 / 148);
        }
    return 0;
}
