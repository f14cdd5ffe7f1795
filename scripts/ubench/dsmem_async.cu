// Micro-benchmark: st.async (remote store + mbarrier complete_tx) all-to-all, and how element-wise
// DSMEM throughput depends on the number of active clusters.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t dst, double a, double b, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(dst), "d"(a), "d"(b), "r"(mbar) : "memory");
}

// MODE 0: st.async all-to-all, completion through mbarrier (no cluster barrier for the data)
// MODE 1: plain st.shared::cluster all-to-all + cluster.sync
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_async(long long* out, int iters, int C) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* buf = reinterpret_cast<double2*>(smem);      // 64 KiB = 4096 double2
    __shared__ __align__(8) uint64_t bar;
    auto cl = cg::this_cluster();
    const int rank = cl.block_rank();
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    cl.sync();
    long long t0 = clock64();
    const uint32_t per_iter_bytes = 256 * 8 * 16;   // each CTA receives as much as it sends
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            if (threadIdx.x == 0) mbar_expect_tx(&bar, per_iter_bytes);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int peer = (rank + 1 + (r % (C - 1))) % C;
                uint32_t dst = mapa(smem_u32(&buf[(r * 256 + threadIdx.x) & 4095]), peer);
                uint32_t mb = mapa(smem_u32(&bar), peer);
                st_async_v2(dst, (double)it, (double)r, mb);
            }
            mbar_wait(&bar, it & 1);
            cl.sync();   // flow control: nobody may send round it+1 into a barrier still in round it
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int peer = (rank + 1 + (r % (C - 1))) % C;
                double2* p = cl.map_shared_rank(buf, peer);
                p[(r * 256 + threadIdx.x) & 4095] = make_double2(it, r);
            }
            cl.sync();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    cl.sync();
}

template <int MODE>
int run(const char* name, int C, int nclusters, int iters) {
    long long* d; CK(cudaMalloc(&d, 8 * 256));
    CK(cudaFuncSetAttribute(k_async<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 65536;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(C * nclusters);
    for (int rep = 0; rep < 2; rep++) { CK(cudaLaunchKernelEx(&cfg, k_async<MODE>, d, iters, C)); CK(cudaDeviceSynchronize()); }
    long long h[256]; CK(cudaMemcpy(h, d, 8 * cfg.gridDim.x, cudaMemcpyDeviceToHost));
    double mx = 0; for (unsigned i = 0; i < cfg.gridDim.x; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s C=%d clusters=%2d  cycles/iter=%.1f  B/clk/SM=%.2f (excl. 400-cycle sync %.2f)\n", name, C, nclusters, mx / iters, 256 * 8 * 16.0 * iters / mx, 256 * 8 * 16.0 / (mx / iters - 400.0)); fflush(stdout);
    cudaFree(d);
    return 0;
}

int main() {
    for (int C : {2, 8}) {
        for (int ncl : {1, 4, 148 / C}) {
            run<0>("st.async + mbarrier (32 KiB/iter)", C, ncl, 1000);
            run<1>("st + cluster.sync (32 KiB/iter)", C, ncl, 1000);
        }
    }
    return 0;
}
