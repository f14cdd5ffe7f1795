// Micro-benchmark: all-to-all inside a cluster with bulk asynchronous DSMEM copies
// (cp.async.bulk.shared::cluster.shared::cta + mbarrier complete_tx) vs element stores.
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
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, const void* src, uint32_t bytes, uint32_t mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(smem_u32(src)), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}

// each iteration: every CTA sends CHUNK bytes to each of the C-1 peers
template <int CHUNK>
__global__ void __launch_bounds__(256, 1) k_bulk(long long* out, int iters, int C, int nissuers) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sendbuf = smem;                       // [C][CHUNK]
    unsigned char* recvbuf = smem + 8 * CHUNK;           // [C][CHUNK]
    __shared__ __align__(8) uint64_t bar;
    auto cl = cg::this_cluster();
    const int rank = cl.block_rank();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); }
    for (int i = threadIdx.x; i < 8 * CHUNK / 8; i += 256) ((double*)sendbuf)[i] = i + rank;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    cl.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (threadIdx.x == 0) mbar_expect_tx(&bar, (C - 1) * CHUNK);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if ((int)threadIdx.x < nissuers) {
            for (int r = 1 + threadIdx.x; r < C; r += nissuers) {
                int peer = (rank + r) % C;
                uint32_t dst = mapa(smem_u32(recvbuf + rank * CHUNK), peer);
                uint32_t mb = mapa(smem_u32(&bar), peer);
                bulk_copy_to_peer(dst, sendbuf + peer * CHUNK, CHUNK, mb);
            }
        }
        mbar_wait(&bar, it & 1);
        // peers may still be reading my sendbuf / I must not be overwritten before consuming: one cluster barrier per round
        cl.sync();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    cl.sync();
}

template <int CHUNK>
int run(int C, int iters, int nissuers) {
    long long* d; CK(cudaMalloc(&d, 8 * 256));
    size_t smem = 16 * CHUNK;
    CK(cudaFuncSetAttribute(k_bulk<CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(C * (148 / C));
    for (int rep = 0; rep < 2; rep++) { CK(cudaLaunchKernelEx(&cfg, k_bulk<CHUNK>, d, iters, C, nissuers)); CK(cudaDeviceSynchronize()); }
    long long h[256]; CK(cudaMemcpy(h, d, 8 * cfg.gridDim.x, cudaMemcpyDeviceToHost));
    double mx = 0; for (unsigned i = 0; i < cfg.gridDim.x; i++) mx = h[i] > mx ? h[i] : mx;
    printf("bulk dsmem all-to-all C=%d chunk=%d B issuers=%d: cycles/iter=%.1f (incl. one cluster.sync ~400)  B/clk/SM=%.2f  (excl. sync: %.2f)\n",
           C, CHUNK, nissuers, mx / iters, (double)(C - 1) * CHUNK * iters / mx, (double)(C - 1) * CHUNK / (mx / iters - 400.0));
    cudaFree(d);
    return 0;
}

int main() {
    for (int C : {2, 4, 8}) {
        run<4096>(C, 500, 1);
        run<4096>(C, 500, 7);
        run<8192>(C, 500, 7);
    }
    return 0;
}
