// Micro-benchmark: the stage-1 exchange of the N = 32768 bootstrap plan (8 CTAs per cluster, every CTA sends
// 4 buffers x 8 chunks of 256 complex values, then reads its 4 x 2048 values back with radix-16 strides)
// routed (a) through distributed shared memory, (b) through global memory (L2), (c) L2 without clusters.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ double2 ldcg2(const double2* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stcg2(double2* p, double2 v) {
    asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// MODE 0: DSMEM scatter + local reads; 1: global scatter + global reads (cluster barriers); 2: as 1 but stores only;
// 3: as 1, loads only
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_xchg(long long* out, double2* xbuf, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* buf = reinterpret_cast<double2*>(smem);     // 4 x 2048 x 16 B = 128 KiB
    auto cl = cg::this_cluster();
    const int rank = cl.block_rank();
    const int tid = threadIdx.x;
    const size_t cluster_id = blockIdx.x / 8;
    double2* X = xbuf + cluster_id * (8 * 4 * 2048);      // [dest rank][buf][2048]
    for (int i = tid; i < 4 * 2048; i += 256) buf[i] = make_double2(i, rank);
    cl.sync();
    long long t_st = 0, t_s1 = 0, t_ld = 0, t_s2 = 0;
    double2 acc = make_double2(0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        long long a = clock64();
        if (MODE != 3) {
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                double2 v = make_double2(it + acc.x, b);
#pragma unroll
                for (int t1 = 0; t1 < 8; t1++) {
                    if (MODE == 0) {
                        double2* p = cl.map_shared_rank(buf, t1);
                        p[b * 2048 + rank * 256 + tid] = v;
                    } else
                        stcg2(&X[(size_t)(t1 * 4 + b) * 2048 + rank * 256 + tid], v);
                }
            }
        }
        long long b_ = clock64();
        cl.sync();
        long long c = clock64();
        if (MODE != 2) {
            const int grp = tid >> 7, w = tid & 127;
#pragma unroll 1
            for (int b = grp; b < 4; b += 2) {
                double2 x[16];
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    if (MODE == 0) x[r] = buf[b * 2048 + w + r * 128];
                    else x[r] = ldcg2(&X[(size_t)(rank * 4 + b) * 2048 + w + r * 128]);
                }
#pragma unroll
                for (int r = 0; r < 16; r++) { acc.x += x[r].x; acc.y += x[r].y; }
            }
        }
        long long d = clock64();
        cl.sync();
        long long e = clock64();
        t_st += b_ - a; t_s1 += c - b_; t_ld += d - c; t_s2 += e - d;
    }
    long long t1 = clock64();
    if (tid == 0) {
        out[blockIdx.x * 8 + 0] = t1 - t0;
        out[blockIdx.x * 8 + 1] = t_st; out[blockIdx.x * 8 + 2] = t_s1;
        out[blockIdx.x * 8 + 3] = t_ld; out[blockIdx.x * 8 + 4] = t_s2;
    }
    if (acc.x == 123.456) out[0] = 0;
}

// no clusters: every CTA writes 128 KiB then reads another CTA's 128 KiB (written in the previous iteration)
__global__ void __launch_bounds__(256, 1) k_l2rw(long long* out, double2* xbuf, int iters, int do_st, int do_ld) {
    const int tid = threadIdx.x;
    double2* mine = xbuf + (size_t)blockIdx.x * 8192;
    const double2* other = xbuf + (size_t)((blockIdx.x + 37) % gridDim.x) * 8192;
    double2 acc = make_double2(0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (do_st) {
#pragma unroll 8
            for (int r = 0; r < 32; r++) stcg2(&mine[r * 256 + tid], make_double2(it, r));
        }
        if (do_ld) {
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                double2 x[16];
#pragma unroll
                for (int r = 0; r < 16; r++) x[r] = ldcg2(&other[(h * 16 + r) * 256 + tid]);
#pragma unroll
                for (int r = 0; r < 16; r++) { acc.x += x[r].x; acc.y += x[r].y; }
            }
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x * 8] = t1 - t0;
    if (acc.x == 123.456) out[1] = 0;
}

template <int MODE>
int run(const char* name, int nclusters, int iters, double2* xbuf) {
    long long* d; CK(cudaMalloc(&d, 8 * 8 * 256));
    CK(cudaFuncSetAttribute(k_xchg<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 131072;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(8 * nclusters);
    for (int rep = 0; rep < 2; rep++) { CK(cudaLaunchKernelEx(&cfg, k_xchg<MODE>, d, xbuf, iters)); CK(cudaDeviceSynchronize()); }
    static long long h[8 * 256]; CK(cudaMemcpy(h, d, 8 * 8 * cfg.gridDim.x, cudaMemcpyDeviceToHost));
    double mx = 0, s[5] = {0, 0, 0, 0, 0};
    for (unsigned i = 0; i < cfg.gridDim.x; i++) {
        mx = h[i * 8] > mx ? h[i * 8] : mx;
        for (int q = 1; q < 5; q++) s[q] += (double)h[i * 8 + q] / cfg.gridDim.x / iters;
    }
    printf("%-34s clusters=%2d  cycles/iter=%7.0f  (store %6.0f  sync %6.0f  load %6.0f  sync %6.0f)  128 KiB each way -> %.1f B/clk/SM\n",
           name, nclusters, mx / iters, s[1], s[2], s[3], s[4], 131072.0 * iters / mx);
    cudaFree(d);
    return 0;
}

int main() {
    double2* xbuf; CK(cudaMalloc(&xbuf, (size_t)148 * 8192 * 16));
    CK(cudaMemset(xbuf, 0, (size_t)148 * 8192 * 16));
    for (int ncl : {1, 4, 15}) {
        run<0>("DSMEM scatter, local radix-16 reads", ncl, 500, xbuf);
        run<1>("global scatter, global reads", ncl, 500, xbuf);
        run<2>("global scatter only", ncl, 500, xbuf);
        run<3>("global reads only", ncl, 500, xbuf);
    }
    long long* d; CK(cudaMalloc(&d, 8 * 8 * 256));
    for (int grid : {120, 148})
        for (int mode = 1; mode <= 3; mode++) {
            int iters = 500;
            for (int rep = 0; rep < 2; rep++) { k_l2rw<<<grid, 256>>>(d, xbuf, iters, mode & 1, mode >> 1); CK(cudaDeviceSynchronize()); }
            static long long h[8 * 256]; CK(cudaMemcpy(h, d, 8 * 8 * grid, cudaMemcpyDeviceToHost));
            double mx = 0; for (int i = 0; i < grid; i++) mx = h[i * 8] > mx ? h[i * 8] : mx;
            printf("no clusters grid=%3d %s%s 128 KiB/CTA/iter: cycles/iter=%7.0f -> %.1f B/clk/SM per direction\n", grid,
                   (mode & 1) ? "st " : "", (mode & 2) ? "ld " : "", mx / iters, 131072.0 * iters / mx);
        }
    return 0;
}
