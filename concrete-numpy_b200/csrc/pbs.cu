// Batched programmable bootstrap for sm_100a: one thread-block cluster of C CTAs
// per ciphertext, the whole blind rotation (n CMUX iterations) runs on-chip.
//
//   * ACC (k+1 torus polynomials) lives in shared memory, split over the C CTAs in a "comb"
//     layout; CTA `rank` owns coefficients  j = h*M + j1*L + rank*Q + q  (h<2, j1<C, q<Q)
//     with M = N/2, L = M/C (local FFT length), Q = L/C.
//   * Negacyclic FFT (fp64): fold N reals to M complex, twist, then a 2-step Cooley-Tukey:
//     step A = radix-C butterflies over j1 in registers (all inputs are CTA-local thanks to
//     the comb layout), results are stored straight into the destination CTA's shared memory
//     (DSMEM), step B = length-L in-place DIF FFT in shared memory (radix-8 register
//     butterflies, XOR-swizzled to stay bank-conflict free).  The spectrum stays in
//     digit-reversed order; the Fourier BSK is produced by the same transform
//     (fhe_bsk_to_fourier) so the pointwise product needs no permutation.  The inverse runs the
//     mirrored DIT passes and the radix-C step reads its inputs from the peers' shared memory.
//   * Per CMUX: rotate+subtract+decompose (integer) -> (k+1)*level forward FFTs -> multiply-
//     accumulate against GGSW_i (streamed from L2/HBM, coalesced 16-byte loads, every CTA rank
//     reads only its 1/C slice) -> k+1 inverse FFTs -> round to torus and add into ACC.
//   * modulus switch is fused at the top, sample extract at the bottom.
//
// Replaces memref_batched_bootstrap_lwe_u64 of the absent concrete-compiler runtime; reached
// from concrete/numpy/compilation/server.py:286 for every FHE.apply_lookup_table the reference
// emits (concrete/numpy/mlir/node_converter.py:1129-1208).
#include <cooperative_groups.h>
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
namespace cg = cooperative_groups;

#define PBS_NT 256
#define PBS_MAXK 3          // k <= 3
#define PBS_MAXC 8
#define PBS_SMEM_MAX 232448 // 227 KB opt-in dynamic shared memory per CTA on sm_100

struct PbsGeom {
    int n, k, N, level, base_log;
    int logN, logM, logL, logQ, logC;
    int J;               // (k+1)*level forward transforms per CMUX
    int npass;           // local FFT passes
    int lr[8];           // log2 radix of each pass (forward order)
    double2 cj[PBS_MAXC];  // exp(i*pi*j1/(2C))
};

__device__ __forceinline__ int SW(int i) { return i ^ ((i >> 3) & 7); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
template <bool INV>
__device__ __forceinline__ double2 mul_mi(double2 a) {  // * (-i) forward, * (+i) inverse
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// natural-order in/out DFTs of size R in registers; forward kernel exp(-2*pi*i*r*m/R)
template <int R, bool INV>
struct Dft;
template <bool INV>
struct Dft<1, INV> {
    static __device__ __forceinline__ void run(double2*) {}
};
template <bool INV>
struct Dft<2, INV> {
    static __device__ __forceinline__ void run(double2* x) {
        double2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};
template <bool INV>
struct Dft<4, INV> {
    static __device__ __forceinline__ void run(double2* x) {
        double2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
        double2 t2 = cadd(x[1], x[3]), t3 = mul_mi<INV>(csub(x[1], x[3]));
        x[0] = cadd(t0, t2);
        x[1] = cadd(t1, t3);
        x[2] = csub(t0, t2);
        x[3] = csub(t1, t3);
    }
};
template <bool INV>
struct Dft<8, INV> {
    static __device__ __forceinline__ void run(double2* x) {
        const double h = 0.70710678118654752440;
        double2 e[4] = {x[0], x[2], x[4], x[6]};
        double2 o[4] = {x[1], x[3], x[5], x[7]};
        Dft<4, INV>::run(e);
        Dft<4, INV>::run(o);
        // o[m] *= w8^m
        double2 o1, o3;
        if (!INV) {
            o1 = make_double2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
            o3 = make_double2((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);
        } else {
            o1 = make_double2((o[1].x - o[1].y) * h, (o[1].x + o[1].y) * h);
            o3 = make_double2(-(o[3].x + o[3].y) * h, (o[3].x - o[3].y) * h);
        }
        double2 o2 = mul_mi<INV>(o[2]);
        x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
    }
};

// ---- one pass of the in-place local FFT over `nbuf` buffers of length L ------------
// forward (DIF): y_m = DFT_R(x)_m * W_Lc^{q m};   inverse (DIT): mirror with conjugates
template <int LR, bool INV>
__device__ __forceinline__ void fft_pass(double2* __restrict__ work, int nbuf, int logL, int logLc,
                                         const double2* __restrict__ tw_local) {
    constexpr int R = 1 << LR;
    const int L = 1 << logL;
    const int logS = logLc - LR;            // S = Lc / R
    const int S = 1 << logS;
    const int per_buf = L >> LR;            // butterflies per buffer
    const int tw_shift = logL - logLc;      // W_Lc^x = W_L^{x << tw_shift}
    for (int it = threadIdx.x; it < nbuf * per_buf; it += PBS_NT) {
        const int buf = it >> (logL - LR);
        const int w = it & (per_buf - 1);
        const int b = w >> logS, q = w & (S - 1);
        double2* base = work + ((size_t)buf << logL);
        const int pos0 = (b << logLc) + q;
        double2 x[R];
#pragma unroll
        for (int r = 0; r < R; r++) x[r] = base[SW(pos0 + (r << logS))];
        if (!INV) {
            Dft<R, false>::run(x);
            if (q != 0) {
#pragma unroll
                for (int m = 1; m < R; m++) x[m] = cmul(x[m], __ldg(&tw_local[(q * m) << tw_shift]));
            }
        } else {
            if (q != 0) {
#pragma unroll
                for (int m = 1; m < R; m++) x[m] = cmulc(x[m], __ldg(&tw_local[(q * m) << tw_shift]));
            }
            Dft<R, true>::run(x);
        }
#pragma unroll
        for (int r = 0; r < R; r++) base[SW(pos0 + (r << logS))] = x[r];
    }
}

template <bool INV>
__device__ __forceinline__ void fft_local(double2* work, int nbuf, const PbsGeom& g,
                                          const double2* __restrict__ tw_local) {
    if (!INV) {
        int logLc = g.logL;
        for (int p = 0; p < g.npass; p++) {
            const int lr = g.lr[p];
            if (lr == 3) fft_pass<3, false>(work, nbuf, g.logL, logLc, tw_local);
            else if (lr == 2) fft_pass<2, false>(work, nbuf, g.logL, logLc, tw_local);
            else fft_pass<1, false>(work, nbuf, g.logL, logLc, tw_local);
            logLc -= lr;
            __syncthreads();
        }
    } else {
        int logLc = 0;
        for (int p = g.npass - 1; p >= 0; p--) {
            const int lr = g.lr[p];
            logLc += lr;
            if (lr == 3) fft_pass<3, true>(work, nbuf, g.logL, logLc, tw_local);
            else if (lr == 2) fft_pass<2, true>(work, nbuf, g.logL, logLc, tw_local);
            else fft_pass<1, true>(work, nbuf, g.logL, logLc, tw_local);
            __syncthreads();
        }
    }
}

template <int C>
__device__ __forceinline__ void csync() {
    if (C == 1) __syncthreads();
    else cg::this_cluster().sync();
}
template <int C, typename T>
__device__ __forceinline__ T* peer(T* p, int rank) {
    if (C == 1) return p;
    return cg::this_cluster().map_shared_rank(p, rank);
}

__device__ __forceinline__ u64 f2torus(double x) {
    x -= 18446744073709551616.0 * rint(x * 5.42101086242752217e-20);
    return (u64)__double2ll_rn(x);
}
__device__ __forceinline__ int modswitch(u64 x, int logN) {  // -> Z_{2N}
    const int lg = logN + 1;
    u64 t = ((x >> (64 - lg - 1)) + 1) >> 1;
    return (int)(t & ((1ULL << lg) - 1));
}

// slot (local index inside one polynomial's CTA slice) -> coefficient index
template <int C>
__device__ __forceinline__ int slot_to_coef(int s, int rank, const PbsGeom& g) {
    const int q = s & ((1 << g.logQ) - 1);
    const int hj = s >> g.logQ;
    const int j1 = hj & (C - 1), h = hj >> g.logC;
    return (h << g.logM) + (j1 << g.logL) + (rank << g.logQ) + q;
}

// ---- the bootstrap kernel ----------------------------------------------------------
template <int C, int MAXL>
__global__ void __launch_bounds__(PBS_NT, 1)
k_pbs(u64* __restrict__ out, const u64* __restrict__ in, long B, const double2* __restrict__ bskF,
      const u64* __restrict__ luts, const int* __restrict__ lut_idx, const PbsGeom g,
      const double2* __restrict__ tw_local, const double2* __restrict__ tw_cross) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = g.N, M = N >> 1, L = 1 << g.logL, Q = 1 << g.logQ;
    const int S = N / C;                       // coefficients per polynomial per CTA
    const int k = g.k, level = g.level, J = g.J;
    u64* acc = reinterpret_cast<u64*>(smem_raw);                    // [(k+1)][S]
    double2* work = reinterpret_cast<double2*>(acc + (size_t)(k + 1) * S);  // [J][L]
    const int tid = threadIdx.x;
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const int logS = g.logN - g.logC;
    const size_t ggsw_stride = (size_t)J * (k + 1) * C * L;   // double2 per CMUX iteration

    for (long ct = cluster_id; ct < B; ct += nclusters) {
        const u64* lwe = in + (size_t)ct * (g.n + 1);
        const u64* lut = luts + (size_t)(lut_idx ? lut_idx[ct] : 0) * (k + 1) * N;
        // ---- ACC = X^{-b~} * LUT
        {
            const int a0 = (2 * N - modswitch(lwe[g.n], g.logN)) & (2 * N - 1);
            for (int w = tid; w < (k + 1) * S; w += PBS_NT) {
                const int c = w >> logS, s = w & (S - 1);
                int idx = slot_to_coef<C>(s, rank, g) - a0;
                bool neg = false;
                if (idx < 0) { idx += N; neg = !neg; }
                if (idx < 0) { idx += N; neg = !neg; }
                const u64 v = lut[(size_t)c * N + idx];
                acc[w] = neg ? (0ULL - v) : v;
            }
        }
        csync<C>();

        for (int i = 0; i < g.n; i++) {
            const int a = modswitch(lwe[i], g.logN);
            if (a == 0) continue;   // uniform over the whole cluster
            // ---- P1: rotate, subtract, decompose, fold+twist, radix-C step, scatter to peers
            for (int it = tid; it < ((k + 1) << g.logQ); it += PBS_NT) {
                const int c = it >> g.logQ, q = it & (Q - 1);
                const int j2 = (rank << g.logQ) + q;
                int dig[2 * C][MAXL];
                const u64* accc = acc + ((size_t)c << logS);
#pragma unroll
                for (int hj = 0; hj < 2 * C; hj++) {
                    const int h = hj / C, j1 = hj % C;
                    const u64 own = accc[(hj << g.logQ) + q];
                    int idx = (h << g.logM) + (j1 << g.logL) + j2 - a;
                    bool neg = false;
                    if (idx < 0) { idx += N; neg = !neg; }
                    if (idx < 0) { idx += N; neg = !neg; }
                    const int hh = idx >> g.logM, r = idx & (M - 1);
                    const int jj1 = r >> g.logL, jj2 = r & (L - 1);
                    const int ow = jj2 >> g.logQ, qq = jj2 & (Q - 1);
                    const int slot = (((hh << g.logC) + jj1) << g.logQ) + qq;
                    const u64 v = *peer<C>(accc + slot, ow);
                    const u64 diff = (neg ? (0ULL - v) : v) - own;
                    signed_decompose<MAXL>(diff, g.base_log, level, dig[hj]);
                }
                double2 T[C];
#pragma unroll
                for (int t1 = 0; t1 < C; t1++) T[t1] = __ldg(&tw_cross[(size_t)j2 * C + t1]);
#pragma unroll
                for (int t = 0; t < MAXL; t++) {
                    if (t < level) {
                        const int lv = level - 1 - t;       // dig[.][t] is the digit of level lv+1
                        double2 z[C];
#pragma unroll
                        for (int j1 = 0; j1 < C; j1++) {
                            double2 d = make_double2((double)dig[j1][t], (double)dig[C + j1][t]);
                            z[j1] = (j1 == 0) ? d : cmul(d, g.cj[j1]);
                        }
                        Dft<C, false>::run(z);
                        double2* dst = work + ((size_t)(c * level + lv) << g.logL) + SW(j2);
#pragma unroll
                        for (int t1 = 0; t1 < C; t1++) *peer<C>(dst, t1) = cmul(z[t1], T[t1]);
                    }
                }
            }
            csync<C>();
            // ---- P2: local forward FFTs of all J buffers
            fft_local<false>(work, J, g, tw_local);
            // ---- P3: multiply-accumulate against GGSW_i (this rank's slice)
            {
                const double2* gg = bskF + (size_t)i * ggsw_stride + ((size_t)rank << g.logL);
                for (int idx = tid; idx < L; idx += PBS_NT) {
                    double2 r[PBS_MAXK + 1];
#pragma unroll
                    for (int o = 0; o <= PBS_MAXK; o++) r[o] = make_double2(0.0, 0.0);
                    const int sidx = SW(idx);
                    for (int j = 0; j < J; j++) {
                        const double2 v = work[((size_t)j << g.logL) + sidx];
#pragma unroll
                        for (int o = 0; o <= PBS_MAXK; o++) {
                            if (o <= k) {
                                const double2 b = __ldg(&gg[((size_t)(j * (k + 1) + o) * C << g.logL) + idx]);
                                r[o].x += v.x * b.x - v.y * b.y;
                                r[o].y += v.x * b.y + v.y * b.x;
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o <= PBS_MAXK; o++)
                        if (o <= k) work[((size_t)o << g.logL) + sidx] = r[o];
                }
            }
            __syncthreads();
            // ---- P4: local inverse FFTs of the k+1 output buffers
            fft_local<true>(work, k + 1, g, tw_local);
            csync<C>();
            // ---- P5: gather from peers, inverse radix-C step, untwist, round, ACC +=
            {
                const double invM = 1.0 / (double)M;
                for (int it = tid; it < ((k + 1) << g.logQ); it += PBS_NT) {
                    const int o = it >> g.logQ, q = it & (Q - 1);
                    const int j2 = (rank << g.logQ) + q;
                    double2 v[C];
                    const double2* src = work + ((size_t)o << g.logL) + SW(j2);
#pragma unroll
                    for (int t1 = 0; t1 < C; t1++)
                        v[t1] = cmulc(*peer<C>(src, t1), __ldg(&tw_cross[(size_t)j2 * C + t1]));
                    Dft<C, true>::run(v);
                    u64* acco = acc + ((size_t)o << logS);
#pragma unroll
                    for (int j1 = 0; j1 < C; j1++) {
                        double2 z = (j1 == 0) ? v[0] : cmulc(v[j1], g.cj[j1]);
                        acco[(j1 << g.logQ) + q] += f2torus(z.x * invM);
                        acco[((C + j1) << g.logQ) + q] += f2torus(z.y * invM);
                    }
                }
            }
            csync<C>();
        }
        // ---- sample extract coefficient 0
        {
            u64* o = out + (size_t)ct * ((size_t)k * N + 1);
            for (int w = tid; w < (k + 1) * S; w += PBS_NT) {
                const int c = w >> logS, s = w & (S - 1);
                const int j = slot_to_coef<C>(s, rank, g);
                const u64 v = acc[w];
                if (c < k) {
                    if (j == 0) o[(size_t)c * N] = v;
                    else o[(size_t)c * N + (N - j)] = 0ULL - v;
                } else if (j == 0)
                    o[(size_t)k * N] = v;
            }
        }
        // the same thread re-initialises the same ACC words at the top of the loop
    }
}

// ---- standard-domain BSK -> Fourier BSK (same transform, same layout as the MAC reads) --
// in : [n][level][k+1 rows r][k+1 polys o][N] u64
// out: [n][J = r*level+lv][o][C ranks][L] double2 (digit-reversed local order)
template <int C>
__global__ void __launch_bounds__(PBS_NT, 1)
k_bsk_fourier(double2* __restrict__ outF, const u64* __restrict__ bsk, long npoly, const PbsGeom g,
              const double2* __restrict__ tw_local, const double2* __restrict__ tw_cross) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* work = reinterpret_cast<double2*>(smem_raw);   // [L]
    const int N = g.N, M = N >> 1, L = 1 << g.logL, Q = 1 << g.logQ;
    const int k = g.k, level = g.level;
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    for (long p = cluster_id; p < npoly; p += nclusters) {
        const u64* src = bsk + (size_t)p * N;
        for (int q = threadIdx.x; q < Q; q += PBS_NT) {
            const int j2 = (rank << g.logQ) + q;
            double2 z[C];
#pragma unroll
            for (int j1 = 0; j1 < C; j1++) {
                double2 d = make_double2((double)(i64)src[(j1 << g.logL) + j2],
                                         (double)(i64)src[M + (j1 << g.logL) + j2]);
                z[j1] = (j1 == 0) ? d : cmul(d, g.cj[j1]);
            }
            Dft<C, false>::run(z);
#pragma unroll
            for (int t1 = 0; t1 < C; t1++)
                *peer<C>(work + SW(j2), t1) = cmul(z[t1], __ldg(&tw_cross[(size_t)j2 * C + t1]));
        }
        csync<C>();
        fft_local<false>(work, 1, g, tw_local);
        // p = ((i*level + lv)*(k+1) + r)*(k+1) + o
        const int o = (int)(p % (k + 1));
        const int r = (int)((p / (k + 1)) % (k + 1));
        const int lv = (int)((p / ((long)(k + 1) * (k + 1))) % level);
        const long i = p / ((long)(k + 1) * (k + 1) * level);
        const int j = r * level + lv;
        double2* dst = outF + (((size_t)(i * g.J + j) * (k + 1) + o) * C + rank) * L;
        for (int idx = threadIdx.x; idx < L; idx += PBS_NT) dst[idx] = work[SW(idx)];
        csync<C>();
    }
}

// ---- debug / test entry: forward then inverse transform of arbitrary torus polynomials,
// optionally multiplied in the Fourier domain by a second (small integer) polynomial.
// Used by tests to validate the FFT against an exact integer negacyclic product.
template <int C>
__global__ void __launch_bounds__(PBS_NT, 1)
k_polymul_test(u64* __restrict__ out, const i64* __restrict__ a_small, const u64* __restrict__ b_torus,
               long npoly, const PbsGeom g, const double2* __restrict__ tw_local,
               const double2* __restrict__ tw_cross) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* work = reinterpret_cast<double2*>(smem_raw);   // [2][L]
    const int N = g.N, M = N >> 1, L = 1 << g.logL, Q = 1 << g.logQ;
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    for (long p = cluster_id; p < npoly; p += nclusters) {
        for (int it = threadIdx.x; it < 2 * Q; it += PBS_NT) {
            const int which = it >> g.logQ, q = it & (Q - 1);
            const int j2 = (rank << g.logQ) + q;
            double2 z[C];
#pragma unroll
            for (int j1 = 0; j1 < C; j1++) {
                double2 d;
                if (which == 0)
                    d = make_double2((double)a_small[(size_t)p * N + (j1 << g.logL) + j2],
                                     (double)a_small[(size_t)p * N + M + (j1 << g.logL) + j2]);
                else
                    d = make_double2((double)(i64)b_torus[(size_t)p * N + (j1 << g.logL) + j2],
                                     (double)(i64)b_torus[(size_t)p * N + M + (j1 << g.logL) + j2]);
                z[j1] = (j1 == 0) ? d : cmul(d, g.cj[j1]);
            }
            Dft<C, false>::run(z);
#pragma unroll
            for (int t1 = 0; t1 < C; t1++)
                *peer<C>(work + ((size_t)which << g.logL) + SW(j2), t1) =
                    cmul(z[t1], __ldg(&tw_cross[(size_t)j2 * C + t1]));
        }
        csync<C>();
        fft_local<false>(work, 2, g, tw_local);
        for (int idx = threadIdx.x; idx < L; idx += PBS_NT)
            work[SW(idx)] = cmul(work[SW(idx)], work[L + SW(idx)]);
        __syncthreads();
        fft_local<true>(work, 1, g, tw_local);
        csync<C>();
        const double invM = 1.0 / (double)M;
        for (int q = threadIdx.x; q < Q; q += PBS_NT) {
            const int j2 = (rank << g.logQ) + q;
            double2 v[C];
#pragma unroll
            for (int t1 = 0; t1 < C; t1++)
                v[t1] = cmulc(*peer<C>(work + SW(j2), t1), __ldg(&tw_cross[(size_t)j2 * C + t1]));
            Dft<C, true>::run(v);
#pragma unroll
            for (int j1 = 0; j1 < C; j1++) {
                double2 z = (j1 == 0) ? v[0] : cmulc(v[j1], g.cj[j1]);
                out[(size_t)p * N + (j1 << g.logL) + j2] = f2torus(z.x * invM);
                out[(size_t)p * N + M + (j1 << g.logL) + j2] = f2torus(z.y * invM);
            }
        }
        csync<C>();
    }
}

// ------------------------------------------------------------------ host side -------
struct PbsPlan {
    PbsGeom g;
    int C;
    size_t smem_pbs;
    double2* tw_local;
    double2* tw_cross;
};

static int choose_cluster(int N, int k, int level) {
    const int M = N / 2;
    for (int C = 1; C <= PBS_MAXC; C *= 2) {
        if (M / C < 64 || M / (C * C) < 1) break;
        size_t smem = (size_t)(k + 1) * (N / C) * 8 + (size_t)(k + 1) * level * (M / C) * 16;
        if (smem <= PBS_SMEM_MAX) return C;
    }
    return -1;
}

static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int, int, int, int>, PbsPlan> g_plans;

static int get_plan(int n, int k, int N, int level, int base_log, PbsPlan* out) {
    int dev = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mu);
    auto key = std::make_tuple(dev, N, k, level, base_log, n);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return 0; }
    const int logN = ilog2_exact(N);
    if (logN < 8 || logN > 15) FHE_FAIL(2, "pbs: polynomial size %d unsupported (256..32768)", N);
    if (k < 1 || k > PBS_MAXK) FHE_FAIL(2, "pbs: glwe dimension %d unsupported (1..%d)", k, PBS_MAXK);
    if (level < 1 || level > 4 || base_log < 1 || level * base_log > 52)
        FHE_FAIL(2, "pbs: decomposition level=%d base_log=%d unsupported", level, base_log);
    const int C = choose_cluster(N, k, level);
    if (C < 0) FHE_FAIL(3, "pbs: N=%d k=%d level=%d does not fit the on-chip layout", N, k, level);
    PbsPlan pl;
    pl.C = C;
    PbsGeom& g = pl.g;
    g.n = n; g.k = k; g.N = N; g.level = level; g.base_log = base_log;
    g.logN = logN; g.logM = logN - 1; g.logC = ilog2_exact(C);
    g.logL = g.logM - g.logC; g.logQ = g.logL - g.logC;
    g.J = (k + 1) * level;
    g.npass = 0;
    int rem = g.logL;
    while (rem >= 3) { g.lr[g.npass++] = 3; rem -= 3; }
    if (rem) g.lr[g.npass++] = rem;
    for (int j1 = 0; j1 < PBS_MAXC; j1++) {
        long double a = M_PIl * (long double)j1 / (2.0L * C);
        g.cj[j1] = make_double2((double)cosl(a), (double)sinl(a));
    }
    const int M = N / 2, L = M / C;
    std::vector<double2> tl(L), tc((size_t)L * C);
    for (int x = 0; x < L; x++) {
        long double a = -2.0L * M_PIl * (long double)x / (long double)L;
        tl[x] = make_double2((double)cosl(a), (double)sinl(a));
    }
    for (int j2 = 0; j2 < L; j2++)
        for (int t1 = 0; t1 < C; t1++) {
            // exp(i*pi*j2*(1-4*t1)/N); reduce the integer angle mod 2N first
            long num = ((long)j2 * (1 - 4 * t1)) % (2L * N);
            long double a = M_PIl * (long double)num / (long double)N;
            tc[(size_t)j2 * C + t1] = make_double2((double)cosl(a), (double)sinl(a));
        }
    FHE_CUDA(cudaMalloc(&pl.tw_local, sizeof(double2) * L));
    FHE_CUDA(cudaMalloc(&pl.tw_cross, sizeof(double2) * (size_t)L * C));
    FHE_CUDA(cudaMemcpy(pl.tw_local, tl.data(), sizeof(double2) * L, cudaMemcpyHostToDevice));
    FHE_CUDA(cudaMemcpy(pl.tw_cross, tc.data(), sizeof(double2) * (size_t)L * C, cudaMemcpyHostToDevice));
    pl.smem_pbs = (size_t)(k + 1) * (N / C) * 8 + (size_t)g.J * L * 16;
    g_plans[key] = pl;
    *out = pl;
    return 0;
}

template <typename Kern, typename... Args>
static int launch_clustered(Kern kern, int C, size_t smem, long work_items, cudaStream_t st, Args... args) {
    FHE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    FHE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(PBS_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int nclusters = 0;
    if (C > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3(C * (sms / C));
        FHE_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg));
        if (nclusters < 1) FHE_FAIL(4, "pbs: no cluster of %d CTAs with %zu B shared memory fits", C, smem);
    } else {
        int per_sm = 0;
        FHE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PBS_NT, smem));
        if (per_sm < 1) FHE_FAIL(4, "pbs: kernel does not fit on an SM (smem %zu)", smem);
        nclusters = per_sm * sms;
    }
    if ((long)nclusters > work_items) nclusters = (int)work_items;
    cfg.gridDim = dim3((unsigned)(nclusters * C));
    FHE_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

extern "C" {

// cluster size (CTAs per ciphertext) the on-chip layout uses for these parameters, <0 if unsupported
int fhe_pbs_cluster_size(int N, int k, int level) { return choose_cluster(N, k, level); }

// number of double2 elements of the Fourier BSK
long fhe_bsk_fourier_elems(int n, int k, int N, int level) {
    return (long)n * (k + 1) * level * (k + 1) * (N / 2);
}

int fhe_bsk_to_fourier(double2* bsk_f, const u64* bsk_std, int n, int k, int N, int level, int base_log,
                       void* stream) {
    PbsPlan pl;
    int rc = get_plan(n, k, N, level, base_log, &pl);
    if (rc) return rc;
    const long npoly = (long)n * level * (k + 1) * (k + 1);
    const size_t smem = sizeof(double2) << pl.g.logL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl.C) {
        case 1: return launch_clustered(k_bsk_fourier<1>, 1, smem, npoly, st, bsk_f, bsk_std, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        case 2: return launch_clustered(k_bsk_fourier<2>, 2, smem, npoly, st, bsk_f, bsk_std, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        case 4: return launch_clustered(k_bsk_fourier<4>, 4, smem, npoly, st, bsk_f, bsk_std, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        default: return launch_clustered(k_bsk_fourier<8>, 8, smem, npoly, st, bsk_f, bsk_std, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
    }
}

#define PBS_DISPATCH_L(CC)                                                                         \
    switch (level) {                                                                               \
        case 1: return launch_clustered(k_pbs<CC, 1>, CC, pl.smem_pbs, B, st, out, in, B, bsk_f, luts, lut_idx, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross); \
        case 2: return launch_clustered(k_pbs<CC, 2>, CC, pl.smem_pbs, B, st, out, in, B, bsk_f, luts, lut_idx, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross); \
        case 3: return launch_clustered(k_pbs<CC, 3>, CC, pl.smem_pbs, B, st, out, in, B, bsk_f, luts, lut_idx, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross); \
        default: return launch_clustered(k_pbs<CC, 4>, CC, pl.smem_pbs, B, st, out, in, B, bsk_f, luts, lut_idx, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross); \
    }

// out [B][kN+1], in [B][n+1] (small key), luts [T][(k+1)N] accumulators, lut_idx [B] or null
int fhe_pbs_batch(u64* out, const u64* in, long B, const double2* bsk_f, const u64* luts,
                  const int* lut_idx, int n, int k, int N, int level, int base_log, void* stream) {
    if (B < 0) FHE_FAIL(2, "pbs_batch: negative batch");
    if (B == 0) return 0;
    PbsPlan pl;
    int rc = get_plan(n, k, N, level, base_log, &pl);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl.C) {
        case 1: PBS_DISPATCH_L(1)
        case 2: PBS_DISPATCH_L(2)
        case 4: PBS_DISPATCH_L(4)
        default: PBS_DISPATCH_L(8)
    }
}

// test hook: out[p] = a_small[p] * b_torus[p] in Z_2^64[X]/(X^N+1) through the fp64 FFT path
int fhe_debug_polymul(u64* out, const i64* a_small, const u64* b_torus, long npoly, int N, int k,
                      int level, void* stream) {
    PbsPlan pl;
    int rc = get_plan(1, k, N, level, 8, &pl);
    if (rc) return rc;
    const size_t smem = (sizeof(double2) * 2) << pl.g.logL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl.C) {
        case 1: return launch_clustered(k_polymul_test<1>, 1, smem, npoly, st, out, a_small, b_torus, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        case 2: return launch_clustered(k_polymul_test<2>, 2, smem, npoly, st, out, a_small, b_torus, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        case 4: return launch_clustered(k_polymul_test<4>, 4, smem, npoly, st, out, a_small, b_torus, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
        default: return launch_clustered(k_polymul_test<8>, 8, smem, npoly, st, out, a_small, b_torus, npoly, pl.g, (const double2*)pl.tw_local, (const double2*)pl.tw_cross);
    }
}

}  // extern "C"
