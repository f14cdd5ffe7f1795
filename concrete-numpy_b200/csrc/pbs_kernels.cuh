// Batched programmable bootstrap for sm_100a -- kernel templates.
//
// One thread-block cluster of C CTAs per ciphertext (C = 1 for N <= 4096); the whole blind rotation
// (n CMUX iterations) runs on-chip:
//
//   * ACC (k+1 torus polynomials) lives in shared memory, split over the C CTAs in a "comb"
//     layout: CTA `rank` owns coefficients  h*M + j1*L1 + rank*Q + q  (h<2, j1<R1, q<Q),
//     M = N/2 = R1*L1, Q = L1/C.  (C == 1: R1 is simply the radix of the first FFT stage.)
//   * Negacyclic product through an fp64 FFT of M complex points (fold + twist), factored as
//       stage 1 : radix-R1 butterflies over j1, fused with rotate/subtract/decompose; all inputs are
//                 CTA-local thanks to the comb layout, outputs go straight to the CTA that owns
//                 sub-transform t1 (distributed shared memory when C > 1);
//       mid     : one or two radix-16/8 passes in shared memory, register butterflies, twiddles in
//                 shared memory and re-used across the J = (k+1)*level buffers;
//       fused   : last radix-RF pass + multiply-accumulate against GGSW_i + first inverse pass, all
//                 in registers (GGSW streamed from L2 with coalesced, software-pipelined loads);
//       then the mirrored inverse passes, and the inverse radix-R1 stage fused with untwist /
//       round-to-torus / ACC update.
//   * The spectrum never leaves its (digit-reversed, thread-major) order: fhe_bsk_to_fourier runs the
//     same forward code, so the pointwise product needs no permutation.
//   * modulus switch is fused at the top, sample extract at the bottom.
//
// Replaces memref_batched_bootstrap_lwe_u64 of the absent concrete-compiler runtime; reached from
// concrete/numpy/compilation/server.py:286 for every FHE.apply_lookup_table the reference emits
// (concrete/numpy/mlir/node_converter.py:1129-1208).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "tmem.cuh"
namespace cg = cooperative_groups;

#define PBS_SMEM_MAX 232448  // 227 KB opt-in dynamic shared memory per CTA on sm_100

struct PbsGeom {
    int n, k, N, level, base_log;
    int logN, logM;
    int J;                // (k+1)*level forward transforms per CMUX
    double inv_M;
    int dbg;              // development ablation bits (FHE_PBS_DBG), 0 in production
    long long* prof;      // phase-cycle buffer (fhe_pbs_set_profile) or null
    int stagger;          // start-up delay per cluster id (SM cycles), de-synchronises the clusters' phases
    // extended bootstrap (k_pbs<P, true>): the modulus switch keeps `theta` zero low bits, so one blind rotation
    // serves 2^theta interleaved test polynomials and coefficients 0..nout-1 are extracted (nout <= 2^theta);
    // 0 / 1 for a plain bootstrap
    int theta, nout;
    double2 cj[16];       // exp(i*pi*j1/(2*R1)), j1 < R1
};

// Compile-time FFT plan.  M = R1 * L1, L1 = RA * RB * RF (RA / RB = 1 when absent).
template <int C_, int LOGR1_, int LRA_, int LRB_, int LRF_, int NT_, int MINB_, int KMAX_, int ACCTM_ = 0, int GROUP_ = 0>
struct PbsPlan {
    // GROUP_ > 1 (single-CTA plans): k = 1 bootstraps whose GGSW fits next to GROUP_ ciphertexts run k_pbs_group
    // (pbs_group.cuh): one CTA per SM carries GROUP_ ciphertexts and stages every GGSW_i once for all of them
    static constexpr int GROUP = GROUP_;
    static constexpr int C = C_, LOGC = (C_ == 1 ? 0 : C_ == 2 ? 1 : C_ == 4 ? 2 : C_ == 8 ? 3 : 4);
    static constexpr int LOGR1 = LOGR1_, R1 = 1 << LOGR1_;
    static constexpr int LRA = LRA_, LRB = LRB_, LRF = LRF_, RF = 1 << LRF_;
    static constexpr int NT = NT_, MINB = MINB_, KMAX = KMAX_;
    static constexpr int LOGL1 = LRA_ + LRB_ + LRF_, L1 = 1 << LOGL1;
    static constexpr int LOGM = LOGR1_ + LOGL1, M = 1 << LOGM, N = 2 << LOGM;
    static constexpr int LOGBUF = (C_ > 1) ? LOGL1 : LOGM, BUF = 1 << LOGBUF;   // double2 per work buffer per CTA
    static constexpr int LOGQ = LOGL1 - LOGC, Q = 1 << LOGQ;
    static constexpr int LOGS = LOGM + 1 - LOGC, S = 1 << LOGS;                  // ACC words per polynomial per CTA
    static constexpr int SA = 1 << (LOGL1 - LRA_);        // stride of mid pass A (sub-length L1)
    static constexpr int SB = 1 << LRF_;                  // stride of mid pass B (sub-length L1 / RA)
    static constexpr int TWA = LRA_ ? ((1 << LRA_) - 1) * SA : 0;
    static constexpr int TWB = LRB_ ? ((1 << LRB_) - 1) * SB : 0;
    static constexpr int TWN = TWA + TWB;                 // entries of the twiddle table of the mid passes
    static constexpr int TW_SMEM = (MINB_ == 1) ? 0 : TWA + TWB;   // ... held in shared memory unless TMEM_TW
    // Tensor-memory columns per thread: the R-1 twiddles of the thread's (fixed) butterfly index in each mid
    // pass and the R1 stage-1 twiddles of each of its NQ item columns, padded to tcgen05 ld/st sizes
    static constexpr int tm_pad(int words) { return words <= 4 ? 4 : words <= 8 ? 8 : words <= 16 ? 16 : words <= 32 ? 32 : 64; }
    static constexpr int TM_A = LRA_ ? tm_pad(4 * ((1 << LRA_) - 1)) : 0;
    static constexpr int TM_B = LRB_ ? tm_pad(4 * ((1 << LRB_) - 1)) : 0;
    // Measured (profiles/r01_ubench_tmem.log + A/B runs): tcgen05.ld costs ~75 cycles when the memory pipe is
    // quiet but queues behind outstanding global / remote accesses of the SM, so tensor memory holds only the
    // mid-pass twiddles and only in the plans that run one CTA per SM; the others keep them in shared memory.
    static constexpr bool TMEM_TW = (MINB_ == 1);
    // ACC_TM (cluster plans, k = 1): the accumulator words a thread owns stay in ITS tensor-memory lane (the same
    // thread reads them in P1 and updates them in P5), and the shared-memory array holds the ROTATED accumulator of
    // the next CMUX instead: every update is pushed (st.shared::cluster, fire and forget) to the CTA that will need
    // it as X^a * ACC.  The latency-bound remote loads of P1 become posted stores that drain behind P5.
    static constexpr bool ACC_TM = (ACCTM_ != 0);
    // PIPE (ACCTM_ == 2): k = 1 / two-level bootstraps of this plan run k_pbs_pipe (pbs_pipe.cuh): bulk-copied forward
    // exchange from local staging buffers, one mbarrier per buffer instead of cluster-wide barriers
    static constexpr bool PIPE = (ACCTM_ == 2);
    static constexpr int ACC_ITEMS = ACC_TM ? (2 << (LOGL1 - LOGC)) / NT_ : 0;     // (k+1)*Q / NT items per thread
    static constexpr int TM_ACC = ACC_TM ? ACC_ITEMS * 4 * (1 << LOGR1_) : 0;      // 2*R1 u64 per item
    static constexpr int TM_THREAD = TM_A + TM_B + TM_ACC;
    static constexpr int TM_NEED = (TM_THREAD > 0 ? TM_THREAD : 1) * ((NT_ + 127) / 128);
    static constexpr int TM_CTA = TM_NEED <= 32 ? 32 : TM_NEED <= 64 ? 64 : TM_NEED <= 128 ? 128 : TM_NEED <= 256 ? 256 : 512;
    static_assert(TM_NEED <= 512, "tensor memory holds 512 columns");
    static_assert(C_ == 1 || LOGR1_ == LOGC, "cluster plans use R1 == C");
    static_assert(LRA_ > 0 || LRB_ == 0, "pass B requires pass A");
    static_assert(!ACC_TM || (C_ > 1 && MINB_ == 1 && KMAX_ == 1 && ((2 << (LOGL1 - LOGC)) % NT_) == 0),
                  "ACC_TM: cluster plan, one CTA per SM, k = 1, whole items per thread");
};

struct PbsPlanMeta {       // what the host needs to know about a plan
    int C, logR1, lra, lrb, lrf, NT, minb, kmax, logM;
    int acctm;             // 1: accumulator in tensor memory + pushed rotated copy (PbsPlan::ACC_TM), 2: and k_pbs_pipe
    int group;             // > 1: k = 1 bootstraps run k_pbs_group with this many ciphertext slots per CTA (when they fit)
    long smem_fixed;       // bytes of shared memory besides ACC and the work buffers (twiddles live in TMEM)
};

// ---- small complex helpers -----------------------------------------------------------
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
// * exp(-+ i*pi/4): (1 -+ i)/sqrt2
template <bool INV>
__device__ __forceinline__ double2 mul_w8(double2 a) {
    const double h = 0.70710678118654752440;
    return INV ? make_double2((a.x - a.y) * h, (a.x + a.y) * h) : make_double2((a.x + a.y) * h, (a.y - a.x) * h);
}
// * exp(-+ 3*i*pi/4): (-1 -+ i)/sqrt2
template <bool INV>
__device__ __forceinline__ double2 mul_w8_3(double2 a) {
    const double h = 0.70710678118654752440;
    return INV ? make_double2(-(a.x + a.y) * h, (a.x - a.y) * h) : make_double2((a.y - a.x) * h, -(a.x + a.y) * h);
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
        double2 e[4] = {x[0], x[2], x[4], x[6]};
        double2 o[4] = {x[1], x[3], x[5], x[7]};
        Dft<4, INV>::run(e);
        Dft<4, INV>::run(o);
        const double2 o1 = mul_w8<INV>(o[1]), o2 = mul_mi<INV>(o[2]), o3 = mul_w8_3<INV>(o[3]);
        x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
    }
};
// 16 = 4 x 4:  r = 4a + b, m = c + 4d:  W16^{rm} = W4^{ac} * W16^{bc} * W4^{bd}
template <bool INV>
struct Dft<16, INV> {
    static __device__ __forceinline__ void run(double2* x) {
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos, sin(pi/8)
        double2 u[4][4];   // u[b][c]
#pragma unroll
        for (int b = 0; b < 4; b++) {
            double2 t[4] = {x[b], x[4 + b], x[8 + b], x[12 + b]};
            Dft<4, INV>::run(t);
#pragma unroll
            for (int c = 0; c < 4; c++) u[b][c] = t[c];
        }
        // twiddles W16^{bc} (conjugated for the inverse)
        const double2 w1 = make_double2(c1, INV ? s1 : -s1);     // W16^1
        const double2 w3 = make_double2(s1, INV ? c1 : -c1);     // W16^3
        const double2 w9 = make_double2(-c1, INV ? -s1 : s1);    // W16^9
        u[1][1] = cmul(u[1][1], w1);
        u[1][2] = mul_w8<INV>(u[1][2]);                          // W16^2
        u[1][3] = cmul(u[1][3], w3);
        u[2][1] = mul_w8<INV>(u[2][1]);                          // W16^2
        u[2][2] = mul_mi<INV>(u[2][2]);                          // W16^4
        u[2][3] = mul_w8_3<INV>(u[2][3]);                        // W16^6
        u[3][1] = cmul(u[3][1], w3);
        u[3][2] = mul_w8_3<INV>(u[3][2]);                        // W16^6
        u[3][3] = cmul(u[3][3], w9);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double2 t[4] = {u[0][c], u[1][c], u[2][c], u[3][c]};
            Dft<4, INV>::run(t);
#pragma unroll
            for (int d = 0; d < 4; d++) x[c + 4 * d] = t[d];
        }
    }
};

// ---- cluster helpers ---------------------------------------------------------------------
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

// explicit distributed-shared-memory accesses (32-bit shared::cluster addresses, no generic LD/ST)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, int rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ u64 ld_cluster_u64(unsigned addr) {
    u64 v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_u64(unsigned addr, u64 v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ double2 ld_cluster_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_f64x2(unsigned addr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
// mbarrier plumbing and bulk asynchronous copies between the shared memories of a cluster: the inverse
// exchange of the FFT is PUSHED by the copy engine in 4 KiB chunks (measured ~2-3x the throughput of
// element-wise remote accesses, profiles/r01_ubench_dsmem_bulk.log) and the receiver waits for its own data
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arm(u64* bar, unsigned tx_bytes) {      // one arrival + expected bytes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_to_cluster(unsigned dst_cluster_addr, const void* src, unsigned bytes,
                                                     unsigned mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(smem_u32(src)), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
// global -> this CTA's shared memory, completion (bytes) on a local mbarrier
__device__ __forceinline__ void bulk_copy_from_global(void* dst_smem, const void* src, unsigned bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// relaxed: the arrival only says "my READS of the spare buffers are done" (their values were consumed before
// the preceding __syncthreads), so no memory fence (MEMBAR.ALL.GPU) is needed
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// L2 prefetch of a contiguous global range (bytes multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// int32 -> double and double -> torus without the slow conversion unit (I2F / FRND / F2I.S64):
// magic-number arithmetic on the FP64 pipe.
__device__ __forceinline__ double i32_to_f64(int d) {
    return __hiloint2double(0x43300000, d ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}
__device__ __forceinline__ u64 f2torus(double x) {
    const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52
    const double r = (x * 5.42101086242752217e-20 + MAGIC) - MAGIC; // rint(x / 2^64), |x| < 2^115
    x = fma(-18446744073709551616.0, r, x);                         // exact: x in [-2^63, 2^63], an integer
    const double hs = fma(x, 2.3283064365386963e-10, MAGIC);        // x / 2^32 rounded to an integer (+ MAGIC)
    const int hi = __double2loint(hs);
    const double lo_d = fma(-4294967296.0, hs - MAGIC, x);          // exact, |lo| <= 2^31
    const int lo = __double2loint(lo_d + MAGIC);
    u64 res = ((u64)(unsigned)hi << 32) + (u64)(i64)lo;
    if (lo_d >= 2147483648.0) res += 1ULL << 32;                    // lo == +2^31 wrapped to -2^31
    return res;
}
// x * scale (scale = 2^-64 / M) -> torus word: the fractional part of the scaled value, read straight out of the mantissa
// of (frac + 1.5) in [1, 2].  5 FP64 + 3 integer operations instead of the 10 + 6 of f2torus; the result is rounded to a
// multiple of 2^12 (2^-52 of the torus), 2^28 times below the fp64 FFT's own error at these sizes.
__device__ __forceinline__ u64 f2torus_frac(double x, double scale) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double y = x * scale;
    const double f = y - ((y + MAGIC) - MAGIC);              // y - rint(y) in [-0.5, 0.5], exact
    const double t = f + 1.5;
    const unsigned hi = (unsigned)__double2hiint(t), lo = (unsigned)__double2loint(t);
    // mantissa << 12 (sign and exponent fall off the top), minus 2^63
    return ((u64)(((hi << 12) | (lo >> 20)) ^ 0x80000000u) << 32) | (u64)(lo << 12);
}

__device__ __forceinline__ u64 f2torus_ref(double x) {

    x -= 18446744073709551616.0 * rint(x * 5.42101086242752217e-20);
    return (u64)__double2ll_rn(x);
}
__device__ __forceinline__ int modswitch(u64 x, int logN) {  // -> Z_{2N}
    const int lg = logN + 1;
    u64 t = ((x >> (64 - lg - 1)) + 1) >> 1;
    return (int)(t & ((1ULL << lg) - 1));
}
// -> multiples of 2^theta in Z_{2N} (rounded): the rotation leaves the low theta bits of every index alone
__device__ __forceinline__ int modswitch_coarse(u64 x, int logN, int theta) {
    const int lg = logN + 1 - theta;
    u64 t = ((x >> (64 - lg - 1)) + 1) >> 1;
    return (int)((t & ((1ULL << lg) - 1)) << theta);
}

// ---- per-thread twiddle factors in tensor memory -------------------------------------------------
struct TwCtx {
    unsigned base;            // tensor-memory allocation base (for dealloc)
    unsigned tA, tB;          // this thread's mid-pass twiddles in tensor memory (TMEM_TW plans)
    unsigned tAcc;            // this thread's accumulator words in tensor memory (ACC_TM plans)
    unsigned tT;              // this thread's R1 stage-1 twiddles in tensor memory (k_pbs_group: one item per thread)
    const double2 *sA, *sB;   // mid-pass twiddle tables in shared memory (other plans)
    const double2* twT;       // stage-1 twiddles [R1][L1] in global memory (L1/L2 resident)
};

// twiddles of one mid pass for this thread's butterfly index, from the global table [(R-1)][S]
template <int LR, int LOGLC, int LOGBUF, int NT>
__device__ __forceinline__ void tw_init_pass(unsigned taddr, const double2* __restrict__ tw_g, int tix = -1) {
    constexpr int R = 1 << LR;
    constexpr int LOGSS = LOGLC - LR, SS = 1 << LOGSS;
    constexpr int NBF = 1 << (LOGBUF - LR);
    constexpr int TMW = (4 * (R - 1) <= 4) ? 4 : (4 * (R - 1) <= 16) ? 16 : (4 * (R - 1) <= 32) ? 32 : 64;
    const int tx = (tix < 0) ? (int)threadIdx.x : tix;
    const int w = (NBF >= NT) ? tx : (tx % NBF);
    const int q = w & (SS - 1);
    double2 t[TMW / 4];
#pragma unroll
    for (int m = 1; m <= TMW / 4; m++) t[m - 1] = (m < R) ? tw_g[((m - 1) << LOGSS) + q] : make_double2(0.0, 0.0);
    tmem_store_c<TMW / 4>(taddr, t);
}

// twiddle set-up: tensor-memory allocation + fill (TMEM_TW plans) or shared-memory tables; all threads call it
template <class P>
__device__ __forceinline__ TwCtx tw_setup(double2* tw_smem, const double2* __restrict__ twT,
                                          const double2* __restrict__ tw_mid) {
    TwCtx tw;
    tw.twT = twT;
    tw.sA = tw_smem;
    tw.sB = tw_smem + P::TWA;
    tw.base = tw.tA = tw.tB = tw.tAcc = tw.tT = 0;
    if constexpr (P::TMEM_TW) {
        __shared__ unsigned s_tmem_base;
        const int warp = threadIdx.x >> 5;
        if (warp == 0) tmem_alloc(&s_tmem_base, P::TM_CTA);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        tw.base = s_tmem_base;
        const unsigned mine = tw.base + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)((warp >> 2) * P::TM_THREAD);
        tw.tA = mine;
        tw.tB = mine + P::TM_A;
        tw.tAcc = mine + P::TM_A + P::TM_B;
        if constexpr (P::LRA > 0) tw_init_pass<P::LRA, P::LOGL1, P::LOGBUF, P::NT>(tw.tA, tw_mid);
        if constexpr (P::LRB > 0) tw_init_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, P::NT>(tw.tB, tw_mid + P::TWA);
        tmem_wait_st();
    } else {
        for (int t = threadIdx.x; t < P::TWN; t += P::NT) tw_smem[t] = tw_mid[t];
        __syncthreads();
    }
    return tw;
}
template <class P>
__device__ __forceinline__ void tw_teardown(const TwCtx& tw) {
    if constexpr (P::TMEM_TW) {
        tmem_fence_before_sync();
        __syncthreads();
        if ((threadIdx.x >> 5) == 0) tmem_dealloc(tw.base, P::TM_CTA);
    }
}

// ---- one mid pass over `nbuf` buffers of BUF points, sub-transform length 2^LOGLC ------------
// forward (DIF): y_m = DFT_R(x)_m * W_Lc^{q m};   inverse (DIT): mirror with conjugates.
// A thread keeps the R-1 twiddles of its butterfly index in registers across the buffers it handles.
template <int LR, int LOGLC, int LOGBUF, int NT, bool INV, bool TM, int TMQ = 0 /* > 0: tensor-memory twiddles are fetched TMQ at a time (128-register kernels) */>
__device__ __forceinline__ void mid_pass(double2* __restrict__ work, int nbuf, unsigned tw_taddr /* TMEM */,
                                         const double2* __restrict__ tw /* shared: [(R-1)][S] */,
                                         int bstep = 1 /* distance between the buffers, in buffers */,
                                         int tix = -1 /* thread index inside the NT threads running the pass (default: threadIdx.x) */) {
    constexpr int R = 1 << LR;
    constexpr int LOGSS = LOGLC - LR, SS = 1 << LOGSS;          // stride inside a sub-transform
    constexpr int NBF = 1 << (LOGBUF - LR);                     // butterflies per buffer
    constexpr int WPT = (NBF >= NT) ? NBF / NT : 1;             // butterfly indices per thread
    constexpr int GROUPS = (NBF >= NT) ? 1 : NT / NBF;          // thread groups sharing the buffers
    const int tx = (tix < 0) ? (int)threadIdx.x : tix;
    const int grp = (GROUPS > 1) ? (tx / NBF) : 0;
#pragma unroll 1
    for (int it = 0; it < WPT; it++) {
        const int w = (GROUPS > 1) ? (tx % NBF) : tx + it * NT;
        const int b = w >> LOGSS, q = w & (SS - 1);
        const int pos0 = (b << LOGLC) + q;
        // the R-1 twiddles W_Lc^{q m} of this thread's butterfly index live in its tensor-memory lane
        static_assert(WPT == 1, "twiddles are per-thread constants");
        constexpr int TMW = (4 * (R - 1) <= 4) ? 4 : (4 * (R - 1) <= 16) ? 16 : (4 * (R - 1) <= 32) ? 32 : 64;
        double2 t[TMW / 4 + 1];
        // tcgen05.ld is warp-collective: it may sit inside the buffer loop only when the thread groups are
        // whole warps (otherwise part of a warp would skip it -- that hung the k = 2, N = 256 plan)
        constexpr bool TM_IN_LOOP = TM && (NBF % 32 == 0);
        if (!TM) {
#pragma unroll
            for (int m = 1; m < R; m++) t[m] = tw[((m - 1) << LOGSS) + q];
        } else if (!TM_IN_LOOP)
            tmem_load_c<TMW / 4>(tw_taddr, t + 1);
#pragma unroll 1
        for (int buf = grp; buf < nbuf; buf += GROUPS) {
            double2* base = work + ((size_t)(buf * bstep) << LOGBUF);
            double2 x[R];
#pragma unroll
            for (int r = 0; r < R; r++) x[r] = base[SW(pos0 + (r << LOGSS))];
            if constexpr (TM_IN_LOOP && TMQ > 0) {
                // 128-register kernels: the twiddles arrive TMQ at a time (TMW / 4 of them next to the R data points spill)
                static_assert((TMW / 4) % TMQ == 0, "whole chunks");
                if (!INV) Dft<R, false>::run(x);
#pragma unroll
                for (int c0 = 0; c0 < TMW / 4; c0 += TMQ) {
                    double2 tq[TMQ];
                    tmem_load_c<TMQ>(tw_taddr + 4 * c0, tq);
#pragma unroll
                    for (int e = 0; e < TMQ; e++)
                        if (c0 + e + 1 < R) x[c0 + e + 1] = INV ? cmulc(x[c0 + e + 1], tq[e]) : cmul(x[c0 + e + 1], tq[e]);
                }
                if (INV) Dft<R, true>::run(x);
            } else if (!INV) {
                Dft<R, false>::run(x);
                // tensor-memory twiddles are re-read per buffer right where they are needed: a ~75-cycle load
                // is cheaper than 64 registers held across the butterfly
                if (TM_IN_LOOP) tmem_load_c<TMW / 4>(tw_taddr, t + 1);
#pragma unroll
                for (int m = 1; m < R; m++) x[m] = cmul(x[m], t[m]);
            } else {
                if (TM_IN_LOOP) tmem_load_c<TMW / 4>(tw_taddr, t + 1);
#pragma unroll
                for (int m = 1; m < R; m++) x[m] = cmulc(x[m], t[m]);
                Dft<R, true>::run(x);
            }
#pragma unroll
            for (int r = 0; r < R; r++) base[SW(pos0 + (r << LOGSS))] = x[r];
        }
    }
}

template <class P, bool INV>
__device__ __forceinline__ void mid_passes(double2* work, int nbuf, const TwCtx& tw) {
    constexpr int TMQ = (P::NT >= 512 && P::TMEM_TW) ? 4 : 0;      // 128-register plans fetch their twiddles 4 at a time
    if (!INV) {
        if constexpr (P::LRA > 0) {
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, P::NT, false, P::TMEM_TW, TMQ>(work, nbuf, tw.tA, tw.sA);
            __syncthreads();
        }
        if constexpr (P::LRB > 0) {
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, P::NT, false, P::TMEM_TW, TMQ>(work, nbuf, tw.tB, tw.sB);
            __syncthreads();
        }
    } else {
        if constexpr (P::LRB > 0) {
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, P::NT, true, P::TMEM_TW, TMQ>(work, nbuf, tw.tB, tw.sB);
            __syncthreads();
        }
        if constexpr (P::LRA > 0) {
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, P::NT, true, P::TMEM_TW, TMQ>(work, nbuf, tw.tA, tw.sA);
            __syncthreads();
        }
    }
}

// ---- stage 1 output: twist, radix-R1 butterfly, stage twiddle, store to the owner of t1 ---------
// re/im: folded inputs of one item (j1 < R1); T: exp(i*pi*j2*(1-4*t1)/N); dst: local address of
// element j2 of the destination buffer (C > 1) or of element (t1 = 0, j2) (C == 1).
template <class P>
__device__ __forceinline__ void stage1_emit(double2* dst_buf, int j2, const double* re, const double* im,
                                            const double2* T, const PbsGeom& g) {
    double2 z[P::R1];
#pragma unroll
    for (int j1 = 0; j1 < P::R1; j1++) {
        const double2 d = make_double2(re[j1], im[j1]);
        z[j1] = (j1 == 0) ? d : cmul(d, g.cj[j1]);
    }
    Dft<P::R1, false>::run(z);
    const unsigned a0 = smem_u32(dst_buf + SW(j2));
#pragma unroll
    for (int t1 = 0; t1 < P::R1; t1++) {
        const double2 v = cmul(z[t1], T[t1]);
        if (P::C > 1) st_cluster_f64x2(mapa_u32(a0, (g.dbg & 4) ? (int)cg::this_cluster().block_rank() : t1), v);
        else dst_buf[SW((t1 << P::LOGL1) + j2)] = v;
    }
}

// inverse of stage 1 for one item: gather sub-transform outputs, conj twiddle, inverse butterfly, untwist
template <class P>
__device__ __forceinline__ void stage1_gather(const double2* src_buf, int j2, const double2* T, const PbsGeom& g,
                                              double2* z /* [R1]: (coef j1*L1+j2, coef M+j1*L1+j2), unnormalised */) {
    double2 v[P::R1];
    const unsigned a0 = smem_u32(src_buf + SW(j2));
#pragma unroll
    for (int t1 = 0; t1 < P::R1; t1++) {
        if (P::C > 1) v[t1] = ld_cluster_f64x2(mapa_u32(a0, (g.dbg & 8) ? (int)cg::this_cluster().block_rank() : t1));
        else v[t1] = src_buf[SW((t1 << P::LOGL1) + j2)];
    }
#pragma unroll
    for (int t1 = 0; t1 < P::R1; t1++) v[t1] = cmulc(v[t1], T[t1]);
    Dft<P::R1, true>::run(v);
#pragma unroll
    for (int j1 = 0; j1 < P::R1; j1++) z[j1] = (j1 == 0) ? v[0] : cmulc(v[j1], g.cj[j1]);
}

// inverse of stage 1 when the sub-transform outputs were pushed into this CTA: chunk t1 of `recv` holds the
// Q values CTA t1 computed for this CTA's j2 range (still in the senders' swizzle)
template <class P>
__device__ __forceinline__ void stage1_gather_local(const double2* recv, int j2, const double2* T, const PbsGeom& g,
                                                    double2* z) {
    double2 v[P::R1];
    const int off = SW(j2) & (P::Q - 1);     // the senders stored element j2 at SW(j2); chunks are 8-aligned
#pragma unroll
    for (int t1 = 0; t1 < P::R1; t1++) v[t1] = recv[(t1 << P::LOGQ) + off];
#pragma unroll
    for (int t1 = 0; t1 < P::R1; t1++) v[t1] = cmulc(v[t1], T[t1]);
    Dft<P::R1, true>::run(v);
#pragma unroll
    for (int j1 = 0; j1 < P::R1; j1++) z[j1] = (j1 == 0) ? v[0] : cmulc(v[j1], g.cj[j1]);
}

template <class P, bool TMT = false /* the thread's only item: twiddles wait in its tensor-memory lane */>
__device__ __forceinline__ void load_T(double2* T, const TwCtx& tw, int j2) {
    if constexpr (TMT) {
        tmem_load_c<P::R1>(tw.tT, T);
    } else {
#pragma unroll
        for (int t1 = 0; t1 < P::R1; t1++) T[t1] = __ldg(&tw.twT[((size_t)t1 << P::LOGL1) + j2]);
    }
}

// slot (index inside one polynomial's CTA slice) -> coefficient index
template <class P>
__device__ __forceinline__ int slot_to_coef(int s, int rank) {
    const int q = s & (P::Q - 1);
    const int hj = s >> P::LOGQ;                 // h*R1 + j1
    return (hj << P::LOGL1) + (rank << P::LOGQ) + q;
}

// Fourier BSK layout: [i][j][o][rank][r][w]  (w = butterfly group of the fused pass, r < RF)
template <class P>
__device__ __forceinline__ size_t ggsw_slice(int j, int o, int k, int rank) {
    return (((size_t)(j * (k + 1) + o) * P::C + rank) << P::LOGBUF);
}

// ---- P1 of a CMUX for every item (c, q) of this CTA: diff = X^a * ACC - ACC, signed decomposition
// (closest representable, balanced digits), one stage-1 emit per level.  ST = u32 when the kept
// level*base_log bits fit 30 bits (halves the registers of the decomposition state).
// OWN_TM (k_pbs_group): the thread's own accumulator words wait in its tensor-memory lane (TwCtx::tAcc, one item per
// thread); the rotated operand still comes from the shared-memory copy
template <class P, typename ST, bool TMT = false, bool OWN_TM = false>
__device__ __forceinline__ void p1_items(const u64* __restrict__ acc, double2* __restrict__ work, int rank, int a,
                                         const PbsGeom& g, const TwCtx& tw,
                                         int tix = -1 /* thread index inside the NT threads of this ciphertext (default: threadIdx.x) */) {
    constexpr int C = P::C, R1 = P::R1, NT = P::NT, Q = P::Q, N = P::N;
    const int k = g.k, level = g.level, base_log = g.base_log;
    const ST mask = (ST)((1ULL << base_log) - 1);
    const int nonrep = 64 - level * base_log;
    const bool do_prof = (g.prof != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
    // small stage-1 radices mean many light items per thread: unroll them for instruction-level parallelism
    constexpr int ITEM_UNROLL = (R1 <= 2) ? 4 : (R1 <= 4) ? 2 : 1;
#pragma unroll ITEM_UNROLL
    for (int it = (tix < 0 ? (int)threadIdx.x : tix); it < ((k + 1) << P::LOGQ); it += NT) {
        const long long t_a = do_prof ? clock64() : 0;
        const int c = it >> P::LOGQ, q = it & (Q - 1);
        const int j2 = (rank << P::LOGQ) + q;
        const u64* accc = acc + ((size_t)c << P::LOGS);
        double2 T[R1];
        load_T<P, TMT>(T, tw, j2);
        ST st[2 * R1];
        {
            // coefficient hj*L1 + j2 of X^a * ACC is +-ACC[(hj*L1 + j2 - a) mod 2N]: the low bits (owner CTA
            // and slot offset) are the same for all hj of the item, only the block index hb + hj moves
            const int base = (j2 - a) & (2 * N - 1);
            const int jj2 = base & (P::L1 - 1);
            const int hb = base >> P::LOGL1;                      // in [0, 4*R1)
            const int owner = (g.dbg & 2) ? rank : (jj2 >> P::LOGQ);
            const u64* src_local = accc + (jj2 & (Q - 1));
            const unsigned src_peer = (C > 1 && !P::ACC_TM) ? mapa_u32(smem_u32(src_local), owner) : 0u;
#pragma unroll
            for (int half = 0; half < 2; half++) {       // two batches bound the registers in flight
                u64 rv[R1];
                unsigned ow[(P::ACC_TM || OWN_TM) ? 2 * R1 : 1];
                if constexpr (OWN_TM) TmemIo<2 * R1>::ld(tw.tAcc + (unsigned)(half * 2 * R1), ow);
                if constexpr (P::ACC_TM) {
                    // the rotated operand was pushed into `acc` by its producers (sign applied), the thread's own
                    // words wait in its tensor-memory lane: no remote access on this side of the exchange
                    TmemIo<2 * R1>::ld(tw.tAcc + (unsigned)((it / NT) * 4 * R1 + half * 2 * R1), ow);
#pragma unroll
                    for (int e = 0; e < R1; e++) rv[e] = accc[((half * R1 + e) << P::LOGQ) + q];
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int e = 0; e < R1; e++) {
                        const int hjp = (hb + half * R1 + e) & (2 * R1 - 1);
                        if (C > 1) rv[e] = ld_cluster_u64(src_peer + ((unsigned)hjp << (P::LOGQ + 3)));
                        else rv[e] = src_local[hjp << P::LOGQ];
                    }
                    if constexpr (OWN_TM) tmem_wait_ld();
                }
#pragma unroll
                for (int e = 0; e < R1; e++) {
                    const int hj = half * R1 + e;
                    const int hjr = (hb + hj) & (4 * R1 - 1);
                    u64 diff;
                    if constexpr (P::ACC_TM) {
                        diff = rv[e] - (((u64)ow[2 * e + 1] << 32) | (u64)ow[2 * e]);
                    } else {
                        u64 own;
                        if constexpr (OWN_TM) own = ((u64)ow[2 * e + 1] << 32) | (u64)ow[2 * e];
                        else own = accc[(hj << P::LOGQ) + q];
                        diff = ((hjr >> (P::LOGR1 + 1)) ? (0ULL - rv[e]) : rv[e]) - own;
                    }
                    if (sizeof(ST) == 4) {
                        // level*base_log <= 30: the state only needs bits [nonrep-1, 64) of diff
                        const unsigned top = (unsigned)(diff >> 32);
                        const unsigned sh = (unsigned)(nonrep - 33);
                        st[hj] = (ST)(((top >> sh) + 1u) >> 1);
                    } else
                        st[hj] = (ST)(((diff >> (nonrep - 1)) + 1) >> 1);      // closest representable
                }
            }
        }
        if (do_prof) {
            // touch the last state word so that the timestamp waits for the loads
            const long long t_b = clock64() + (long long)(st[2 * R1 - 1] & 0);
            g.prof[9] += t_b - t_a;
        }
        const long long t_c = do_prof ? clock64() : 0;
#pragma unroll 1
        for (int t = 0; t < level; t++) {
            const int lv = level - 1 - t;       // digits come out least significant level first
            double re[R1], im[R1];
#pragma unroll
            for (int hj = 0; hj < 2 * R1; hj++) {
                const ST res = st[hj] & mask;
                const ST s2 = st[hj] >> base_log;
                const ST carry = (((res - 1) | s2) & res) >> (base_log - 1);
                st[hj] = s2 + carry;
                double d;
                if (sizeof(ST) == 4) d = i32_to_f64((int)res - (int)(carry << base_log));
                else d = (double)((i64)res - (i64)((u64)carry << base_log));
                if (hj < R1) re[hj] = d;
                else im[hj - R1] = d;
            }
            stage1_emit<P>(work + ((size_t)(c * level + lv) << P::LOGBUF), j2, re, im, T, g);
        }
        if (do_prof) g.prof[10] += clock64() - t_c;
    }
}

// ---- the bootstrap kernel ------------------------------------------------------------------
// EXT (extended bootstrap, used by the wide-integer lookup, wide.py):
//  * bsk_ct_stride > 0: every ciphertext brings its own list of n Fourier GGSWs (bsk_ct_stride double2 apart) --
//    the blind rotation of vertical packing, where the GGSWs come out of circuit bootstrapping;
//  * g.theta / g.nout: one blind rotation evaluates up to 2^theta test polynomials interleaved in the accumulator
//    (coefficient i belongs to function i mod 2^theta) and extracts coefficients 0..nout-1 (out: [B][nout][kN+1]).
template <class P, bool EXT = false>
__global__ void __launch_bounds__(P::NT, P::MINB)
k_pbs(u64* __restrict__ out, const u64* __restrict__ in, long B, const double2* __restrict__ bskF_all,
      const u64* __restrict__ luts, const int* __restrict__ lut_idx, const PbsGeom g,
      const double2* __restrict__ twT, const double2* __restrict__ tw_mid, long long* prof, size_t bsk_ct_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, R1 = P::R1, NT = P::NT, RF = P::RF, BUF = P::BUF, Q = P::Q, S = P::S;
    constexpr int N = P::N;
    // optional phase timing (development aid, fhe_pbs_set_profile): thread 0 of CTA 0 accumulates
    // clock64() deltas per phase
    // (CTA b < 16 writes prof[16*b + slot]: CTA 0 is the reference, the others show the skew inside cluster 0)
    const bool do_prof = (prof != nullptr) && blockIdx.x < 16 && threadIdx.x == 0;
    long long t_last = do_prof ? clock64() : 0;
#define PBS_PROF(slot)                                   \
    if (do_prof) {                                       \
        const long long t_now = clock64();               \
        prof[16 * blockIdx.x + slot] += t_now - t_last;  \
        t_last = t_now;                                  \
    }
    constexpr int NG = BUF / RF;                    // butterfly groups of the fused pass
    const int k = g.k, level = g.level, J = g.J, base_log = g.base_log;
    u64* acc = reinterpret_cast<u64*>(smem_raw);                                   // [(k+1)][S]
    double2* work = reinterpret_cast<double2*>(acc + (size_t)(k + 1) * S);         // [J][BUF]
    const int tid = threadIdx.x;
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const size_t ggsw_stride = (size_t)J * (k + 1) * C * BUF;   // double2 per CMUX iteration
    const double tscale = g.inv_M * 5.42101086242752217e-20;    // 2^-64 / M (f2torus_frac)

    const TwCtx tw = tw_setup<P>(work + (size_t)J * BUF, twT, tw_mid);
    // cluster plans with at least k+1 spare work buffers (level >= 2): the inverse exchange is pushed with
    // bulk copies into the peers' spare buffers and counted on one mbarrier per output polynomial
    __shared__ __align__(8) u64 s_bar_inv[4];
    const bool push_inv = (C >= 4) && (J >= 2 * (k + 1)) && !(g.dbg & 32);   // C == 2: the TPC-pair path is fast enough
    unsigned inv_phase = 0;
    if (C > 1) {
        if (tid == 0) {
            for (int o = 0; o <= k; o++) mbar_init(&s_bar_inv[o], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    if (g.stagger > 0) {
        // all clusters run the same phase sequence; started together they would hit the L2 (GGSW
        // streaming) and the SM-to-SM network (stage-1 scatter) at the same instants
        const long long t_end = clock64() + (long long)(cluster_id % 64) * g.stagger;
        while (clock64() < t_end) __nanosleep(100);
    }

    for (long ct = cluster_id; ct < B; ct += nclusters) {
        const double2* bskF = EXT ? bskF_all + (size_t)ct * bsk_ct_stride : bskF_all;
        const u64* lwe = in + (size_t)ct * (g.n + 1);
        const u64* lut = luts + (size_t)(lut_idx ? lut_idx[ct] : 0) * (k + 1) * N;
        // ---- ACC = X^{-b~} * LUT
        int a_next = EXT ? modswitch_coarse(lwe[0], g.logN, g.theta) : modswitch(lwe[0], g.logN);
        {
            const int a0 = (2 * N - (EXT ? modswitch_coarse(lwe[g.n], g.logN, g.theta) : modswitch(lwe[g.n], g.logN))) & (2 * N - 1);
            // ACC_TM: the shared array holds X^{a_0} * ACC, the operand of the first CMUX, ACC itself goes to tensor memory
            const int a0s = P::ACC_TM ? ((a0 + a_next) & (2 * N - 1)) : a0;
            for (int w = tid; w < (k + 1) * S; w += NT) {
                const int c = w >> P::LOGS, s = w & (S - 1);
                const int v = (slot_to_coef<P>(s, rank) - a0s) & (2 * N - 1);
                const u64 x = lut[(size_t)c * N + (v & (N - 1))];
                acc[w] = (v >> (P::LOGM + 1)) ? (0ULL - x) : x;
            }
            if constexpr (P::ACC_TM) {
#pragma unroll 1
                for (int it = tid; it < ((k + 1) << P::LOGQ); it += NT) {
                    const int c = it >> P::LOGQ, q = it & (Q - 1);
                    unsigned ow[4 * R1];
#pragma unroll
                    for (int hj = 0; hj < 2 * R1; hj++) {
                        const int v = ((hj << P::LOGL1) + (rank << P::LOGQ) + q - a0) & (2 * N - 1);
                        u64 x = lut[(size_t)c * N + (v & (N - 1))];
                        if (v >> (P::LOGM + 1)) x = 0ULL - x;
                        ow[2 * hj] = (unsigned)x;
                        ow[2 * hj + 1] = (unsigned)(x >> 32);
                    }
                    TmemIo<4 * R1>::st(tw.tAcc + (unsigned)((it / NT) * 4 * R1), ow);
                }
                tmem_wait_st();
            }
        }
        csync<C>();

        for (int i = 0; i < g.n; i++) {
            const int a = a_next;
            a_next = (i + 1 < g.n) ? (EXT ? modswitch_coarse(lwe[i + 1], g.logN, g.theta) : modswitch(lwe[i + 1], g.logN)) : 0;
            // GGSW_{i+1} slices of this rank -> L2 one iteration ahead (HBM latency off the MAC's path)
            if (i + 1 < g.n && tid < J * (k + 1) && !(g.dbg & 16))
                prefetch_l2_bulk(bskF + (size_t)(i + 1) * ggsw_stride + (((size_t)tid * C + rank) << P::LOGBUF),
                                 (unsigned)(sizeof(double2) << P::LOGBUF));
            if (push_inv && tid == 0)
                for (int o = 0; o <= k; o++) mbar_arm(&s_bar_inv[o], (unsigned)(sizeof(double2) << P::LOGBUF));
            // ---- P1: rotate, subtract, decompose, fold+twist, radix-R1 stage, scatter to owners
            if (level * base_log <= 30)
                p1_items<P, unsigned>(acc, work, rank, a, g, tw);
            else
                p1_items<P, u64>(acc, work, rank, a, g, tw);
            PBS_PROF(0)
            csync<C>();
            PBS_PROF(1)
            // ---- P2: mid passes of all J buffers
            mid_passes<P, false>(work, J, tw);
            PBS_PROF(2)
            // ---- P3: last forward pass + multiply-accumulate against GGSW_i + first inverse pass
            if constexpr ((NG < NT) && (NT % NG == 0) && (P::KMAX + 1 <= NT / NG)) {
                // fewer butterfly groups than threads (N = 2048 with 256 threads): the thread groups split the k+1
                // outputs between them -- each thread transforms its inputs and accumulates ONE output polynomial
                // (half the GGSW loads and FMAs per thread instead of half the threads idle)
                const double2* gg = bskF + (size_t)i * ggsw_stride;
                const int w = tid % NG, o = tid / NG;
                const bool act = o <= k;
                const size_t o_stride = (size_t)P::C << P::LOGBUF;
                const double2* gp = gg + ((size_t)rank << P::LOGBUF) + w + (act ? o : 0) * o_stride;
                double2 r[RF], gv[RF];
#pragma unroll
                for (int m = 0; m < RF; m++) {
                    r[m] = make_double2(0.0, 0.0);
                    gv[m] = __ldg(gp + m * NG);
                }
#pragma unroll 1
                for (int j = 0; j < J; j++) {
                    double2 x[RF];
                    const double2* base = work + ((size_t)j << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) x[m] = base[SW(w * RF + m)];
                    Dft<RF, false>::run(x);
                    const double2* gn = gp + (size_t)(j + 1) * (k + 1) * o_stride;
                    const bool more = (j + 1 < J);
#pragma unroll
                    for (int m = 0; m < RF; m++) {
                        const double2 b = gv[m];
                        r[m].x = fma(x[m].x, b.x, r[m].x);
                        r[m].x = fma(-x[m].y, b.y, r[m].x);
                        r[m].y = fma(x[m].x, b.y, r[m].y);
                        r[m].y = fma(x[m].y, b.x, r[m].y);
                        if (more) gv[m] = __ldg(gn + m * NG);
                    }
                }
                __syncthreads();      // the other groups still read the buffers the outputs overwrite
                if (act) {
                    Dft<RF, true>::run(r);
                    double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) base[SW(w * RF + m)] = r[m];
                }
            } else {
                const double2* gg = bskF + (size_t)i * ggsw_stride;
#pragma unroll 1
                for (int w = tid; w < NG; w += NT) {
                    double2 r[P::KMAX + 1][RF];
#pragma unroll
                    for (int o = 0; o <= P::KMAX; o++)
#pragma unroll
                        for (int m = 0; m < RF; m++) r[o][m] = make_double2(0.0, 0.0);
                    // rolling prefetch: every GGSW register is re-loaded for the next input polynomial right
                    // after its last use, so (k+1)*RF 16-byte loads stay in flight behind the FMAs
                    double2 gv[P::KMAX + 1][RF];
                    const double2* gp = gg + ((size_t)rank << P::LOGBUF) + w;
                    const size_t o_stride = (size_t)P::C << P::LOGBUF;
#pragma unroll
                    for (int o = 0; o <= P::KMAX; o++)
                        if (o <= k) {
#pragma unroll
                            for (int m = 0; m < RF; m++) gv[o][m] = (g.dbg & 1) ? make_double2(1.0, 0.5) : __ldg(gp + o * o_stride + m * NG);
                        }
#pragma unroll 1
                    for (int j = 0; j < J; j++) {
                        double2 x[RF];
                        const double2* base = work + ((size_t)j << P::LOGBUF);
#pragma unroll
                        for (int m = 0; m < RF; m++) x[m] = base[SW(w * RF + m)];
                        Dft<RF, false>::run(x);
                        const double2* gn = gp + (size_t)(j + 1) * (k + 1) * o_stride;
                        const bool more = (j + 1 < J);
#pragma unroll
                        for (int o = 0; o <= P::KMAX; o++)
                            if (o <= k) {
#pragma unroll
                                for (int m = 0; m < RF; m++) {
                                    const double2 b = gv[o][m];
                                    r[o][m].x = fma(x[m].x, b.x, r[o][m].x);
                                    r[o][m].x = fma(-x[m].y, b.y, r[o][m].x);
                                    r[o][m].y = fma(x[m].x, b.y, r[o][m].y);
                                    r[o][m].y = fma(x[m].y, b.x, r[o][m].y);
                                    if (more && !(g.dbg & 1)) gv[o][m] = __ldg(gn + o * o_stride + m * NG);
                                }
                            }
                    }
#pragma unroll
                    for (int o = 0; o <= P::KMAX; o++)
                        if (o <= k) {
                            Dft<RF, true>::run(r[o]);
                            double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                            for (int m = 0; m < RF; m++) base[SW(w * RF + m)] = r[o][m];
                        }
                }
            }
            __syncthreads();
            if (push_inv) cluster_arrive();      // "my spare buffers are free": completes behind the inverse passes
            PBS_PROF(3)
            // ---- P4: inverse mid passes of the k+1 output buffers
            mid_passes<P, true>(work, k + 1, tw);
            PBS_PROF(4)
            if (push_inv) {
                cluster_wait();
                if (tid < (k + 1) * C) {
                    // chunk [d*Q, (d+1)*Q) of my output buffer o -> chunk `rank` of CTA d's spare buffer k+1+o
                    const int o = tid / C, d = tid % C;
                    fence_proxy_async();
                    const double2* src = work + ((size_t)o << P::LOGBUF) + (d << P::LOGQ);
                    const double2* dst = work + ((size_t)(k + 1 + o) << P::LOGBUF) + (rank << P::LOGQ);
                    bulk_copy_to_cluster(mapa_u32(smem_u32(dst), d), src, (unsigned)(sizeof(double2) << P::LOGQ),
                                         mapa_u32(smem_u32(&s_bar_inv[o]), d));
                }
            } else if (C > 1)
                csync<C>();
            PBS_PROF(5)
            // ---- P5: gather from owners, inverse radix-R1 stage, untwist, round, ACC +=
            int o_ready = -1;
            for (int it = tid; it < ((k + 1) << P::LOGQ); it += NT) {
                const int o = it >> P::LOGQ, q = it & (Q - 1);
                const int j2 = (rank << P::LOGQ) + q;
                double2 T[R1];
                load_T<P>(T, tw, j2);
                double2 z[R1];
                if (push_inv) {
                    if (o != o_ready) {              // wait for the C chunks of output o (the others still fly)
                        mbar_wait(&s_bar_inv[o], inv_phase & 1);
                        o_ready = o;
                    }
                    stage1_gather_local<P>(work + ((size_t)(k + 1 + o) << P::LOGBUF), j2, T, g, z);
                } else
                    stage1_gather<P>(work + ((size_t)o << P::LOGBUF), j2, T, g, z);
                u64* acco = acc + ((size_t)o << P::LOGS);
                if constexpr (P::ACC_TM) {
                    // own words: tensor memory; every updated word is pushed to the CTA that reads it as X^{a'} * ACC in
                    // the next CMUX: coefficient p = hj*L1 + j2 lands on (p + a') mod N, negated when it wraps once
                    const bool push = (i + 1 < g.n);
                    const int u = j2 + (a_next & (P::L1 - 1));
                    const int j2d = u & (P::L1 - 1);
                    const int hbd = (a_next >> P::LOGL1) + (u >> P::LOGL1);
                    const unsigned dst = mapa_u32(smem_u32(acco + (j2d & (Q - 1))), j2d >> P::LOGQ);
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        unsigned ow[2 * R1];
                        const unsigned ta = tw.tAcc + (unsigned)((it / NT) * 4 * R1 + half * 2 * R1);
                        TmemIo<2 * R1>::ld(ta, ow);
                        tmem_wait_ld();
#pragma unroll
                        for (int j1 = 0; j1 < R1; j1++) {
                            const u64 v = (((u64)ow[2 * j1 + 1] << 32) | (u64)ow[2 * j1]) +
                                          f2torus_frac(half ? z[j1].y : z[j1].x, tscale);
                            ow[2 * j1] = (unsigned)v;
                            ow[2 * j1 + 1] = (unsigned)(v >> 32);
                            if (push) {
                                const int bt = (half * R1 + j1 + hbd) & (4 * R1 - 1);
                                st_cluster_u64(dst + ((unsigned)(bt & (2 * R1 - 1)) << (P::LOGQ + 3)),
                                               (bt >> (P::LOGR1 + 1)) ? (0ULL - v) : v);
                            }
                        }
                        TmemIo<2 * R1>::st(ta, ow);
                    }
                } else {
#pragma unroll
                    for (int j1 = 0; j1 < R1; j1++) {
                        acco[(j1 << P::LOGQ) + q] += f2torus_frac(z[j1].x, tscale);
                        acco[((R1 + j1) << P::LOGQ) + q] += f2torus_frac(z[j1].y, tscale);
                    }
                }
            }
            if constexpr (P::ACC_TM) tmem_wait_st();
            inv_phase++;
            PBS_PROF(6)
            csync<C>();
            PBS_PROF(7)
        }
        // ---- sample extract coefficient 0 (extended: coefficients 0..nout-1)
        if constexpr (P::ACC_TM) {
            // the accumulator words are in the owners' tensor-memory lanes
            const int nout = EXT ? g.nout : 1;
#pragma unroll 1
            for (int it = tid; it < ((k + 1) << P::LOGQ); it += NT) {
                const int c = it >> P::LOGQ, q = it & (Q - 1);
                unsigned ow[4 * R1];
                TmemIo<4 * R1>::ld(tw.tAcc + (unsigned)((it / NT) * 4 * R1), ow);
                tmem_wait_ld();
                for (int u = 0; u < nout; u++) {
                    u64* o = out + ((size_t)ct * nout + u) * ((size_t)k * N + 1);
#pragma unroll
                    for (int hj = 0; hj < 2 * R1; hj++) {
                        const int j = (hj << P::LOGL1) + (rank << P::LOGQ) + q;
                        const u64 v = ((u64)ow[2 * hj + 1] << 32) | (u64)ow[2 * hj];
                        if (c < k) {
                            if (j <= u) o[(size_t)c * N + (u - j)] = v;
                            else o[(size_t)c * N + (N + u - j)] = 0ULL - v;
                        } else if (j == u)
                            o[(size_t)k * N] = v;
                    }
                }
            }
        } else if (!EXT) {
            u64* o = out + (size_t)ct * ((size_t)k * N + 1);
            for (int w = tid; w < (k + 1) * S; w += NT) {
                const int c = w >> P::LOGS, s = w & (S - 1);
                const int j = slot_to_coef<P>(s, rank);
                const u64 v = acc[w];
                if (c < k) {
                    if (j == 0) o[(size_t)c * N] = v;
                    else o[(size_t)c * N + (N - j)] = 0ULL - v;
                } else if (j == 0)
                    o[(size_t)k * N] = v;
            }
        } else {
            const int nout = g.nout;
            for (int u = 0; u < nout; u++) {
                // coefficient u of the product: mask word t of polynomial c is A_c[u - t] (t <= u), -A_c[N + u - t] (t > u)
                u64* o = out + ((size_t)ct * nout + u) * ((size_t)k * N + 1);
                for (int w = tid; w < (k + 1) * S; w += NT) {
                    const int c = w >> P::LOGS, s = w & (S - 1);
                    const int j = slot_to_coef<P>(s, rank);
                    const u64 v = acc[w];
                    if (c < k) {
                        if (j <= u) o[(size_t)c * N + (u - j)] = v;
                        else o[(size_t)c * N + (N + u - j)] = 0ULL - v;
                    } else if (j == u)
                        o[(size_t)k * N] = v;
                }
            }
        }
        // the same thread re-initialises the same ACC words at the top of the loop
        PBS_PROF(8)
    }
    tw_teardown<P>(tw);
#undef PBS_PROF
}

// forward transform of one polynomial held as folded (re, im) inputs by the caller's functor:
// stage 1 for every item of this CTA, cluster sync, mid passes.  Leaves buffer `buf` ready for the
// last radix-RF pass.
template <class P, class Src>
__device__ __forceinline__ void forward_poly(double2* work, int buf, int rank, const PbsGeom& g,
                                             const TwCtx& tw, Src src) {
#pragma unroll 1
    for (int q = threadIdx.x; q < P::Q; q += P::NT) {
        const int j2 = (rank << P::LOGQ) + q;
        double2 T[P::R1];
        load_T<P>(T, tw, j2);
        double re[P::R1], im[P::R1];
#pragma unroll
        for (int j1 = 0; j1 < P::R1; j1++) {
            re[j1] = src((j1 << P::LOGL1) + j2);
            im[j1] = src(P::M + (j1 << P::LOGL1) + j2);
        }
        stage1_emit<P>(work + ((size_t)buf << P::LOGBUF), j2, re, im, T, g);
    }
}

// ---- standard-domain BSK -> Fourier BSK (same transform, same layout as the MAC reads) --
// in : [n][level][k+1 rows r][k+1 polys o][N] u64
// out: [n][J = r*level+lv][o][C ranks][RF][NG] double2
template <class P>
__global__ void __launch_bounds__(P::NT, P::MINB)
k_bsk_fourier(double2* __restrict__ outF, const u64* __restrict__ bsk, long npoly, const PbsGeom g,
              const double2* __restrict__ twT, const double2* __restrict__ tw_mid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, NT = P::NT, RF = P::RF, BUF = P::BUF, N = P::N;
    constexpr int NG = BUF / RF;
    double2* work = reinterpret_cast<double2*>(smem_raw);   // [BUF]
    const int k = g.k, level = g.level;
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const TwCtx tw = tw_setup<P>(work + BUF, twT, tw_mid);
    csync<C>();
    for (long p = cluster_id; p < npoly; p += nclusters) {
        const u64* src = bsk + (size_t)p * N;
        forward_poly<P>(work, 0, rank, g, tw, [&](int coef) { return (double)(i64)src[coef]; });
        csync<C>();
        mid_passes<P, false>(work, 1, tw);
        // p = ((i*level + lv)*(k+1) + r)*(k+1) + o
        const int o = (int)(p % (k + 1));
        const int r = (int)((p / (k + 1)) % (k + 1));
        const int lv = (int)((p / ((long)(k + 1) * (k + 1))) % level);
        const long i = p / ((long)(k + 1) * (k + 1) * level);
        const int j = r * level + lv;
        double2* dst = outF + (size_t)i * ((size_t)g.J * (k + 1) * C * BUF) + ggsw_slice<P>(j, o, k, rank);
        for (int w = threadIdx.x; w < NG; w += NT) {
            double2 x[RF];
#pragma unroll
            for (int m = 0; m < RF; m++) x[m] = work[SW(w * RF + m)];
            Dft<RF, false>::run(x);
#pragma unroll
            for (int m = 0; m < RF; m++) dst[m * NG + w] = x[m];
        }
        csync<C>();
    }
    tw_teardown<P>(tw);
}

// ---- debug / test entry: out = a_small * b_torus in Z_2^64[X]/(X^N+1) through the same fp64
// pipeline (forward both, pointwise product in the fused pass, inverse).  Used by tests to validate
// the transform against an exact integer negacyclic product.
template <class P>
__global__ void __launch_bounds__(P::NT, P::MINB)
k_polymul_test(u64* __restrict__ out, const i64* __restrict__ a_small, const u64* __restrict__ b_torus,
               long npoly, const PbsGeom g, const double2* __restrict__ twT,
               const double2* __restrict__ tw_mid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, R1 = P::R1, NT = P::NT, RF = P::RF, BUF = P::BUF, N = P::N, M = P::M, Q = P::Q;
    constexpr int NG = BUF / RF;
    double2* work = reinterpret_cast<double2*>(smem_raw);   // [2][BUF]
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const TwCtx tw = tw_setup<P>(work + 2 * BUF, twT, tw_mid);
    csync<C>();
    for (long p = cluster_id; p < npoly; p += nclusters) {
        const i64* pa = a_small + (size_t)p * N;
        const u64* pb = b_torus + (size_t)p * N;
        forward_poly<P>(work, 0, rank, g, tw, [&](int coef) { return (double)pa[coef]; });
        forward_poly<P>(work, 1, rank, g, tw, [&](int coef) { return (double)(i64)pb[coef]; });
        csync<C>();
        mid_passes<P, false>(work, 2, tw);
        for (int w = threadIdx.x; w < NG; w += NT) {
            double2 x[RF], y[RF];
#pragma unroll
            for (int m = 0; m < RF; m++) {
                x[m] = work[SW(w * RF + m)];
                y[m] = work[BUF + SW(w * RF + m)];
            }
            Dft<RF, false>::run(x);
            Dft<RF, false>::run(y);
#pragma unroll
            for (int m = 0; m < RF; m++) x[m] = cmul(x[m], y[m]);
            Dft<RF, true>::run(x);
#pragma unroll
            for (int m = 0; m < RF; m++) work[SW(w * RF + m)] = x[m];
        }
        __syncthreads();
        mid_passes<P, true>(work, 1, tw);
        csync<C>();
        for (int q = threadIdx.x; q < Q; q += NT) {
            const int j2 = (rank << P::LOGQ) + q;
            double2 T[R1];
            load_T<P>(T, tw, j2);
            double2 z[R1];
            stage1_gather<P>(work, j2, T, g, z);
#pragma unroll
            for (int j1 = 0; j1 < R1; j1++) {
                out[(size_t)p * N + (j1 << P::LOGL1) + j2] = f2torus(z[j1].x * g.inv_M);
                out[(size_t)p * N + M + (j1 << P::LOGL1) + j2] = f2torus(z[j1].y * g.inv_M);
            }
        }
        csync<C>();
    }
    tw_teardown<P>(tw);
}

// ---- batched CMUX on GLWE ciphertexts (selection tree of the wide-integer lookup) ------------------
//   out[b] = pool[i0[b]] + GGSW[gi[b]] (x) (pool[i1[b]] - pool[i0[b]])
// pool: [P][(k+1)N] u64, GGSWs in the Fourier layout of this plan (fhe_bsk_to_fourier with n = number of GGSWs).
// `plain`: the inputs are plaintext polynomials (trivial GLWEs, zero masks -- the leaves of the tree): only the
// body polynomial is decomposed and transformed.  The phases are the bootstrap's (stage 1 with fused
// decomposition, mid passes, fused MAC, inverse passes, gather); the difference is read from global memory
// instead of the rotated accumulator.
template <class P>
__global__ void __launch_bounds__(P::NT, P::MINB)
k_cmux(u64* __restrict__ out, const u64* __restrict__ pool, const int* __restrict__ i0, const int* __restrict__ i1,
       const double2* __restrict__ ggswF, const int* __restrict__ gi, long B, const PbsGeom g,
       const double2* __restrict__ twT, const double2* __restrict__ tw_mid, int plain) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, R1 = P::R1, NT = P::NT, RF = P::RF, BUF = P::BUF, N = P::N, M = P::M, Q = P::Q;
    constexpr int NG = BUF / RF;
    constexpr int MAXL = 4;                                  // get_plan() admits at most 4 levels
    const int k = g.k, level = g.level, J = g.J, base_log = g.base_log;
    double2* work = reinterpret_cast<double2*>(smem_raw);   // [J][BUF]
    int rank = 0;
    if (C > 1) rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const size_t glwe = (size_t)(k + 1) * N;
    const size_t ggsw_stride = (size_t)J * (k + 1) * C * BUF;
    const double tscale = g.inv_M * 5.42101086242752217e-20; // 2^-64 / M (f2torus_frac)
    const int c_first = plain ? k : 0;                      // first GLWE component with a non-zero difference
    const int j_first = c_first * level;
    const TwCtx tw = tw_setup<P>(work + (size_t)J * BUF, twT, tw_mid);
    csync<C>();
    for (long b = cluster_id; b < B; b += nclusters) {
        const u64* d0 = pool + (size_t)i0[b] * glwe;
        const u64* d1 = pool + (size_t)i1[b] * glwe;
        const double2* gg = ggswF + (size_t)gi[b] * ggsw_stride;
        // ---- decomposition of d1 - d0 fused into stage 1 (one item = the 2*R1 coefficients of a butterfly)
#pragma unroll 1
        for (int it = threadIdx.x; it < ((k + 1 - c_first) << P::LOGQ); it += NT) {
            const int c = c_first + (it >> P::LOGQ), q = it & (Q - 1);
            const int j2 = (rank << P::LOGQ) + q;
            double2 T[R1];
            load_T<P>(T, tw, j2);
            int dg[2 * R1][MAXL];
#pragma unroll
            for (int hj = 0; hj < 2 * R1; hj++) {
                const size_t coef = (size_t)c * N + (hj << P::LOGL1) + j2;
                signed_decompose<MAXL>(d1[coef] - d0[coef], base_log, level, dg[hj]);
            }
#pragma unroll
            for (int t = 0; t < MAXL; t++)
                if (t < level) {                              // dg[.][t] is the digit of level index level-1-t
                    double re[R1], im[R1];
#pragma unroll
                    for (int j1 = 0; j1 < R1; j1++) {
                        re[j1] = (double)dg[j1][t];
                        im[j1] = (double)dg[R1 + j1][t];
                    }
                    stage1_emit<P>(work + ((size_t)(c * level + level - 1 - t) << P::LOGBUF), j2, re, im, T, g);
                }
        }
        csync<C>();
        mid_passes<P, false>(work + ((size_t)j_first << P::LOGBUF), J - j_first, tw);
#pragma unroll 1
        for (int w = threadIdx.x; w < NG; w += NT) {
            double2 r[P::KMAX + 1][RF];
#pragma unroll
            for (int o = 0; o <= P::KMAX; o++)
#pragma unroll
                for (int m = 0; m < RF; m++) r[o][m] = make_double2(0.0, 0.0);
            const double2* gp = gg + ((size_t)rank << P::LOGBUF) + w;
            const size_t o_stride = (size_t)P::C << P::LOGBUF;
#pragma unroll 1
            for (int j = j_first; j < J; j++) {
                double2 x[RF];
                const double2* base = work + ((size_t)j << P::LOGBUF);
                const double2* gj = gp + (size_t)j * (k + 1) * o_stride;
                double2 gv[P::KMAX + 1][RF];
#pragma unroll
                for (int o = 0; o <= P::KMAX; o++)
                    if (o <= k) {
#pragma unroll
                        for (int m = 0; m < RF; m++) gv[o][m] = __ldg(gj + o * o_stride + m * NG);
                    }
#pragma unroll
                for (int m = 0; m < RF; m++) x[m] = base[SW(w * RF + m)];
                Dft<RF, false>::run(x);
#pragma unroll
                for (int o = 0; o <= P::KMAX; o++)
                    if (o <= k) {
#pragma unroll
                        for (int m = 0; m < RF; m++) {
                            const double2 v = gv[o][m];
                            r[o][m].x = fma(x[m].x, v.x, r[o][m].x);
                            r[o][m].x = fma(-x[m].y, v.y, r[o][m].x);
                            r[o][m].y = fma(x[m].x, v.y, r[o][m].y);
                            r[o][m].y = fma(x[m].y, v.x, r[o][m].y);
                        }
                    }
            }
            // outputs overwrite buffers 0..k at the positions this thread alone reads: no barrier needed
#pragma unroll
            for (int o = 0; o <= P::KMAX; o++)
                if (o <= k) {
                    Dft<RF, true>::run(r[o]);
                    double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) base[SW(w * RF + m)] = r[o][m];
                }
        }
        __syncthreads();
        mid_passes<P, true>(work, k + 1, tw);
        csync<C>();
#pragma unroll 1
        for (int it = threadIdx.x; it < ((k + 1) << P::LOGQ); it += NT) {
            const int o = it >> P::LOGQ, q = it & (Q - 1);
            const int j2 = (rank << P::LOGQ) + q;
            double2 T[R1];
            load_T<P>(T, tw, j2);
            double2 z[R1];
            stage1_gather<P>(work + ((size_t)o << P::LOGBUF), j2, T, g, z);
            const u64* src = d0 + (size_t)o * N;
            u64* dst = out + (size_t)b * glwe + (size_t)o * N;
#pragma unroll
            for (int j1 = 0; j1 < R1; j1++) {
                const int c0 = (j1 << P::LOGL1) + j2;
                dst[c0] = src[c0] + f2torus_frac(z[j1].x, tscale);
                dst[M + c0] = src[M + c0] + f2torus_frac(z[j1].y, tscale);
            }
        }
        csync<C>();
    }
    tw_teardown<P>(tw);
}

#include "pbs_pipe.cuh"
#include "pbs_group.cuh"

// ------------------------------------------------------------------ host launch helpers ----
template <typename Kern, typename... Args>
static int launch_clustered(Kern kern, int C, int NT, int tmem_cols, size_t smem, long work_items, cudaStream_t st,
                            Args... args) {
    FHE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    FHE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int nclusters = 0;
    if (C > 8) FHE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
        FHE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
        if (per_sm < 1) FHE_FAIL(4, "pbs: kernel does not fit on an SM (smem %zu)", smem);
        if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;      // tensor-memory columns per SM
        nclusters = per_sm * sms;
    }
    if ((long)nclusters > work_items) nclusters = (int)work_items;
    cfg.gridDim = dim3((unsigned)(nclusters * C));
    FHE_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

extern int g_pbs_pipeline;    // pbs.cu: k_pbs_pipe enabled (default 1, FHE_PBS_PIPE=0 / fhe_pbs_set_pipeline)
// pbs.cu: a zeroed device counter for the wave alignment of one k_pbs_pipe launch (null when FHE_PBS_WAVE=0)
int pbs_wave_counter(cudaStream_t st, unsigned** ctr);
extern int g_pbs_group;       // pbs.cu: k_pbs_group enabled (default 1, FHE_PBS_GROUP=0 / fhe_pbs_set_group)

// one CTA per SM, GROUP ciphertexts per CTA (k_pbs_group)
template <typename Kern, typename... Args>
static int launch_grouped(Kern kern, int threads, size_t smem, long B, cudaStream_t st, Args... args) {
    FHE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    FHE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.gridDim = dim3((unsigned)((long)sms < B ? (long)sms : B));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    FHE_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

struct PbsLaunchArgs {
    // common
    PbsGeom g;
    const double2* twT;
    const double2* tw_mid;
    cudaStream_t st;
    // bootstrap
    u64* out; const u64* in; long B; const double2* bskF; const u64* luts; const int* lut_idx; long long* prof;
    size_t bsk_ct_stride;      // > 0: one GGSW list per ciphertext (k_pbs<P, true>)
    // batched CMUX
    const u64* cm_pool; const int* cm_i0; const int* cm_i1; const int* cm_gi; int cm_plain;
    // bsk conversion
    double2* outF; const u64* bsk; long npoly;
    // polymul
    u64* pm_out; const i64* pm_a; const u64* pm_b;
};

template <class P>
struct PbsPlanOps {
    static PbsPlanMeta meta() {
        return PbsPlanMeta{P::C, P::LOGR1, P::LRA, P::LRB, P::LRF, P::NT, P::MINB, P::KMAX, P::LOGM, P::PIPE ? 2 : (P::ACC_TM ? 1 : 0),
                           P::GROUP, (long)sizeof(double2) * P::TW_SMEM};
    }
    static size_t smem_pbs(const PbsGeom& g) {
        return (size_t)(g.k + 1) * P::S * 8 + (size_t)g.J * P::BUF * 16 + (size_t)P::TW_SMEM * 16;
    }
    // ciphertexts (clusters) that can be resident at once for this geometry, or < 0 on error
    static int capacity(const PbsGeom& g) {
        const size_t smem = smem_pbs(g);
        if (cudaFuncSetAttribute(k_pbs<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (P::C > 1) {
            if (P::C > 8) cudaFuncSetAttribute(k_pbs<P>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(P::NT);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = P::C;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(P::C * (sms / P::C));
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_pbs<P>, &cfg) != cudaSuccess) return -1;
            return n;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pbs<P>, P::NT, smem) != cudaSuccess) return -1;
        if (per_sm > 512 / P::TM_CTA) per_sm = 512 / P::TM_CTA;
        return per_sm * sms;
    }
    static size_t smem_pipe() { return PipeStage<P>::SMEM; }
    static int pbs(const PbsLaunchArgs& a) {
        if constexpr (P::GROUP >= 1) {
            // fhe_pbs_set_group(0) / FHE_PBS_GROUP=0 run the plan's generic kernel instead (bit-identical outputs)
            using GP = PbsGroup<P, P::GROUP>;
            const bool ext = a.bsk_ct_stride || a.g.nout > 1 || a.g.theta > 0;
            // g_pbs_group: 1 = two slots per CTA, GGSW_i streamed from L2 (default: measured fastest); 2 = GGSW_i staged once
            // per SM in the shared-memory ring when it fits (north-star design, measured 4 % slower); 0 = k_pbs
            const bool ring = g_pbs_group == 2 && !ext && 2 * a.g.J <= GP::MAX_SLICES && GP::smem(a.g, true) + 2048 <= PBS_SMEM_MAX;
            if (g_pbs_group != 0 && a.g.k == 1 && GP::smem(a.g, ring) + 2048 <= PBS_SMEM_MAX) {
                if (ring)
                    return launch_grouped(k_pbs_group<P, P::GROUP, true, false>, GP::NTC, GP::smem(a.g, true), a.B, a.st, a.out,
                                          a.in, a.B, a.bskF, a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, (size_t)0);
                if (ext)
                    return launch_grouped(k_pbs_group<P, P::GROUP, false, true>, GP::NTC, GP::smem(a.g, false), a.B, a.st, a.out,
                                          a.in, a.B, a.bskF, a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, a.bsk_ct_stride);
                return launch_grouped(k_pbs_group<P, P::GROUP, false, false>, GP::NTC, GP::smem(a.g, false), a.B, a.st, a.out,
                                      a.in, a.B, a.bskF, a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, (size_t)0);
            }
        }
        if constexpr (P::PIPE) {
            // fhe_pbs_set_pipeline(0) / FHE_PBS_PIPE=0 run the plan's generic kernel instead: both kernels perform the same
            // floating-point operations in the same order, so their outputs must be BIT-IDENTICAL (the race test)
            const bool use_pipe = g_pbs_pipeline != 0;
            const bool plain = !(a.bsk_ct_stride || a.g.nout > 1 || a.g.theta > 0);
            const bool fits = use_pipe && plain && a.g.k == 1 && a.g.level == 2 && 2 * a.g.base_log <= 30 && a.g.base_log >= 2;
            if (fits) {
                unsigned* wave_ctr = nullptr;
                const int rc = pbs_wave_counter(a.st, &wave_ctr);
                if (rc) return rc;
                return launch_clustered(k_pbs_pipe<P>, P::C, P::NT, P::TM_CTA, smem_pipe(), a.B, a.st, a.out, a.in, a.B,
                                        a.bskF, a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, wave_ctr);
            }
        }
        if (a.bsk_ct_stride || a.g.nout > 1 || a.g.theta > 0)
            return launch_clustered(k_pbs<P, true>, P::C, P::NT, P::TM_CTA, smem_pbs(a.g), a.B, a.st, a.out, a.in, a.B,
                                    a.bskF, a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, a.bsk_ct_stride);
        return launch_clustered(k_pbs<P, false>, P::C, P::NT, P::TM_CTA, smem_pbs(a.g), a.B, a.st, a.out, a.in, a.B, a.bskF,
                                a.luts, a.lut_idx, a.g, a.twT, a.tw_mid, a.prof, (size_t)0);
    }
    static int cmux(const PbsLaunchArgs& a) {
        const size_t smem = (size_t)a.g.J * P::BUF * 16 + (size_t)P::TW_SMEM * 16;
        return launch_clustered(k_cmux<P>, P::C, P::NT, P::TM_CTA, smem, a.B, a.st, a.out, a.cm_pool, a.cm_i0, a.cm_i1,
                                a.bskF, a.cm_gi, a.B, a.g, a.twT, a.tw_mid, a.cm_plain);
    }
    static int bsk(const PbsLaunchArgs& a) {
        const size_t smem = (size_t)P::BUF * 16 + (size_t)P::TW_SMEM * 16;
        return launch_clustered(k_bsk_fourier<P>, P::C, P::NT, P::TM_CTA, smem, a.npoly, a.st, a.outF, a.bsk, a.npoly, a.g,
                                a.twT, a.tw_mid);
    }
    static int polymul(const PbsLaunchArgs& a) {
        const size_t smem = (size_t)2 * P::BUF * 16 + (size_t)P::TW_SMEM * 16;
        return launch_clustered(k_polymul_test<P>, P::C, P::NT, P::TM_CTA, smem, a.npoly, a.st, a.pm_out, a.pm_a, a.pm_b,
                                a.npoly, a.g, a.twT, a.tw_mid);
    }
};

struct PbsPlanEntry {
    PbsPlanMeta meta;
    int (*capacity)(const PbsGeom&);
    int (*pbs)(const PbsLaunchArgs&);
    int (*bsk)(const PbsLaunchArgs&);
    int (*polymul)(const PbsLaunchArgs&);
    int (*cmux)(const PbsLaunchArgs&);
};

#define PBS_PLAN_ENTRY(...) \
    PbsPlanEntry { PbsPlanOps<__VA_ARGS__>::meta(), &PbsPlanOps<__VA_ARGS__>::capacity, &PbsPlanOps<__VA_ARGS__>::pbs, &PbsPlanOps<__VA_ARGS__>::bsk, \
                   &PbsPlanOps<__VA_ARGS__>::polymul, &PbsPlanOps<__VA_ARGS__>::cmux }
