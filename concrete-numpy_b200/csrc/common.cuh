// Shared device/host helpers of the B200 TFHE backend (sm_100a only).
//
// Torus = u64 (q = 2^64).  Data layouts:
//   LWE ct  : d mask words then the body word                 [d+1] u64
//   GLWE ct : k mask polynomials then the body polynomial      [(k+1)][N] u64
//   BSK std : [n][level][k+1 rows][k+1 polys][N] u64, level lv <-> 2^(64-(lv+1)*base_log)
//   KSK     : [kN][level][n+1] u64
//   keys    : binary, one u64 (0/1) per bit
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

// ---- error plumbing (fhe_last_error() in api.cu) ---------------------------------
extern thread_local char g_fhe_err[512];
#define FHE_FAIL(code, ...)                                  \
    do {                                                     \
        snprintf(g_fhe_err, sizeof(g_fhe_err), __VA_ARGS__); \
        return (code);                                       \
    } while (0)
#define FHE_CUDA(expr)                                                                    \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            FHE_FAIL(1000 + (int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                 \
    } while (0)
#define FHE_CHECK_LAUNCH() FHE_CUDA(cudaGetLastError())

// Stream-ordered scratch memory for hot paths: the device's default pool keeps what is freed (release threshold = max)
// instead of handing it back to the OS at every synchronisation point, so cudaMallocAsync inside Server.run costs
// microseconds, not a driver allocation (measured: 0.3-20 ms per call with the default threshold of 0)
static inline cudaError_t fhe_malloc_async(void** p, size_t bytes, cudaStream_t st) {
    static bool prepared[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !prepared[dev]) {
        cudaMemPool_t pool;
        e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ULL;
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        prepared[dev] = true;
    }
    return cudaMallocAsync(p, bytes, st);
}

static inline int ilog2_exact(long v) {
    int l = 0;
    while ((1L << l) < v) l++;
    return ((1L << l) == v) ? l : -1;
}

// ---- counter-based ChaCha20 (same definition as oracle/tfhe_ref.c) ---------------
__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

struct ChaKey {
    u32 w[8];
};
// Key material of the C ABI: 16 host words = a SECRET 256-bit ChaCha key (secret-key bits, noise) followed by a
// separate 256-bit key for the PUBLIC masks of BSK / KSK / ciphertexts -- nothing published is a stream of the key
// that produced the secrets.  Production callers pass 64 bytes of OS entropy.
struct ChaKeys {
    ChaKey sec, pub;
};
static inline ChaKeys chacha_keys_load(const u32* key16) {
    ChaKeys k;
    for (int i = 0; i < 8; i++) { k.sec.w[i] = key16[i]; k.pub.w[i] = key16[8 + i]; }
    return k;
}
// deterministic expansion of a 64-bit TEST seed (fhe_seed_expand; same definition as oracle/tfhe_ref.c)
#define FHE_PUBLIC_SEED_TWEAK 0xA5A5C3C39696F00FULL
static inline void chacha_key_from_seed(u64 seed, u32 k[8]) {
    for (int i = 0; i < 4; i++) {
        u64 kw = splitmix64(seed + (u64)i * 0xD1342543DE82EF95ULL);
        k[2 * i] = (u32)kw;
        k[2 * i + 1] = (u32)(kw >> 32);
    }
}

__device__ __forceinline__ u32 rotl32(u32 v, int c) { return __funnelshift_l(v, v, c); }
#define CHA_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); \
    c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  \
    c += d; b ^= c; b = rotl32(b, 7);

// one 64-byte block = 8 torus words
__device__ __forceinline__ void chacha20_block(const ChaKey& key, u64 stream, u64 block, u64 out[8]) {
    u32 s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) s[4 + i] = key.w[i];
    s[12] = (u32)block; s[13] = (u32)(block >> 32);
    s[14] = (u32)stream; s[15] = (u32)(stream >> 32);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        CHA_QR(x[0], x[4], x[8], x[12]) CHA_QR(x[1], x[5], x[9], x[13])
        CHA_QR(x[2], x[6], x[10], x[14]) CHA_QR(x[3], x[7], x[11], x[15])
        CHA_QR(x[0], x[5], x[10], x[15]) CHA_QR(x[1], x[6], x[11], x[12])
        CHA_QR(x[2], x[7], x[8], x[13]) CHA_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
        out[i] = (u64)(x[2 * i] + s[2 * i]) | ((u64)(x[2 * i + 1] + s[2 * i + 1]) << 32);
}

__device__ __forceinline__ u64 rng_word(const ChaKey& key, u64 stream, u64 idx) {
    u64 blk[8];
    chacha20_block(key, stream, idx >> 3, blk);
    u64 r = blk[0];
#pragma unroll
    for (int i = 1; i < 8; i++) r = ((idx & 7) == (u64)i) ? blk[i] : r;
    return r;
}

enum {
    DOM_SK_SMALL = 1, DOM_SK_BIG = 2, DOM_BSK_MASK = 3, DOM_BSK_NOISE = 4,
    DOM_KSK_MASK = 5, DOM_KSK_NOISE = 6, DOM_ENC_MASK = 7, DOM_ENC_NOISE = 8
};
__host__ __device__ __forceinline__ u64 stream_id(int dom, u64 obj) { return ((u64)dom << 56) | obj; }

// ---- deterministic Gaussian: only correctly-rounded IEEE operations, written with
// explicit _rn intrinsics so nvcc never contracts a*b+c differently from the CPU port.
__device__ __forceinline__ double det_log(double x) {
    u64 bits = (u64)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7FF) - 1022;
    bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FE0000000000000ULL;
    double m = __longlong_as_double((long long)bits);
    if (m < 0.70710678118654752440) { m = __dmul_rn(m, 2.0); e -= 1; }
    double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    double s2 = __dmul_rn(s, s);
    double acc = 1.0 / 21.0;
    acc = __fma_rn(acc, s2, 1.0 / 19.0);
    acc = __fma_rn(acc, s2, 1.0 / 17.0);
    acc = __fma_rn(acc, s2, 1.0 / 15.0);
    acc = __fma_rn(acc, s2, 1.0 / 13.0);
    acc = __fma_rn(acc, s2, 1.0 / 11.0);
    acc = __fma_rn(acc, s2, 1.0 / 9.0);
    acc = __fma_rn(acc, s2, 1.0 / 7.0);
    acc = __fma_rn(acc, s2, 1.0 / 5.0);
    acc = __fma_rn(acc, s2, 1.0 / 3.0);
    acc = __fma_rn(acc, s2, 1.0);
    double lnm = __dmul_rn(2.0, __dmul_rn(s, acc));
    return __fma_rn((double)e, 0.69314718055994530942, lnm);
}
__device__ __forceinline__ double det_cos_small(double y) {
    double y2 = -__dmul_rn(y, y);
    double a = 1.0 / 6402373705728000.0;
    a = __fma_rn(a, y2, 1.0 / 20922789888000.0);
    a = __fma_rn(a, y2, 1.0 / 87178291200.0);
    a = __fma_rn(a, y2, 1.0 / 479001600.0);
    a = __fma_rn(a, y2, 1.0 / 3628800.0);
    a = __fma_rn(a, y2, 1.0 / 40320.0);
    a = __fma_rn(a, y2, 1.0 / 720.0);
    a = __fma_rn(a, y2, 1.0 / 24.0);
    a = __fma_rn(a, y2, 0.5);
    a = __fma_rn(a, y2, 1.0);
    return a;
}
__device__ __forceinline__ double det_sin_small(double y) {
    double y2 = -__dmul_rn(y, y);
    double a = 1.0 / 121645100408832000.0;
    a = __fma_rn(a, y2, 1.0 / 355687428096000.0);
    a = __fma_rn(a, y2, 1.0 / 1307674368000.0);
    a = __fma_rn(a, y2, 1.0 / 6227020800.0);
    a = __fma_rn(a, y2, 1.0 / 39916800.0);
    a = __fma_rn(a, y2, 1.0 / 362880.0);
    a = __fma_rn(a, y2, 1.0 / 5040.0);
    a = __fma_rn(a, y2, 1.0 / 120.0);
    a = __fma_rn(a, y2, 1.0 / 6.0);
    a = __fma_rn(a, y2, 1.0);
    return __dmul_rn(y, a);
}
__device__ __forceinline__ double det_cos2pi(double u) {
    double t = __dmul_rn(u, 4.0);
    int q = (int)t;
    double f = __dadd_rn(t, -(double)q);
    const double HALF_PI = 1.57079632679489661923;
    double c, s;
    if (f <= 0.5) {
        double y = __dmul_rn(f, HALF_PI);
        c = det_cos_small(y); s = det_sin_small(y);
    } else {
        double y = __dmul_rn(__dadd_rn(1.0, -f), HALF_PI);
        c = det_sin_small(y); s = det_cos_small(y);
    }
    switch (q & 3) {
        case 0: return c;
        case 1: return -s;
        case 2: return -c;
        default: return s;
    }
}
// j-th Gaussian sample of a noise stream as a wrapped torus value; std relative to 2^64
__device__ __forceinline__ u64 gaussian_torus(const ChaKey& key, u64 stream, u64 j, double std) {
    u64 blk[8];
    chacha20_block(key, stream, (2 * j) >> 3, blk);   // words 2j and 2j+1 share a block
    u64 x1 = blk[0], x2 = blk[1];
    int w = (int)((2 * j) & 7);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        x1 = (w == i) ? blk[i] : x1;
        x2 = (w == i) ? blk[i + 1] : x2;
    }
    double u1 = __dmul_rn((double)((x1 >> 11) + 1), 0x1.0p-53);
    double u2 = __dmul_rn((double)(x2 >> 11), 0x1.0p-53);
    double r = __dsqrt_rn(__dmul_rn(-2.0, det_log(u1)));
    double z = __dmul_rn(r, det_cos2pi(u2));
    double e = rint(__dmul_rn(z, __dmul_rn(std, 0x1.0p64)));
    return (u64)(i64)__double2ll_rn(e);
}

// ---- signed gadget decomposition: closest representable, balanced digits,
// produced from level `level` down to 1.  digits_desc[t] is the digit of level (level - t),
// t = 0..level-1 (compile-time indices keep the array in registers).
template <int MAXL>
__device__ __forceinline__ void signed_decompose(u64 a, int base_log, int level, int* digits_desc) {
    const int nonrep = 64 - level * base_log;
    const u64 mask = (1ULL << base_log) - 1;
    u64 state = ((a >> (nonrep - 1)) + 1) >> 1;
#pragma unroll
    for (int t = 0; t < MAXL; t++) {
        if (t < level) {
            u64 res = state & mask;
            state >>= base_log;
            u64 carry = ((res - 1) | state) & res;
            carry >>= (base_log - 1);
            state += carry;
            digits_desc[t] = (int)((i64)res - (i64)(carry << base_log));
        }
    }
}
