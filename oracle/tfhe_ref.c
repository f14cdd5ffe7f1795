/*
 * oracle/tfhe_ref.c -- CPU restatement of the TFHE arithmetic that sits behind
 * concrete-numpy's `Server.run` (reference: concrete/numpy/compilation/server.py:270-286,
 * `self._support.server_call(...)`), `Client.keygen/encrypt/decrypt`
 * (concrete/numpy/compilation/client.py:96-106, 108-182, 184-256).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (concrete-numpy_b200/) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and the
 * `cpu_baseline` / `--impl reference` legs of bench.py use it, as the checker / the
 * CPU arm.
 *
 * PARITY UNPINNED at ciphertext level: the arithmetic of this path lives in the
 * third-party wheel `concrete-compiler==0.24.0rc5` (reference pyproject.toml:47,
 * poetry.lock:484-496), which is absent from /root/reference and not installable
 * here (cp37-cp310 wheels only).  The reference holds no ciphertext-level golden
 * vectors.  This file therefore restates the *published*
 * TFHE algorithms as implemented by concrete-core / concrete-cpu at that version
 *  and is pinned against the reference only at the
 * decrypted-output level (tests/golden/, generated from the reference's own
 * frontend and clear evaluator).
 *
 * Conventions:
 *   torus = u64 (q = 2^64); LWE ct = d mask words then the body (d+1 words);
 *   GLWE ct = k mask polynomials then the body polynomial, N words each;
 *   secret keys are binary, stored one u64 (0/1) per bit;
 *   BSK (standard domain): [n][level][k+1 rows][k+1 polys][N], level index lv
 *   stands for gadget factor 2^(64-(lv+1)*base_log);
 *   KSK: [kN][level][n+1], same level convention.
 *
 * Randomness: a counter-based ChaCha20 generator (key, stream, index) so that the
 * CUDA path can reproduce every key / mask / noise word bit-for-bit.  Key material is
 * 16 words: a SECRET 256-bit ChaCha key (secret-key bits, noise) followed by a separate
 * 256-bit key for the PUBLIC masks of BSK / KSK / ciphertexts, so that nothing published
 * is a stream of the key that produced the secrets.  Production callers pass 64 bytes of
 * OS entropy; `ref_seed_expand` is the deterministic 64-bit-seed expansion of the tests.  Gaussian
 * noise is Box-Muller evaluated with +,-,*,/,sqrt,fma only (all correctly rounded
 * IEEE-754 operations) so CPU and GPU agree to the last bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

/* ------------------------------------------------------------------ RNG --- */

static inline u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

#define ROTL32(v, c) (((v) << (c)) | ((v) >> (32 - (c))))
#define QR(a, b, c, d)                                                        \
    a += b; d ^= a; d = ROTL32(d, 16);                                        \
    c += d; b ^= c; b = ROTL32(b, 12);                                        \
    a += b; d ^= a; d = ROTL32(d, 8);                                         \
    c += d; b ^= c; b = ROTL32(b, 7);

/* deterministic expansion of a 64-bit TEST seed into one 256-bit ChaCha key */
static void key_from_seed(u64 seed, u32 k[8]) {
    for (int i = 0; i < 4; i++) {
        u64 kw = splitmix64(seed + (u64)i * 0xD1342543DE82EF95ULL);
        k[2 * i] = (u32)kw;
        k[2 * i + 1] = (u32)(kw >> 32);
    }
}
#define PUBLIC_SEED_TWEAK 0xA5A5C3C39696F00FULL
/* test / reproducibility mode only: 64-bit seed -> (secret key, public-mask key) */
void ref_seed_expand(u64 seed, u32 key16[16]) {
    key_from_seed(seed, key16);
    key_from_seed(seed ^ PUBLIC_SEED_TWEAK, key16 + 8);
}
#define KSEC(key16) (key16)
#define KPUB(key16) ((key16) + 8)

/* One ChaCha20 block (RFC 8439 round function; 64-bit block counter in words
 * 12-13, 64-bit stream id in words 14-15). */
static void chacha20_block(const u32 key[8], u64 stream, u64 block, u32 out[16]) {
    u32 st[16], x[16];
    st[0] = 0x61707865u; st[1] = 0x3320646eu; st[2] = 0x79622d32u; st[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) st[4 + i] = key[i];
    st[12] = (u32)block; st[13] = (u32)(block >> 32);
    st[14] = (u32)stream; st[15] = (u32)(stream >> 32);
    memcpy(x, st, sizeof(x));
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

/* idx-th 64-bit word of stream `stream` under `key` (8 words per block). */
static inline u64 rng_u64(const u32 *key, u64 stream, u64 idx) {
    u32 blk[16];
    chacha20_block(key, stream, idx >> 3, blk);
    int w = (int)(idx & 7) * 2;
    return (u64)blk[w] | ((u64)blk[w + 1] << 32);
}

/* fill out[0..count) with words idx0.. of a stream, one block at a time */
static void rng_fill(const u32 *key, u64 stream, u64 idx0, u64 count, u64 *out) {
    u32 blk[16];
    u64 cur = ~0ULL;
    for (u64 j = 0; j < count; j++) {
        u64 idx = idx0 + j;
        if ((idx >> 3) != cur) { cur = idx >> 3; chacha20_block(key, stream, cur, blk); }
        int w = (int)(idx & 7) * 2;
        out[j] = (u64)blk[w] | ((u64)blk[w + 1] << 32);
    }
}

void ref_rng_u64(u64 seed, u64 stream, u64 idx0, u64 count, u64 *out) {
    u32 k[8];
    key_from_seed(seed, k);
    rng_fill(k, stream, idx0, count, out);
}

/* stream-id domains */
enum { DOM_SK_SMALL = 1, DOM_SK_BIG = 2, DOM_BSK_MASK = 3, DOM_BSK_NOISE = 4,
       DOM_KSK_MASK = 5, DOM_KSK_NOISE = 6, DOM_ENC_MASK = 7, DOM_ENC_NOISE = 8 };
#define STREAM(dom, obj) ((((u64)(dom)) << 56) | (u64)(obj))

/* ---- deterministic math: only correctly rounded IEEE ops (file is compiled with
 * -ffp-contract=off so a*b+c is never fused behind our back) ---- */

static double det_log(double x) { /* x in (0,1] (any positive normal works) */
    u64 bits; memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7FF) - 1022;               /* x = m * 2^e, m in [0.5,1) */
    bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FE0000000000000ULL;
    double m; memcpy(&m, &bits, 8);
    if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }  /* m in [sqrt(.5), sqrt(2)) */
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double acc = 1.0 / 21.0;
    acc = fma(acc, s2, 1.0 / 19.0);
    acc = fma(acc, s2, 1.0 / 17.0);
    acc = fma(acc, s2, 1.0 / 15.0);
    acc = fma(acc, s2, 1.0 / 13.0);
    acc = fma(acc, s2, 1.0 / 11.0);
    acc = fma(acc, s2, 1.0 / 9.0);
    acc = fma(acc, s2, 1.0 / 7.0);
    acc = fma(acc, s2, 1.0 / 5.0);
    acc = fma(acc, s2, 1.0 / 3.0);
    acc = fma(acc, s2, 1.0);
    double lnm = 2.0 * (s * acc);
    return fma((double)e, 0.69314718055994530942, lnm);
}

static double det_cos_small(double y) { /* y in [0, pi/4] */
    double y2 = y * y;
    double a = 1.0 / 6402373705728000.0;          /* 1/18! */
    a = fma(a, -y2, 1.0 / 20922789888000.0);      /* 1/16! */
    a = fma(a, -y2, 1.0 / 87178291200.0);         /* 1/14! */
    a = fma(a, -y2, 1.0 / 479001600.0);           /* 1/12! */
    a = fma(a, -y2, 1.0 / 3628800.0);             /* 1/10! */
    a = fma(a, -y2, 1.0 / 40320.0);               /* 1/8!  */
    a = fma(a, -y2, 1.0 / 720.0);                 /* 1/6!  */
    a = fma(a, -y2, 1.0 / 24.0);                  /* 1/4!  */
    a = fma(a, -y2, 0.5);                         /* 1/2!  */
    a = fma(a, -y2, 1.0);
    return a;
}

static double det_sin_small(double y) { /* y in [0, pi/4] */
    double y2 = y * y;
    double a = 1.0 / 121645100408832000.0;        /* 1/19! */
    a = fma(a, -y2, 1.0 / 355687428096000.0);     /* 1/17! */
    a = fma(a, -y2, 1.0 / 1307674368000.0);       /* 1/15! */
    a = fma(a, -y2, 1.0 / 6227020800.0);          /* 1/13! */
    a = fma(a, -y2, 1.0 / 39916800.0);            /* 1/11! */
    a = fma(a, -y2, 1.0 / 362880.0);              /* 1/9!  */
    a = fma(a, -y2, 1.0 / 5040.0);                /* 1/7!  */
    a = fma(a, -y2, 1.0 / 120.0);                 /* 1/5!  */
    a = fma(a, -y2, 1.0 / 6.0);                   /* 1/3!  */
    a = fma(a, -y2, 1.0);
    return y * a;
}

static double det_cos2pi(double u) { /* u in [0,1) -> cos(2*pi*u) */
    double t = u * 4.0;
    int q = (int)t;                 /* quadrant 0..3 */
    double f = t - (double)q;       /* exact */
    const double HALF_PI = 1.57079632679489661923;
    double c, s;                    /* cos, sin of (pi/2)*f */
    if (f <= 0.5) { double y = f * HALF_PI; c = det_cos_small(y); s = det_sin_small(y); }
    else { double y = (1.0 - f) * HALF_PI; c = det_sin_small(y); s = det_cos_small(y); }
    switch (q & 3) {
        case 0: return c;
        case 1: return -s;
        case 2: return -c;
        default: return s;
    }
}

/* j-th Gaussian sample of a noise stream, as a torus (u64, wrapping) value with
 * standard deviation `std` (relative to q = 2^64). */
static u64 gaussian_torus(const u32 *key, u64 stream, u64 j, double std) {
    u64 x1 = rng_u64(key, stream, 2 * j), x2 = rng_u64(key, stream, 2 * j + 1);
    double u1 = (double)((x1 >> 11) + 1) * 0x1.0p-53; /* (0,1] */
    double u2 = (double)(x2 >> 11) * 0x1.0p-53;       /* [0,1) */
    double r = sqrt(-2.0 * det_log(u1));
    double z = r * det_cos2pi(u2);
    double e = rint(z * (std * 0x1.0p64));
    return (u64)(i64)e;
}

void ref_gaussian(u64 seed, u64 stream, u64 j0, u64 count, double std, u64 *out) {
    u32 k[8];
    key_from_seed(seed, k);
    for (u64 j = 0; j < count; j++) out[j] = gaussian_torus(k, stream, j0 + j, std);
}

/* ------------------------------------------------------------- keygen ---- */

/* binary secret key: bit i = bit (i%64) of word i/64 of stream (domain, 0) */
void ref_keygen_lwe_secret(u64 *sk, int dim, const u32 *key16, int domain) {
    for (int i = 0; i < dim; i++)
        sk[i] = (rng_u64(KSEC(key16), STREAM(domain, 0), (u64)(i >> 6)) >> (i & 63)) & 1ULL;
}

/* out += a * s in Z[X]/(X^N+1), s binary (one u64 per coefficient) */
static void negacyclic_mul_binary_acc(u64 *out, const u64 *a, const u64 *s, int N) {
    for (int j = 0; j < N; j++) {
        if (!s[j]) continue;
        /* X^j * a : out[t] += a[t-j] (t>=j), out[t] -= a[N+t-j] (t<j) */
        for (int t = j; t < N; t++) out[t] += a[t - j];
        for (int t = 0; t < j; t++) out[t] -= a[N + t - j];
    }
}

/* GLWE encryption of zero, row index R selects the mask / noise streams */
static void glwe_encrypt_zero(u64 *ct, const u64 *sk_glwe, int k, int N, double std,
                              const u32 *key16, u64 R) {
    u64 *body = ct + (size_t)k * N;
    for (int t = 0; t < N; t++) body[t] = gaussian_torus(KSEC(key16), STREAM(DOM_BSK_NOISE, R), (u64)t, std);
    for (int c = 0; c < k; c++) {
        rng_fill(KPUB(key16), STREAM(DOM_BSK_MASK, R), (u64)c * N, (u64)N, ct + (size_t)c * N);
        negacyclic_mul_binary_acc(body, ct + (size_t)c * N, sk_glwe + (size_t)c * N, N);
    }
}

void ref_keygen_bsk(u64 *bsk, const u64 *sk_small, int n, const u64 *sk_glwe, int k, int N,
                    int level, int base_log, double std, const u32 *key16) {
    const size_t row_words = (size_t)(k + 1) * N;
    const long rows = (long)n * level * (k + 1);
#pragma omp parallel for schedule(dynamic, 4)
    for (long R = 0; R < rows; R++) {
        int r = (int)(R % (k + 1));
        int lv = (int)((R / (k + 1)) % level);
        int i = (int)(R / ((long)(k + 1) * level));
        u64 *row = bsk + (size_t)R * row_words;
        glwe_encrypt_zero(row, sk_glwe, k, N, std, key16, (u64)R);
        /* key entries are 0/1 for an LWE key; the extended key of the packing keyswitch (wide-integer
         * path, ref_keygen_bsk is reused for it) also holds -1 for the body slot */
        row[(size_t)r * N] += sk_small[i] * (1ULL << (64 - (lv + 1) * base_log));
    }
}

void ref_keygen_ksk(u64 *ksk, const u64 *sk_big, int kN, const u64 *sk_small, int n,
                    int level, int base_log, double std, const u32 *key16) {
    const long rows = (long)kN * level;
#pragma omp parallel for schedule(static)
    for (long R = 0; R < rows; R++) {
        int lv = (int)(R % level);
        int i = (int)(R / level);
        u64 *row = ksk + (size_t)R * (n + 1);
        rng_fill(KPUB(key16), STREAM(DOM_KSK_MASK, R), 0, (u64)n, row);
        u64 body = gaussian_torus(KSEC(key16), STREAM(DOM_KSK_NOISE, R), 0, std);
        for (int j = 0; j < n; j++) if (sk_small[j]) body += row[j];
        if (sk_big[i]) body += 1ULL << (64 - (lv + 1) * base_log);
        row[n] = body;
    }
}

/* ------------------------------------------------- encrypt / decrypt ------ */

void ref_encrypt_lwe_batch(u64 *ct, const u64 *pt, long B, const u64 *sk, int dim, double std,
                           const u32 *key16, u64 first_index) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < B; e++) {
        u64 *row = ct + (size_t)e * (dim + 1);
        u64 id = first_index + (u64)e;
        rng_fill(KPUB(key16), STREAM(DOM_ENC_MASK, id), 0, (u64)dim, row);
        u64 body = pt[e] + gaussian_torus(KSEC(key16), STREAM(DOM_ENC_NOISE, id), 0, std);
        for (int j = 0; j < dim; j++) if (sk[j]) body += row[j];
        row[dim] = body;
    }
}

void ref_trivial_lwe_batch(u64 *ct, const u64 *pt, long B, int dim) {
    for (long e = 0; e < B; e++) {
        u64 *row = ct + (size_t)e * (dim + 1);
        memset(row, 0, sizeof(u64) * dim);
        row[dim] = pt[e];
    }
}

/* phase = body - <a, s> */
void ref_decrypt_lwe_batch(u64 *phase, const u64 *ct, long B, const u64 *sk, int dim) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < B; e++) {
        const u64 *row = ct + (size_t)e * (dim + 1);
        u64 acc = row[dim];
        for (int j = 0; j < dim; j++) if (sk[j]) acc -= row[j];
        phase[e] = acc;
    }
}

/* message decoding of a phase at precision p (one padding bit): SURVEY A.2 */
void ref_decode_batch(u64 *msg, const u64 *phase, long B, int p) {
    for (long e = 0; e < B; e++) {
        u64 t = phase[e] >> (64 - p - 2);
        msg[e] = ((t >> 1) + (t & 1)) & ((1ULL << (p + 1)) - 1);
    }
}

/* ------------------------------------------------ signed decomposition ---- */

/* closest representable + balanced digits, levels produced from `level` down to 1
 * (SURVEY A.3).  digits[lv-1] receives the digit of level lv. */
static inline void signed_decompose(u64 a, int base_log, int level, i64 *digits) {
    const int nonrep = 64 - level * base_log;
    const u64 mask = (1ULL << base_log) - 1;
    u64 state = ((a >> (nonrep - 1)) + 1) >> 1;
    for (int lv = level; lv >= 1; lv--) {
        u64 res = state & mask;
        state >>= base_log;
        u64 carry = ((res - 1) | state) & res;
        carry >>= (base_log - 1);
        state += carry;
        digits[lv - 1] = (i64)res - (i64)(carry << base_log);
    }
}

void ref_signed_decompose(const u64 *a, long count, int base_log, int level, i64 *digits) {
    for (long e = 0; e < count; e++) signed_decompose(a[e], base_log, level, digits + e * level);
}

/* ------------------------------------------------------- keyswitch -------- */

void ref_keyswitch_batch(u64 *out, const u64 *in, long B, const u64 *ksk, int kN, int n,
                         int level, int base_log) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long e = 0; e < B; e++) {
        const u64 *ct = in + (size_t)e * (kN + 1);
        u64 *o = out + (size_t)e * (n + 1);
        memset(o, 0, sizeof(u64) * n);
        o[n] = ct[kN];
        i64 digits[64];
        for (int i = 0; i < kN; i++) {
            signed_decompose(ct[i], base_log, level, digits);
            for (int lv = 0; lv < level; lv++) {
                const u64 d = (u64)digits[lv];
                if (!d) continue;
                const u64 *row = ksk + ((size_t)i * level + lv) * (n + 1);
                for (int j = 0; j <= n; j++) o[j] -= d * row[j];
            }
        }
    }
}

/* ------------------------------------------------------------ FFT --------- */

typedef struct { double re, im; } cplx;

typedef struct {
    int N, M, logM;
    cplx *twist;   /* exp(i*pi*j/N), j < M */
    cplx *w;       /* exp(-2*pi*i*j/M), j < M/2 */
    int *rev;      /* bit reversal of M points */
} fft_plan;

static fft_plan *plan_cache[32];

static fft_plan *get_plan(int N) {
    int lg = 0; while ((1 << lg) < N) lg++;
    fft_plan *p;
#pragma omp critical(fftplan)
    {
        p = plan_cache[lg];
        if (!p) {
            p = (fft_plan *)malloc(sizeof(fft_plan));
            p->N = N; p->M = N / 2; p->logM = lg - 1;
            p->twist = (cplx *)malloc(sizeof(cplx) * p->M);
            p->w = (cplx *)malloc(sizeof(cplx) * (p->M / 2 + 1));
            p->rev = (int *)malloc(sizeof(int) * p->M);
            for (int j = 0; j < p->M; j++) {
                double a = M_PI * (double)j / (double)N;
                p->twist[j].re = cos(a); p->twist[j].im = sin(a);
            }
            for (int j = 0; j < p->M / 2; j++) {
                double a = -2.0 * M_PI * (double)j / (double)p->M;
                p->w[j].re = cos(a); p->w[j].im = sin(a);
            }
            for (int j = 0; j < p->M; j++) {
                int r = 0;
                for (int b = 0; b < p->logM; b++) if (j & (1 << b)) r |= 1 << (p->logM - 1 - b);
                p->rev[j] = r;
            }
            plan_cache[lg] = p;
        }
    }
    return p;
}

/* in-place radix-2 DIT, natural order in/out; inverse==1 uses conjugate twiddles */
static void fft_inplace(const fft_plan *p, cplx *x, int inverse) {
    const int M = p->M;
    for (int j = 0; j < M; j++) {
        int r = p->rev[j];
        if (r > j) { cplx t = x[j]; x[j] = x[r]; x[r] = t; }
    }
    for (int len = 2; len <= M; len <<= 1) {
        int half = len >> 1, step = M / len;
        for (int base = 0; base < M; base += len) {
            for (int j = 0; j < half; j++) {
                cplx w = p->w[j * step];
                if (inverse) w.im = -w.im;
                cplx a = x[base + j], b = x[base + j + half];
                double tr = b.re * w.re - b.im * w.im;
                double ti = b.re * w.im + b.im * w.re;
                x[base + j].re = a.re + tr; x[base + j].im = a.im + ti;
                x[base + j + half].re = a.re - tr; x[base + j + half].im = a.im - ti;
            }
        }
    }
}

/* forward negacyclic transform of a signed-integer polynomial given as doubles */
static void negacyclic_fwd(const fft_plan *p, const double *poly, cplx *out) {
    const int M = p->M;
    for (int j = 0; j < M; j++) {
        double a = poly[j], b = poly[j + M];
        out[j].re = a * p->twist[j].re - b * p->twist[j].im;
        out[j].im = a * p->twist[j].im + b * p->twist[j].re;
    }
    fft_inplace(p, out, 0);
}

/* inverse: Fourier -> wrapped torus polynomial, *added* into acc */
static void negacyclic_inv_add(const fft_plan *p, cplx *f, u64 *acc) {
    const int M = p->M;
    fft_inplace(p, f, 1);
    const double scale = 1.0 / (double)M;
    for (int j = 0; j < M; j++) {
        double zr = f[j].re * scale, zi = f[j].im * scale;
        double a = zr * p->twist[j].re + zi * p->twist[j].im;   /* multiply by conj(twist) */
        double b = zi * p->twist[j].re - zr * p->twist[j].im;
        a -= 0x1.0p64 * rint(a * 0x1.0p-64);
        b -= 0x1.0p64 * rint(b * 0x1.0p-64);
        acc[j] += (u64)(i64)rint(a);
        acc[j + M] += (u64)(i64)rint(b);
    }
}

/* Fourier BSK layout of the oracle: [n][level][k+1 rows][k+1 polys][M] cplx (natural order) */
void ref_bsk_to_fourier(double *bsk_f, const u64 *bsk, int n, int k, int N, int level) {
    const fft_plan *p = get_plan(N);
    const long polys = (long)n * level * (k + 1) * (k + 1);
#pragma omp parallel
    {
        double *tmp = (double *)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
        for (long q = 0; q < polys; q++) {
            const u64 *src = bsk + (size_t)q * N;
            for (int t = 0; t < N; t++) tmp[t] = (double)(i64)src[t];
            negacyclic_fwd(p, tmp, (cplx *)bsk_f + (size_t)q * (N / 2));
        }
        free(tmp);
    }
}

/* ------------------------------------------- programmable bootstrap ------- */

/* dst = X^a * src  (a in [0, 2N)), negacyclic */
static void monomial_mul(u64 *dst, const u64 *src, int a, int N) {
    for (int j = 0; j < N; j++) {
        int idx = j - a;                 /* in (-2N, N) */
        int neg = 0;
        while (idx < 0) { idx += N; neg ^= 1; }
        dst[j] = neg ? (u64)0 - src[idx] : src[idx];
    }
}

static inline int modswitch(u64 x, int log2N) { /* to Z_{2N} */
    int lg = log2N + 1;
    u64 t = ((x >> (64 - lg - 1)) + 1) >> 1;
    return (int)(t & ((1ULL << lg) - 1));
}

/* one CMUX: acc += ggsw_i (x) (X^a * acc - acc) */
static void cmux_rotate(const fft_plan *p, u64 *acc, int a, const cplx *ggsw_f, int k, int N,
                        int level, int base_log, u64 *diff, double *dig, cplx *fin, cplx *fout) {
    const int M = N / 2;
    for (int c = 0; c <= k; c++) {
        monomial_mul(diff + (size_t)c * N, acc + (size_t)c * N, a, N);
        for (int t = 0; t < N; t++) diff[(size_t)c * N + t] -= acc[(size_t)c * N + t];
    }
    memset(fout, 0, sizeof(cplx) * (size_t)(k + 1) * M);
    i64 digits[64];
    for (int c = 0; c <= k; c++) {
        /* dig holds `level` decomposed polynomials of component c */
        for (int t = 0; t < N; t++) {
            signed_decompose(diff[(size_t)c * N + t], base_log, level, digits);
            for (int lv = 0; lv < level; lv++) dig[(size_t)lv * N + t] = (double)digits[lv];
        }
        for (int lv = 0; lv < level; lv++) {
            negacyclic_fwd(p, dig + (size_t)lv * N, fin);
            /* row (lv, c): k+1 Fourier polynomials */
            const cplx *row = ggsw_f + ((size_t)lv * (k + 1) + c) * (k + 1) * M;
            for (int o = 0; o <= k; o++) {
                const cplx *g = row + (size_t)o * M;
                cplx *dst = fout + (size_t)o * M;
                for (int t = 0; t < M; t++) {
                    dst[t].re += fin[t].re * g[t].re - fin[t].im * g[t].im;
                    dst[t].im += fin[t].re * g[t].im + fin[t].im * g[t].re;
                }
            }
        }
    }
    for (int o = 0; o <= k; o++) negacyclic_inv_add(p, fout + (size_t)o * M, acc + (size_t)o * N);
}

/* luts: [T][(k+1)*N] GLWE accumulators (mask polys normally zero); lut_idx[e] in [0,T) or NULL */
/* bsk_ct_stride: 0 for a bootstrap (one key for the batch); for the blind rotation of the
 * wide-integer lookup every ciphertext brings its own list of n Fourier GGSWs (cplx elements
 * between the lists of consecutive ciphertexts). */
/* theta / nout: multi-output bootstrap (PBSmanyLUT, Chillotti-Ligier-Orfila-Tap 2021): the modulus switch
 * keeps theta zero low bits, the accumulator interleaves up to 2^theta functions (coefficient i belongs to
 * function i mod 2^theta) and coefficients 0..nout-1 are extracted; out is [B][nout][kN+1].  0 / 1 = plain. */
static inline int modswitch_coarse(u64 x, int log2N, int theta) {
    int lg = log2N + 1 - theta;
    u64 t = ((x >> (64 - lg - 1)) + 1) >> 1;
    return (int)((t & ((1ULL << lg) - 1)) << theta);
}

void ref_pbs_batch_ex(u64 *out, const u64 *in, long B, const double *bsk_f, const u64 *luts,
                      const int32_t *lut_idx, int n, int k, int N, int level, int base_log,
                      long bsk_ct_stride, int theta, int nout) {
    const fft_plan *p = get_plan(N);
    int log2N = 0; while ((1 << log2N) < N) log2N++;
    const int M = N / 2;
    const size_t ggsw_c = (size_t)level * (k + 1) * (k + 1) * M;
#pragma omp parallel
    {
        u64 *acc = (u64 *)malloc(sizeof(u64) * (size_t)(k + 1) * N);
        u64 *tmp = (u64 *)malloc(sizeof(u64) * (size_t)(k + 1) * N);
        double *dig = (double *)malloc(sizeof(double) * (size_t)level * N);
        cplx *fin = (cplx *)malloc(sizeof(cplx) * M);
        cplx *fout = (cplx *)malloc(sizeof(cplx) * (size_t)(k + 1) * M);
#pragma omp for schedule(dynamic, 1)
        for (long e = 0; e < B; e++) {
            const u64 *ct = in + (size_t)e * (n + 1);
            const u64 *lut = luts + (size_t)(lut_idx ? lut_idx[e] : 0) * (k + 1) * N;
            int bt = modswitch_coarse(ct[n], log2N, theta);
            for (int c = 0; c <= k; c++)
                monomial_mul(acc + (size_t)c * N, lut + (size_t)c * N, (2 * N - bt) % (2 * N), N);
            for (int i = 0; i < n; i++) {
                int at = modswitch_coarse(ct[i], log2N, theta);
                if (at == 0) continue;
                cmux_rotate(p, acc, at, (const cplx *)bsk_f + (size_t)e * bsk_ct_stride + (size_t)i * ggsw_c,
                            k, N, level, base_log, tmp, dig, fin, fout);
            }
            /* sample extract coefficients 0..nout-1: mask word m of polynomial c is A[u-m] (m <= u), -A[N+u-m] */
            for (int u = 0; u < nout; u++) {
                u64 *o = out + ((size_t)e * nout + u) * ((size_t)k * N + 1);
                for (int c = 0; c < k; c++) {
                    const u64 *A = acc + (size_t)c * N;
                    for (int m = 0; m <= u; m++) o[(size_t)c * N + m] = A[u - m];
                    for (int m = u + 1; m < N; m++) o[(size_t)c * N + m] = (u64)0 - A[N + u - m];
                }
                o[(size_t)k * N] = acc[(size_t)k * N + u];
            }
        }
        free(acc); free(tmp); free(dig); free(fin); free(fout);
    }
}

void ref_pbs_batch(u64 *out, const u64 *in, long B, const double *bsk_f, const u64 *luts,
                   const int32_t *lut_idx, int n, int k, int N, int level, int base_log) {
    ref_pbs_batch_ex(out, in, B, bsk_f, luts, lut_idx, n, k, N, level, base_log, 0, 0, 1);
}

/* Batched CMUX on GLWE ciphertexts (selection tree of the wide-integer lookup, "vertical packing",
 * Bergerat et al., "Parameter Optimization & Larger Precision for (T)FHE", 2022):
 *   out[b] = pool[i0[b]] + ggsw[gi[b]] (x) (pool[i1[b]] - pool[i0[b]])
 * pool: [P][(k+1)N] GLWE ciphertexts (plaintext polynomials are trivial GLWEs: zero mask);
 * ggsw_f: [G][level][k+1][k+1][M] as produced by ref_bsk_to_fourier. */
void ref_cmux_batch(u64 *out, const u64 *pool, const int32_t *i0, const int32_t *i1,
                    const double *ggsw_f, const int32_t *gi, long B, int k, int N, int level,
                    int base_log) {
    const fft_plan *p = get_plan(N);
    const int M = N / 2;
    const size_t glwe = (size_t)(k + 1) * N;
    const size_t ggsw_c = (size_t)level * (k + 1) * (k + 1) * M;
#pragma omp parallel
    {
        u64 *diff = (u64 *)malloc(sizeof(u64) * glwe);
        double *dig = (double *)malloc(sizeof(double) * (size_t)level * N);
        cplx *fin = (cplx *)malloc(sizeof(cplx) * M);
        cplx *fout = (cplx *)malloc(sizeof(cplx) * (size_t)(k + 1) * M);
        i64 digits[64];
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < B; b++) {
            const u64 *d0 = pool + (size_t)i0[b] * glwe, *d1 = pool + (size_t)i1[b] * glwe;
            const cplx *g = (const cplx *)ggsw_f + (size_t)gi[b] * ggsw_c;
            u64 *o = out + (size_t)b * glwe;
            for (size_t t = 0; t < glwe; t++) { diff[t] = d1[t] - d0[t]; o[t] = d0[t]; }
            memset(fout, 0, sizeof(cplx) * (size_t)(k + 1) * M);
            for (int c = 0; c <= k; c++) {
                for (int t = 0; t < N; t++) {
                    signed_decompose(diff[(size_t)c * N + t], base_log, level, digits);
                    for (int lv = 0; lv < level; lv++) dig[(size_t)lv * N + t] = (double)digits[lv];
                }
                for (int lv = 0; lv < level; lv++) {
                    negacyclic_fwd(p, dig + (size_t)lv * N, fin);
                    const cplx *row = g + ((size_t)lv * (k + 1) + c) * (k + 1) * M;
                    for (int oo = 0; oo <= k; oo++) {
                        const cplx *gg = row + (size_t)oo * M;
                        cplx *dst = fout + (size_t)oo * M;
                        for (int t = 0; t < M; t++) {
                            dst[t].re += fin[t].re * gg[t].re - fin[t].im * gg[t].im;
                            dst[t].im += fin[t].re * gg[t].im + fin[t].im * gg[t].re;
                        }
                    }
                }
            }
            for (int oo = 0; oo <= k; oo++) negacyclic_inv_add(p, fout + (size_t)oo * M, o + (size_t)oo * N);
        }
        free(diff); free(dig); free(fin); free(fout);
    }
}

/* LUT accumulator body polynomial: table of 2^p entries (values already
 * reduced mod 2^(p_out+1) by the caller or not -- they are wrapped by the shift). */
void ref_build_lut_poly(u64 *poly, const i64 *table, int p, int p_out, int N) {
    const int entries = 1 << p, box = N / entries, half = box / 2;
    for (int j = 0; j < N; j++) {
        int src = j + half;                 /* rotate left by half a box, negacyclic */
        int neg = 0;
        if (src >= N) { src -= N; neg = 1; }
        u64 v = ((u64)table[src / box]) << (64 - (p_out + 1));
        poly[j] = neg ? (u64)0 - v : v;
    }
}

/* ------------------------------------------------------- linear ops ------- */
/* all on LWE tensors stored [E][L] u64 (L = kN+1, body last) */

void ref_lin_add(u64 *out, const u64 *a, const u64 *b, long words) {
    for (long t = 0; t < words; t++) out[t] = a[t] + b[t];
}
void ref_lin_sub(u64 *out, const u64 *a, const u64 *b, long words) {
    for (long t = 0; t < words; t++) out[t] = a[t] - b[t];
}
void ref_lin_neg(u64 *out, const u64 *a, long words) {
    for (long t = 0; t < words; t++) out[t] = (u64)0 - a[t];
}
/* out[e] = a[e] with plaintext (already encoded) added to the body */
void ref_lin_add_plain(u64 *out, const u64 *a, const u64 *pt, long E, int L) {
    for (long e = 0; e < E; e++) {
        memcpy(out + (size_t)e * L, a + (size_t)e * L, sizeof(u64) * L);
        out[(size_t)e * L + L - 1] += pt[e];
    }
}
void ref_lin_mul_clear(u64 *out, const u64 *a, const i64 *c, long E, int L) {
    for (long e = 0; e < E; e++)
        for (int t = 0; t < L; t++) out[(size_t)e * L + t] = a[(size_t)e * L + t] * (u64)c[e];
}
/* general clear-weighted gather: out[e] = sum_j w[e][j] * x[idx[e][j]] + bias_pt[e] on the
 * body; idx < 0 means "no term".  Covers matmul / dot / conv2d / sum / broadcast. */
void ref_lin_gather_mac(u64 *out, const u64 *x, const int32_t *idx, const i64 *w, const u64 *bias,
                        long E, int K, int L) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < E; e++) {
        u64 *o = out + (size_t)e * L;
        memset(o, 0, sizeof(u64) * L);
        for (int j = 0; j < K; j++) {
            int32_t s = idx[(size_t)e * K + j];
            if (s < 0) continue;
            const u64 wj = (u64)w[(size_t)e * K + j];
            const u64 *src = x + (size_t)s * L;
            for (int t = 0; t < L; t++) o[t] += wj * src[t];
        }
        if (bias) o[L - 1] += bias[e];
    }
}
