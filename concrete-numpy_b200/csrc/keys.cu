// Batched key generation, LWE encryption / decryption on device.
// Replaces ClientSupport.key_set / encrypt_arguments / decrypt_result of the absent
// concrete-compiler wheel (reference call sites: concrete/numpy/compilation/client.py:106,
// 178-182, 201).  Every random word comes from the counter-based ChaCha20 streams of
// common.cuh, so outputs are a pure function of (key material, object index); masks (public) and secrets / noise
// are streams of two different 256-bit keys (ChaKeys).
#include "common.cuh"

// ---------------------------------------------------------------- secret keys ---
__global__ void k_keygen_secret(u64* sk, int dim, ChaKey key, int domain) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    u64 w = rng_word(key, stream_id(domain, 0), (u64)(i >> 6));
    sk[i] = (w >> (i & 63)) & 1ULL;
}

// block-wide sum of one u64 per thread (wrapping)
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    u64 r = 0;
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; w++) r += sh[w];
    }
    return r;  // valid in thread 0
}

// ------------------------------------------------------------- LWE rows ---------
// One CTA per LWE row: mask words from (mask_dom, id), body = <a,s> + pt + e.
// Used for input encryption (pt = encoded message) and for KSK rows (pt = gadget term).
// mode 0: pt[row]                       (encrypt_lwe_batch)
// mode 1: sk_big[row/level] << (64-(row%level+1)*base_log)   (KSK)
__global__ void __launch_bounds__(128)
k_lwe_rows(u64* __restrict__ out, const u64* __restrict__ pt, const u64* __restrict__ sk, int dim,
           double std, ChaKeys keys, int mask_dom, int noise_dom, u64 first_id, int mode,
           const u64* __restrict__ sk_big, int level, int base_log) {
    __shared__ u64 sh[4];
    const long row = blockIdx.x;
    const u64 id = first_id + (u64)row;
    u64* dst = out + (size_t)row * (dim + 1);
    u64 partial = 0;
    const int nblk = (dim + 7) >> 3;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        u64 w[8];
        chacha20_block(keys.pub, stream_id(mask_dom, id), (u64)b, w);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            int j = b * 8 + t;
            if (j < dim) {
                dst[j] = w[t];
                partial += sk[j] ? w[t] : 0ULL;
            }
        }
    }
    u64 total = block_sum_u64(partial, sh);
    if (threadIdx.x == 0) {
        u64 m;
        if (mode == 0) m = pt[row];
        else {
            int lv = (int)(row % level);
            long i = row / level;
            m = sk_big[i] ? (1ULL << (64 - (lv + 1) * base_log)) : 0ULL;
        }
        dst[dim] = total + m + gaussian_torus(keys.sec, stream_id(noise_dom, id), 0, std);
    }
}

__global__ void k_trivial_lwe(u64* __restrict__ out, const u64* __restrict__ pt, long B, int dim) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = B * (long)(dim + 1);
    for (; t < total; t += (long)gridDim.x * blockDim.x) {
        long e = t / (dim + 1);
        int j = (int)(t - e * (dim + 1));
        out[t] = (j == dim) ? pt[e] : 0ULL;
    }
}

// phase[e] = body - <a, s>; one CTA per ciphertext.  If p >= 0 also decode the message.
__global__ void __launch_bounds__(128)
k_decrypt(u64* __restrict__ phase, u64* __restrict__ msg, const u64* __restrict__ ct,
          const u64* __restrict__ sk, int dim, int p) {
    __shared__ u64 sh[4];
    const long e = blockIdx.x;
    const u64* row = ct + (size_t)e * (dim + 1);
    u64 partial = 0;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) partial += sk[j] ? row[j] : 0ULL;
    u64 total = block_sum_u64(partial, sh);
    if (threadIdx.x == 0) {
        u64 ph = row[dim] - total;
        if (phase) phase[e] = ph;
        if (msg) {
            u64 t = ph >> (64 - p - 2);
            msg[e] = ((t >> 1) + (t & 1)) & ((1ULL << (p + 1)) - 1);
        }
    }
}

// --------------------------------------------------------------- BSK ------------
// step 1: masks (uniform) and body initialised with noise; grid = (rows, k+1)
__global__ void __launch_bounds__(256)
k_bsk_init(u64* __restrict__ bsk, int k, int N, double std, ChaKeys keys) {
    const long R = blockIdx.x;
    const int c = blockIdx.y;
    u64* poly = bsk + ((size_t)R * (k + 1) + c) * N;
    if (c < k) {
        const int nblk = N >> 3;
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
            u64 w[8];
            chacha20_block(keys.pub, stream_id(DOM_BSK_MASK, (u64)R), (u64)c * nblk + b, w);
#pragma unroll
            for (int t = 0; t < 8; t++) poly[b * 8 + t] = w[t];
        }
    } else {
        for (int t = threadIdx.x; t < N; t += blockDim.x)
            poly[t] = gaussian_torus(keys.sec, stream_id(DOM_BSK_NOISE, (u64)R), (u64)t, std);
    }
}

// step 2: body += sum_c A_c * S_c (negacyclic, S binary given as list of set positions)
// grid = (rows, N/256)
__global__ void __launch_bounds__(256)
k_bsk_body(u64* __restrict__ bsk, const int* __restrict__ ones, const int* __restrict__ ones_cnt,
           int k, int N) {
    const long R = blockIdx.x;
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    u64* row = bsk + (size_t)R * (k + 1) * N;
    u64 acc = 0;
    for (int c = 0; c < k; c++) {
        const u64* A = row + (size_t)c * N;
        const int* pos = ones + (size_t)c * N;
        const int cnt = ones_cnt[c];
        for (int u = 0; u < cnt; u++) {
            int idx = t - pos[u];
            u64 v = A[idx & (N - 1)];
            acc += (idx < 0) ? (0ULL - v) : v;
        }
    }
    row[(size_t)k * N + t] += acc;
}

// step 3: gadget terms  (after the body has consumed the pure masks)
__global__ void k_bsk_gadget(u64* __restrict__ bsk, const u64* __restrict__ sk_small, long rows,
                             int k, int N, int level, int base_log) {
    long R = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= rows) return;
    const int r = (int)(R % (k + 1));
    const int lv = (int)((R / (k + 1)) % level);
    const long i = R / ((long)(k + 1) * level);
    // key entries are 0/1 for an LWE key; the extended key of the packing keyswitch (wide-integer path) also
    // holds -1 for the body slot
    bsk[((size_t)R * (k + 1) + r) * N] += sk_small[i] * (1ULL << (64 - (lv + 1) * base_log));
}

// compact list of set positions of each GLWE key polynomial (single CTA, serial scan is fine)
__global__ void k_list_ones(const u64* __restrict__ sk_glwe, int k, int N, int* ones, int* ones_cnt) {
    int c = blockIdx.x;
    if (threadIdx.x != 0) return;
    int cnt = 0;
    for (int j = 0; j < N; j++)
        if (sk_glwe[(size_t)c * N + j]) ones[(size_t)c * N + cnt++] = j;
    ones_cnt[c] = cnt;
}

// ------------------------------------------------------------------ C ABI -------
extern "C" {

// test / reproducibility mode only: 64-bit seed -> the 16 words of key material (production: 64 bytes of OS entropy)
int fhe_seed_expand(u64 seed, u32* key16) {
    if (!key16) FHE_FAIL(2, "seed_expand: null output");
    chacha_key_from_seed(seed, key16);
    chacha_key_from_seed(seed ^ FHE_PUBLIC_SEED_TWEAK, key16 + 8);
    return 0;
}

int fhe_keygen_lwe_secret(u64* sk, int dim, const u32* key16, int domain, void* stream) {
    if (dim <= 0 || !key16 || (domain != DOM_SK_SMALL && domain != DOM_SK_BIG)) FHE_FAIL(2, "keygen_lwe_secret: bad arguments");
    k_keygen_secret<<<(dim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sk, dim, chacha_keys_load(key16).sec, domain);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_keygen_ksk(u64* ksk, const u64* sk_big, int kN, const u64* sk_small, int n, int level,
                   int base_log, double std, const u32* key16, void* stream) {
    if (kN <= 0 || n <= 0 || !key16 || level <= 0 || level * base_log >= 64) FHE_FAIL(2, "keygen_ksk: bad arguments");
    long rows = (long)kN * level;
    k_lwe_rows<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(
        ksk, nullptr, sk_small, n, std, chacha_keys_load(key16), DOM_KSK_MASK, DOM_KSK_NOISE, 0,
        1, sk_big, level, base_log);
    FHE_CHECK_LAUNCH();
    return 0;
}

// scratch: device buffer of at least (k*N + k) int32 words
int fhe_keygen_bsk(u64* bsk, const u64* sk_small, int n, const u64* sk_glwe, int k, int N,
                   int level, int base_log, double std, const u32* key16, int* scratch, void* stream) {
    if (n <= 0 || k <= 0 || !key16 || ilog2_exact(N) < 4 || N < 256 || level <= 0 || level * base_log >= 64)
        FHE_FAIL(2, "keygen_bsk: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)n * level * (k + 1);
    int* ones = scratch;
    int* ones_cnt = scratch + (size_t)k * N;
    k_list_ones<<<k, 32, 0, st>>>(sk_glwe, k, N, ones, ones_cnt);
    FHE_CHECK_LAUNCH();
    k_bsk_init<<<dim3((unsigned)rows, k + 1), 256, 0, st>>>(bsk, k, N, std, chacha_keys_load(key16));
    FHE_CHECK_LAUNCH();
    k_bsk_body<<<dim3((unsigned)rows, N / 256), 256, 0, st>>>(bsk, ones, ones_cnt, k, N);
    FHE_CHECK_LAUNCH();
    k_bsk_gadget<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(bsk, sk_small, rows, k, N, level, base_log);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_encrypt_lwe_batch(u64* ct, const u64* plaintexts, long B, const u64* sk, int dim, double std,
                          const u32* key16, u64 first_index, void* stream) {
    if (B < 0 || dim <= 0 || !key16) FHE_FAIL(2, "encrypt_lwe_batch: bad arguments");
    if (B == 0) return 0;
    k_lwe_rows<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(
        ct, plaintexts, sk, dim, std, chacha_keys_load(key16), DOM_ENC_MASK, DOM_ENC_NOISE,
        first_index, 0, nullptr, 1, 1);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_trivial_lwe_batch(u64* ct, const u64* plaintexts, long B, int dim, void* stream) {
    if (B < 0 || dim <= 0) FHE_FAIL(2, "trivial_lwe_batch: bad arguments");
    if (B == 0) return 0;
    long total = B * (long)(dim + 1);
    int grid = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
    k_trivial_lwe<<<grid, 256, 0, (cudaStream_t)stream>>>(ct, plaintexts, B, dim);
    FHE_CHECK_LAUNCH();
    return 0;
}

// phase and/or msg may be null; msg is decoded at precision p (one padding bit)
int fhe_decrypt_lwe_batch(u64* phase, u64* msg, const u64* ct, long B, const u64* sk, int dim, int p,
                          void* stream) {
    if (B < 0 || dim <= 0 || (msg && (p < 1 || p > 60))) FHE_FAIL(2, "decrypt_lwe_batch: bad arguments");
    if (B == 0) return 0;
    k_decrypt<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(phase, msg, ct, sk, dim, p);
    FHE_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
