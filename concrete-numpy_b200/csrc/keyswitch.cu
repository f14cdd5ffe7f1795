// Batched LWE keyswitch big -> small key, bit-exact on the torus.
//   out[b] = (0,...,0, body_in[b]) - sum_{i<kN} sum_{lv<level} digit_lv(a_in[b][i]) * KSK[i][lv][:]
// A digit-decomposition contraction over Z_2^64: D[B x kN*level] (small signed digits, built on
// the fly) times KSK[kN*level x (n+1)] (u64).  CUDA-core kernel: each CTA owns a tile of
// BT ciphertexts x 128 output columns; each thread keeps BT wrapping u64 accumulators in
// registers, KSK rows are read coalesced (128 consecutive columns) exactly once per CTA and
// the digits are broadcast from shared memory.  All arithmetic wraps mod 2^64, so the result
// is independent of summation order and matches the CPU oracle bit for bit.
//
// Replaces memref_batched_keyswitch_lwe_u64 of the absent concrete-compiler runtime, reached
// from concrete/numpy/compilation/server.py:286 for every table lookup
// (concrete/numpy/mlir/node_converter.py:1190-1206).
#include "common.cuh"

#define KS_COLS 128   // output columns per CTA (= threads)
#define KS_IC 32      // mask words decomposed per shared-memory chunk
#define KS_MAXL 16

// SPLITK: small batches (steps of sequential circuits) do not fill the GPU with (batch tile x column tile)
// CTAs, so blockIdx.z additionally splits the kN mask words; partial sums are combined with 64-bit atomic
// adds into a zero-initialised output -- wrapping sums are order independent, the result stays bit-exact.
template <int BT, bool SPLITK>
__global__ void __launch_bounds__(KS_COLS)
k_keyswitch(u64* __restrict__ out, const u64* __restrict__ in, long B, const u64* __restrict__ ksk,
            int kN, int n, int level, int base_log) {
    extern __shared__ int s_dig[];   // [KS_IC][level][BT]
    const int j = blockIdx.y * KS_COLS + threadIdx.x;       // output column
    const long b0 = (long)blockIdx.x * BT;
    const bool col_ok = j <= n;
    const int L_in = kN + 1, L_out = n + 1;
    int i_begin = 0, i_end = kN;
    if (SPLITK) {
        const int per = ((kN + gridDim.z - 1) / gridDim.z + KS_IC - 1) / KS_IC * KS_IC;
        i_begin = blockIdx.z * per;
        i_end = min(kN, i_begin + per);
    }

    u64 acc[BT];
#pragma unroll
    for (int b = 0; b < BT; b++) {
        long e = b0 + b;
        acc[b] = (j == n && e < B && (!SPLITK || blockIdx.z == 0)) ? in[(size_t)e * L_in + kN] : 0ULL;
    }

    for (int i0 = i_begin; i0 < i_end; i0 += KS_IC) {
        __syncthreads();
        // decompose BT x KS_IC mask words into shared memory
        for (int t = threadIdx.x; t < BT * KS_IC; t += KS_COLS) {
            int b = t / KS_IC, ii = t - b * KS_IC;
            long e = b0 + b;
            int i = i0 + ii;
            int dg[KS_MAXL];
#pragma unroll
            for (int q = 0; q < KS_MAXL; q++) dg[q] = 0;
            if (e < B && i < i_end) signed_decompose<KS_MAXL>(in[(size_t)e * L_in + i], base_log, level, dg);
#pragma unroll
            for (int q = 0; q < KS_MAXL; q++)
                if (q < level) s_dig[(ii * level + (level - 1 - q)) * BT + b] = dg[q];  // dg[q] <-> level (level-q)
        }
        __syncthreads();
        if (col_ok) {
            const int imax = min(KS_IC, i_end - i0);
            const u64* krow = ksk + (size_t)i0 * level * L_out + j;
            for (int r = 0; r < imax * level; r++) {
                const u64 kv = __ldg(krow + (size_t)r * L_out);
                const int4* dp = reinterpret_cast<const int4*>(s_dig + r * BT);
#pragma unroll
                for (int b4 = 0; b4 < BT / 4; b4++) {
                    int4 d = dp[b4];
                    acc[b4 * 4 + 0] -= (u64)(i64)d.x * kv;
                    acc[b4 * 4 + 1] -= (u64)(i64)d.y * kv;
                    acc[b4 * 4 + 2] -= (u64)(i64)d.z * kv;
                    acc[b4 * 4 + 3] -= (u64)(i64)d.w * kv;
                }
            }
        }
    }
    if (col_ok) {
#pragma unroll
        for (int b = 0; b < BT; b++) {
            long e = b0 + b;
            if (e < B) {
                if (SPLITK) atomicAdd(reinterpret_cast<unsigned long long*>(out + (size_t)e * L_out + j),
                                      (unsigned long long)acc[b]);
                else out[(size_t)e * L_out + j] = acc[b];
            }
        }
    }
}

extern "C" int fhe_keyswitch_batch(u64* out, const u64* in, long B, const u64* ksk, int kN, int n,
                                   int level, int base_log, void* stream) {
    if (B < 0 || kN <= 0 || n <= 0 || level <= 0 || level > KS_MAXL || base_log <= 0 ||
        level * base_log >= 64)
        FHE_FAIL(2, "keyswitch_batch: bad arguments (kN=%d n=%d level=%d base_log=%d)", kN, n, level, base_log);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int ycols = (n + 1 + KS_COLS - 1) / KS_COLS;
    if (B >= 512) {
        const int BT = 32;
        size_t smem = (size_t)KS_IC * level * BT * sizeof(int);
        k_keyswitch<BT, false><<<dim3((unsigned)((B + BT - 1) / BT), ycols), KS_COLS, smem, st>>>(
            out, in, B, ksk, kN, n, level, base_log);
    } else {
        // small batch: split the mask words too until a few CTAs per SM are in flight
        const int BT = 8;
        size_t smem = (size_t)KS_IC * level * BT * sizeof(int);
        const long tiles = ((B + BT - 1) / BT) * ycols;
        int dev = 0, sms = 0;
        FHE_CUDA(cudaGetDevice(&dev));
        FHE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        long ksplit = (4L * sms + tiles - 1) / tiles;
        const long max_split = (kN + KS_IC - 1) / KS_IC;
        if (ksplit > max_split) ksplit = max_split;
        if (ksplit <= 1) {
            k_keyswitch<BT, false><<<dim3((unsigned)((B + BT - 1) / BT), ycols), KS_COLS, smem, st>>>(
                out, in, B, ksk, kN, n, level, base_log);
        } else {
            FHE_CUDA(cudaMemsetAsync(out, 0, sizeof(u64) * (size_t)B * (n + 1), st));
            k_keyswitch<BT, true><<<dim3((unsigned)((B + BT - 1) / BT), ycols, (unsigned)ksplit), KS_COLS, smem, st>>>(
                out, in, B, ksk, kN, n, level, base_log);
        }
    }
    FHE_CHECK_LAUNCH();
    return 0;
}
