// Batched LWE keyswitch on the 5th-generation tensor cores (tcgen05, int8 limbs), bit-exact on the torus.
//
//   out[b][c] = [c == n] * body_in[b] - sum_{kk} digit(b, kk) * KSK[kk][c]        (mod 2^64)
//
// is an integer GEMM D[B x K] * KSK[K x (n+1)] with K = level * kN.  The signed gadget digits fit int8
// (base_log <= 7) and every u64 key word is split into 8 unsigned bytes, so the product becomes an
// int8 x uint8 -> int32 GEMM with N = 8 * (n+1) "limb columns"; the 8 int32 limb sums of one output word
// are recombined with shifts mod 2^64 in the epilogue, which is exact as long as no int32 sum overflows
// (K * 2^(base_log-1) * 255 < 2^31, checked by the host; otherwise the CUDA-core kernel is used).
//
// Data layouts are prepared so that every (tile, K-stage) operand block is CONTIGUOUS in global memory in
// exactly the shared-memory layout tcgen05.mma wants (K-major, no swizzle: 8-row x 16-byte core matrices),
// hence plain 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) feed the tensor cores, no tensor maps:
//   KSK_tc : [n-tile of 32 words][K/16][256 limb rows = (word c, limb l)][16 bytes of K]   (built once per key)
//   A_tc   : [m-tile of 128 cts][K/16][128 rows][16 bytes of K]                            (digits, per call)
// with the K index ordered level-major, kk = lv * kN + i, padded with zeros to a multiple of 128.
//
// Kernel (one 128 x 256 output tile per CTA, 192 threads): warp 0 = bulk-copy producer over a 4-stage
// ring, warp 1 = MMA issuer (one elected thread, accumulator 128 lanes x 256 columns of tensor memory),
// warps 2-5 = epilogue (tcgen05.ld, limb recombination, subtraction from the body).
//
// Replaces memref_batched_keyswitch_lwe_u64 of the absent concrete-compiler runtime (reached from
// concrete/numpy/compilation/server.py:286) for large batches; small batches keep csrc/keyswitch.cu.
#include "common.cuh"
#include "tmem.cuh"

#define KTC_BM 128          // ciphertexts per tile (TMEM lanes)
#define KTC_WORDS 32        // output words per tile
#define KTC_BN (8 * KTC_WORDS)   // limb columns per tile (TMEM columns)
#define KTC_BK 128          // K bytes per pipeline stage (4 MMAs of K = 32)
#define KTC_STAGES 4
#define KTC_A_STAGE (KTC_BM * KTC_BK)        // 16 KiB
#define KTC_B_STAGE (KTC_BN * KTC_BK)        // 32 KiB
#define KTC_THREADS 192

__device__ __forceinline__ unsigned ktc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ktc_mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ktc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ktc_mbar_expect_tx(u64* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ktc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ktc_mbar_wait(u64* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(ktc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ktc_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ktc_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(ktc_smem_u32(bar)) : "memory");
}
// tcgen05.commit: the mbarrier gets one arrival when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void ktc_commit(u64* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(ktc_smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (rows 16 B apart);
// SBO = byte distance between 8-row groups, LBO = byte distance between 16-byte K chunks
__device__ __forceinline__ u64 ktc_smem_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    u64 d = 0;
    d |= (u64)((smem_addr >> 4) & 0x3FFF);
    d |= (u64)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (u64)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (u64)1 << 46;           // descriptor version of sm_100
    return d;                     // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// instruction descriptor for kind::i8: D = s32, A = s8 (digits), B = u8 (key bytes), both K-major
__host__ __device__ constexpr unsigned ktc_instr_desc(int M, int N) {
    return (2u << 4)            // c_format = S32
         | (1u << 7)            // a_format = INT8
         | (0u << 10)           // b_format = UINT8
         | (0u << 15) | (0u << 16)   // a_major = b_major = K
         | ((unsigned)(N >> 3) << 17)
         | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void ktc_mma(unsigned tmem_c, u64 desc_a, u64 desc_b, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

// ---- key preparation: KSK [kN][level][n+1] u64 -> KSK_tc (see header) --------------------------------
__global__ void k_ksk_to_tc(unsigned char* __restrict__ dst, const u64* __restrict__ ksk, int kN, int n, int level,
                            long Kp, int ntiles) {
    // one thread per (ntile, kc, limb row): writes 16 bytes
    const long total = (long)ntiles * (Kp / 16) * KTC_BN;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int row = (int)(t % KTC_BN);
        const long kc = (t / KTC_BN) % (Kp / 16);
        const int nt = (int)(t / ((long)KTC_BN * (Kp / 16)));
        const int c = row >> 3, l = row & 7;
        const int col = nt * KTC_WORDS + c;
        unsigned char b[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const long kk = kc * 16 + e;                 // kk = lv * kN + i
            unsigned char v = 0;
            if (kk < (long)level * kN && col <= n) {
                const int lv = (int)(kk / kN), i = (int)(kk % kN);
                v = (unsigned char)(ksk[((size_t)i * level + lv) * (n + 1) + col] >> (8 * l));
            }
            b[e] = v;
        }
        *reinterpret_cast<uint4*>(dst + (size_t)t * 16) = *reinterpret_cast<const uint4*>(b);
    }
}

// ---- per call: signed digits of every mask word -> A_tc --------------------------------------------
// one thread per (row of the padded batch, group of 16 consecutive mask words): `level` 16-byte rows.
// Consecutive threads take consecutive rows, so the 16-byte stores of a warp are contiguous.
__global__ void k_ks_digits(unsigned char* __restrict__ dst, const u64* __restrict__ in, long B, int kN, int level,
                            int base_log, long Kp, int mtiles) {
    const long groups = kN / 16;
    const long kcs = Kp / 16;
    const long total = (long)mtiles * groups * KTC_BM;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int ml = (int)(t % KTC_BM);
        const long gi = (t / KTC_BM) % groups;
        const int mt = (int)(t / (KTC_BM * groups));
        const long m = (long)mt * KTC_BM + ml;
        unsigned char* base = dst + ((size_t)mt * kcs * KTC_BM + ml) * 16;       // + kc * 128 * 16
        int dig[16][8];      // [word][level], level <= 8 on this path
        if (m < B) {
            const u64* w = in + (size_t)m * (kN + 1) + gi * 16;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                int d[8];
                signed_decompose<8>(w[e], base_log, level, d);       // d[t] <-> level index level-1-t
#pragma unroll
                for (int t2 = 0; t2 < 8; t2++) dig[e][t2] = d[t2];
            }
        }
#pragma unroll
        for (int t2 = 0; t2 < 8; t2++) {
            if (t2 < level) {
                const int lv = level - 1 - t2;
                unsigned char b[16];
#pragma unroll
                for (int e = 0; e < 16; e++) b[e] = (m < B) ? (unsigned char)(signed char)dig[e][t2] : 0;
                const long kc = ((long)lv * kN) / 16 + gi;
                *reinterpret_cast<uint4*>(base + (size_t)kc * KTC_BM * 16) = *reinterpret_cast<const uint4*>(b);
            }
        }
        if (gi == 0)         // zero rows of the K padding
            for (long kc = ((long)level * kN) / 16; kc < kcs; kc++)
                *reinterpret_cast<uint4*>(base + (size_t)kc * KTC_BM * 16) = make_uint4(0, 0, 0, 0);
    }
}

// ---- the GEMM -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KTC_THREADS, 1)
k_ks_gemm(u64* __restrict__ out, const u64* __restrict__ in, long B, const unsigned char* __restrict__ A_tc,
          const unsigned char* __restrict__ ksk_tc, int kN, int n, long Kp, int mtiles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) u64 bar_full[KTC_STAGES], bar_empty[KTC_STAGES], bar_acc;
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.x % mtiles, nt = blockIdx.x / mtiles;       // CTAs sharing a key tile run together
    const long kcs = Kp / 16;
    const int nstages = (int)(Kp / KTC_BK);

    if (tid == 0) {
        for (int s = 0; s < KTC_STAGES; s++) {
            ktc_mbar_init(&bar_full[s], 1);
            ktc_mbar_init(&bar_empty[s], 1);
        }
        ktc_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&s_tmem, KTC_BN);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const unsigned tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* a_src = A_tc + (size_t)mt * kcs * KTC_BM * 16;
            const unsigned char* b_src = ksk_tc + (size_t)nt * kcs * KTC_BN * 16;
            for (int s = 0; s < nstages; s++) {
                const int slot = s % KTC_STAGES;
                if (s >= KTC_STAGES) ktc_mbar_wait(&bar_empty[slot], ((s / KTC_STAGES) - 1) & 1);
                unsigned char* a_dst = smem + (size_t)slot * (KTC_A_STAGE + KTC_B_STAGE);
                ktc_mbar_expect_tx(&bar_full[slot], KTC_A_STAGE + KTC_B_STAGE);
                ktc_bulk_g2s(a_dst, a_src + (size_t)s * KTC_A_STAGE, KTC_A_STAGE, &bar_full[slot]);
                ktc_bulk_g2s(a_dst + KTC_A_STAGE, b_src + (size_t)s * KTC_B_STAGE, KTC_B_STAGE, &bar_full[slot]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = ktc_instr_desc(KTC_BM, KTC_BN);
            for (int s = 0; s < nstages; s++) {
                const int slot = s % KTC_STAGES;
                ktc_mbar_wait(&bar_full[slot], (s / KTC_STAGES) & 1);
                tmem_fence_after_sync();
                const unsigned a_addr = ktc_smem_u32(smem + (size_t)slot * (KTC_A_STAGE + KTC_B_STAGE));
                const unsigned b_addr = a_addr + KTC_A_STAGE;
#pragma unroll
                for (int j = 0; j < KTC_BK / 32; j++) {
                    // one MMA consumes two 16-byte K chunks; chunk stride (LBO) = rows * 16 bytes
                    const u64 da = ktc_smem_desc(a_addr + j * 2 * (KTC_BM * 16), KTC_BM * 16, 128);
                    const u64 db = ktc_smem_desc(b_addr + j * 2 * (KTC_BN * 16), KTC_BN * 16, 128);
                    ktc_mma(tmem, da, db, idesc, (s | j) != 0 ? 1u : 0u);
                }
                ktc_commit(&bar_empty[slot]);          // slot reusable when these MMAs have read it
            }
            ktc_commit(&bar_acc);
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        ktc_mbar_wait(&bar_acc, 0);
        tmem_fence_after_sync();
        const int quarter = warp & 3;
        const int ml = quarter * 32 + lane;
        const long m = (long)mt * KTC_BM + ml;
        const unsigned taddr = tmem + ((unsigned)(quarter * 32) << 16);
#pragma unroll 1
        for (int chunk = 0; chunk < KTC_BN / 64; chunk++) {
            unsigned r[64];
            TmemIo<64>::ld(taddr + chunk * 64, r);
            tmem_wait_ld();
            if (m < B) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    u64 val = 0;
#pragma unroll
                    for (int l = 0; l < 8; l++) val += (u64)(i64)(int)r[c * 8 + l] << (8 * l);
                    const int col = nt * KTC_WORDS + chunk * 8 + c;
                    if (col <= n) {
                        const u64 body = (col == n) ? in[(size_t)m * (kN + 1) + kN] : 0ULL;
                        out[(size_t)m * (n + 1) + col] = body - val;
                    }
                }
            }
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, KTC_BN);
}

extern "C" {

static long ktc_Kp(int kN, int level) { return (((long)kN * level + KTC_BK - 1) / KTC_BK) * KTC_BK; }

// 1 when the tensor-core path is exact for these parameters (int8 digits, no int32 overflow)
int fhe_ks_tc_supported(int kN, int n, int level, int base_log) {
    if (kN < 16 || (kN % 16) != 0 || n < 1 || level < 1 || level > 8 || base_log < 1 || base_log > 7) return 0;
    const double worst = (double)kN * level * (double)(1 << (base_log - 1)) * 255.0;
    return worst < 2147483647.0 ? 1 : 0;
}
long fhe_ksk_tc_bytes(int kN, int n, int level) {
    const long ntiles = (n + 1 + KTC_WORDS - 1) / KTC_WORDS;
    return ntiles * ktc_Kp(kN, level) * KTC_BN;
}
long fhe_keyswitch_tc_scratch_bytes(long B, int kN, int level) {
    const long mtiles = (B + KTC_BM - 1) / KTC_BM;
    return mtiles * ktc_Kp(kN, level) * KTC_BM;
}
int fhe_ksk_to_tc(void* dst, const u64* ksk, int kN, int n, int level, void* stream) {
    const long Kp = ktc_Kp(kN, level);
    const int ntiles = (n + 1 + KTC_WORDS - 1) / KTC_WORDS;
    const long total = (long)ntiles * (Kp / 16) * KTC_BN;
    const int blocks = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    k_ksk_to_tc<<<blocks, 256, 0, (cudaStream_t)stream>>>((unsigned char*)dst, ksk, kN, n, level, Kp, ntiles);
    FHE_CHECK_LAUNCH();
    return 0;
}
int fhe_keyswitch_batch_tc(u64* out, const u64* in, long B, const void* ksk_tc, void* scratch, int kN, int n,
                           int level, int base_log, void* stream) {
    if (!fhe_ks_tc_supported(kN, n, level, base_log))
        FHE_FAIL(3, "keyswitch_tc: parameters outside the exact int8 range (kN=%d level=%d base_log=%d)", kN, level, base_log);
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long Kp = ktc_Kp(kN, level);
    const int mtiles = (int)((B + KTC_BM - 1) / KTC_BM);
    const int ntiles = (n + 1 + KTC_WORDS - 1) / KTC_WORDS;
    {
        const long total = (long)mtiles * (kN / 16) * KTC_BM;
        const int blocks = (int)((total + 127) / 128 < 1 << 20 ? (total + 127) / 128 : 1 << 20);
        k_ks_digits<<<blocks, 128, 0, st>>>((unsigned char*)scratch, in, B, kN, level, base_log, Kp, mtiles);
        FHE_CHECK_LAUNCH();
    }
    const size_t smem = (size_t)KTC_STAGES * (KTC_A_STAGE + KTC_B_STAGE) + 1024;
    FHE_CUDA(cudaFuncSetAttribute(k_ks_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ks_gemm<<<mtiles * ntiles, KTC_THREADS, smem, st>>>(out, in, B, (const unsigned char*)scratch,
                                                          (const unsigned char*)ksk_tc, kN, n, Kp, mtiles);
    FHE_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
