// Torus linear algebra on LWE tensors stored [E][L] u64 (L = k*N+1, body last).
// HBM-bound u64 kernels: every global access is a warp-contiguous run of 8-byte words
// (16-byte vectors on the flat identity path).  Wrapping u64 arithmetic => bit-exact vs the
// CPU oracle regardless of summation order.
//
// Semantics follow the FHE / FHELinalg ops the reference emits
// (concrete/numpy/mlir/node_converter.py:216-242 add, :1062-1092 sub, :630-669 mul/neg,
//  :497-524 dot, :596-616 matmul, :436-483 conv2d, :1094-1127 sum, :244-264 broadcast).
#include "common.cuh"

// ---- flat elementwise (identity index maps), 16-byte vectors + scalar tail ----------
template <int OP>  // 0 add, 1 sub, 2 neg
__global__ void __launch_bounds__(256)
k_flat(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b, long words) {
    const long nvec = words >> 1;
    const long stride = (long)gridDim.x * blockDim.x;
    const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(a);
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(b);
    ulonglong2* o2 = reinterpret_cast<ulonglong2*>(out);
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += stride) {
        ulonglong2 x = a2[t], r;
        if (OP == 2) { r.x = 0ULL - x.x; r.y = 0ULL - x.y; }
        else {
            ulonglong2 y = b2[t];
            if (OP == 0) { r.x = x.x + y.x; r.y = x.y + y.y; }
            else { r.x = x.x - y.x; r.y = x.y - y.y; }
        }
        o2[t] = r;
    }
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        long t = words - 1;
        out[t] = (OP == 2) ? (0ULL - a[t]) : (OP == 0 ? a[t] + b[t] : a[t] - b[t]);
    }
}

// ---- row-mapped elementwise: out[e] = f(a[ia[e]], b[ib[e]]) with numpy broadcasting resolved
// on the host into int32 row maps (null = identity).  grid = (E, column chunks)
// OP: 0 add, 1 sub, 2 neg, 3 copy + plaintext on body, 4 multiply by clear int,
//     5 plaintext - ct (sub_int_eint): out = -a, body += pt
template <int OP>
__device__ __forceinline__ u64 row_op(u64 x, u64 y, u64 s, bool last) {
    if (OP == 0) return x + y;
    if (OP == 1) return x - y;
    if (OP == 2) return 0ULL - x;
    if (OP == 3) return x + (last ? s : 0ULL);
    if (OP == 4) return x * s;
    return (0ULL - x) + (last ? s : 0ULL);
}

// Rows are L = k*N+1 words long (odd), so consecutive rows alternate between 16-byte aligned and 8 bytes off.  When the
// rows an output row reads share its phase (always the case for identity maps), the row moves as 16-byte vectors with a
// scalar head / tail word; otherwise every thread keeps four independent 8-byte accesses in flight.
template <int OP>
__global__ void __launch_bounds__(256)
k_rows(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b,
       const int* __restrict__ ia, const int* __restrict__ ib, const u64* __restrict__ scal,
       int L) {
    const long e = blockIdx.x;
    const u64* ra = a + (size_t)(ia ? ia[e] : e) * L;
    const u64* rb = (OP <= 1) ? b + (size_t)(ib ? ib[e] : e) * L : ra;
    u64* ro = out + (size_t)e * L;
    const u64 s = (OP >= 3) ? scal[e] : 0ULL;
    const int nthr = gridDim.y * blockDim.x, tid = blockIdx.y * blockDim.x + threadIdx.x;
    const unsigned pa = (unsigned)((uintptr_t)ra & 15), pb = (unsigned)((uintptr_t)rb & 15), po = (unsigned)((uintptr_t)ro & 15);
    if (pa == po && pb == po) {
        const int head = po ? 1 : 0;                        // one scalar word brings the row to a 16-byte boundary
        if (tid == 0 && head) ro[0] = row_op<OP>(ra[0], rb[0], s, L == 1);
        const int nvec = (L - head) >> 1;
        const ulonglong2* va = reinterpret_cast<const ulonglong2*>(ra + head);
        const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(rb + head);
        ulonglong2* vo = reinterpret_cast<ulonglong2*>(ro + head);
        for (int v = tid; v < nvec; v += nthr) {
            const ulonglong2 x = va[v];
            const ulonglong2 y = (OP <= 1) ? vb[v] : x;
            const int t = head + 2 * v;
            ulonglong2 r;
            r.x = row_op<OP>(x.x, y.x, s, t == L - 1);
            r.y = row_op<OP>(x.y, y.y, s, t + 1 == L - 1);
            vo[v] = r;
        }
        const int done = head + 2 * nvec;
        if (tid == 0 && done < L) ro[done] = row_op<OP>(ra[done], rb[done], s, true);
        return;
    }
    for (int t0 = tid; t0 < L; t0 += 4 * nthr) {
        u64 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = t0 + u * nthr;
            x[u] = (t < L) ? ra[t] : 0ULL;
            y[u] = (OP <= 1 && t < L) ? rb[t] : 0ULL;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = t0 + u * nthr;
            if (t < L) ro[t] = row_op<OP>(x[u], y[u], s, t == L - 1);
        }
    }
}

// ---- general clear-weighted gather: out[e] = sum_j w[e][j] * x[idx[e][j]] (+ bias on body)
// idx < 0 or w == 0 => no term.  The taps of an output row are compacted into shared memory once per CTA by its first
// warp (ballot + prefix count), grouped as [+1 taps | -1 taps | general taps]: zero weights of ternary matrices and
// padding cost nothing, and +-1 weights (convolution stencils, sums, concatenations) are plain 64-bit adds instead of
// three 32-bit multiply-adds.  Covers dot / conv2d / sum and the matmuls the tiled kernels do not take;
// grid = (E, column chunks), GM_TPT columns per thread, coalesced 8-byte accesses (operand rows start on either
// 16-byte phase).  Operand rows shared by neighbouring outputs (convolutions) are re-read from L2, so with many taps
// per output this kernel is bound by L2 -> SM traffic and 64-bit integer adds, not by HBM: see DESIGN.md 4.3.
#define GM_TPT 8
#define GM_MAXTAPS 1024
__global__ void __launch_bounds__(256)
k_gather_mac(u64* __restrict__ out, const u64* __restrict__ x, const int* __restrict__ idx,
             const i64* __restrict__ w, const u64* __restrict__ bias, int K, int L) {
    __shared__ int s_src[GM_MAXTAPS];
    __shared__ i64 s_w[GM_MAXTAPS];
    __shared__ int s_cnt[3];
    const long e = blockIdx.x;
    const int* ie = idx + (size_t)e * K;
    const i64* we = w + (size_t)e * K;
    const u64 be = bias ? bias[e] : 0ULL;
    for (int k0 = 0; k0 < K; k0 += GM_MAXTAPS) {        // (K > GM_MAXTAPS: several rounds, accumulating into out)
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int k1 = (K - k0 < GM_MAXTAPS) ? K - k0 : GM_MAXTAPS;
            int c = 0;
            for (int cls = 0; cls < 3; cls++) {             // +1, -1, everything else
                const int c_start = c;
                for (int j0 = 0; j0 < k1; j0 += 32) {
                    const int j = j0 + lane;
                    const int sidx = (j < k1) ? ie[k0 + j] : -1;
                    const i64 wj = (j < k1) ? we[k0 + j] : 0;
                    const int kind = (wj == 1) ? 0 : (wj == -1) ? 1 : 2;
                    const bool keep = sidx >= 0 && wj != 0 && kind == cls;
                    const unsigned mask = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int pos = c + __popc(mask & ((1u << lane) - 1u));
                        s_src[pos] = sidx;
                        s_w[pos] = wj;
                    }
                    c += __popc(mask);
                }
                if (lane == 0) s_cnt[cls] = c - c_start;
            }
        }
        __syncthreads();
        const int cp = s_cnt[0], cn = s_cnt[1], cg = s_cnt[2];
        for (int t0 = blockIdx.y * blockDim.x * GM_TPT + threadIdx.x; t0 < L;
             t0 += gridDim.y * blockDim.x * GM_TPT) {
            u64 acc[GM_TPT];
#pragma unroll
            for (int u = 0; u < GM_TPT; u++) acc[u] = 0;
            int j = 0;
            for (; j < cp; j++) {
                const u64* src = x + (size_t)s_src[j] * L;
#pragma unroll
                for (int u = 0; u < GM_TPT; u++) {
                    const int t = t0 + u * blockDim.x;
                    if (t < L) acc[u] += src[t];
                }
            }
            for (; j < cp + cn; j++) {
                const u64* src = x + (size_t)s_src[j] * L;
#pragma unroll
                for (int u = 0; u < GM_TPT; u++) {
                    const int t = t0 + u * blockDim.x;
                    if (t < L) acc[u] -= src[t];
                }
            }
            for (; j < cp + cn + cg; j++) {
                const u64 wj = (u64)s_w[j];
                const u64* src = x + (size_t)s_src[j] * L;
#pragma unroll
                for (int u = 0; u < GM_TPT; u++) {
                    const int t = t0 + u * blockDim.x;
                    if (t < L) acc[u] += wj * src[t];
                }
            }
#pragma unroll
            for (int u = 0; u < GM_TPT; u++) {
                const int t = t0 + u * blockDim.x;
                if (t < L) {
                    u64* o = out + (size_t)e * L + t;
                    if (k0 == 0) *o = acc[u] + ((t == L - 1) ? be : 0ULL);
                    else *o += acc[u];
                }
            }
        }
    }
}

// ---- dense matmul  out[m][n] = sum_k W[k][n] * x[m][k]   (x: [Mr][K][L] ct, W: [K][Nc] clear)
// One CTA owns (row m, tile of TN output columns n, chunk of L): every x word is read once per
// n-tile and reused for TN outputs from registers, weights come from shared memory.
// Algorithmic bytes: (Mr*K + Mr*Nc)*L*8 when Nc <= TN.
#define MM_TN 16
__global__ void __launch_bounds__(256)
k_matmul(u64* __restrict__ out, const u64* __restrict__ x, const i64* __restrict__ W, int K, int Nc,
         int L, const int* __restrict__ nnz /* tap count of W (null: always run) */) {
    extern __shared__ i64 s_w[];   // [K][MM_TN]
    if (nnz && 2L * *nnz <= (long)K * Nc) return;     // sparse weights: the tap-list kernel took this matmul
    const long m = blockIdx.x;
    const int n0 = blockIdx.y * MM_TN;
    for (int t = threadIdx.x; t < K * MM_TN; t += blockDim.x) {
        int kk = t / MM_TN, nn = t - kk * MM_TN;
        s_w[t] = (n0 + nn < Nc) ? W[(size_t)kk * Nc + n0 + nn] : 0;
    }
    __syncthreads();
    const u64* xm = x + (size_t)m * K * L;
    for (int t = blockIdx.z * blockDim.x + threadIdx.x; t < L; t += gridDim.z * blockDim.x) {
        u64 acc[MM_TN];
#pragma unroll
        for (int u = 0; u < MM_TN; u++) acc[u] = 0;
        for (int kk = 0; kk < K; kk++) {
            const u64 v = xm[(size_t)kk * L + t];
            const i64* wr = s_w + kk * MM_TN;
#pragma unroll
            for (int u = 0; u < MM_TN; u++) acc[u] += (u64)wr[u] * v;
        }
#pragma unroll
        for (int u = 0; u < MM_TN; u++)
            if (n0 + u < Nc) out[((size_t)m * Nc + n0 + u) * L + t] = acc[u];
    }
}

// ---- matmul with the operand tile staged in shared memory: x is read from HBM exactly ONCE.
// The weight matrices of FHE circuits are small clear integers and mostly sparse (ternary layers, convolution
// stencils), so the columns of W are first compacted -- once per launch, by k_mm_taps -- into tap lists
//   [+1 taps | -1 taps | general taps]   (k indices, i64 weights for the general ones: any 64-bit weight is exact)
// and the main kernel does one add / subtract / multiply-add per NON-ZERO weight.
// CTA (m, chunk of MS_TC words): x[m][0..K)[chunk] -> shared memory with 8-byte cp.async (rows start on either 16-byte
// phase), then warp w owns output columns n = w, w + 8, ...: its lanes hold two word columns each (lane, lane + 32),
// tap lists are read warp-uniformly through the read-only cache, operand words come from shared memory.
// Algorithmic bytes = HBM traffic: (Mr*K + Mr*Nc)*L*8.  Dense matrices (more than half the weights non-zero) go to the
// register-blocked k_matmul above, which is bound by the 64-bit integer multiply-adds, not by HBM.
#define MS_TC 64
struct MmTaps {            // scratch layout, one launch
    int* cnt;              // [Nc][3]  number of +1, -1, general taps
    unsigned short* kk;    // [Nc][K]  their k, grouped in that order
    i64* ww;               // [Nc][K]  weights (aligned with kk; only the general ones are read)
    int* nnz;              // non-zero weights: at most half of K*Nc -> k_matmul_taps runs, else k_matmul
};
// one warp per column: three ballot-compaction sweeps over k (+1, -1, general); the total tap count goes to *t.nnz
// (zeroed by the host on the stream) and the consumers derive "sparse" from it
__global__ void __launch_bounds__(256)
k_mm_taps(const i64* __restrict__ W, int K, int Nc, MmTaps t) {
    const int lane = threadIdx.x & 31;
    const int n = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n >= Nc) return;
    unsigned short* kn = t.kk + (size_t)n * K;
    i64* wn = t.ww + (size_t)n * K;
    int c = 0;
    for (int cls = 0; cls < 3; cls++) {
        const int c_start = c;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const i64 w = (k < K) ? W[(size_t)k * Nc + n] : 0;
            const int kind = (w == 1) ? 0 : (w == -1) ? 1 : 2;
            const bool keep = w != 0 && kind == cls;
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = c + __popc(mask & ((1u << lane) - 1u));
                kn[pos] = (unsigned short)k;
                wn[pos] = w;
            }
            c += __popc(mask);
        }
        if (lane == 0) t.cnt[3 * n + cls] = c - c_start;
    }
    if (lane == 0) atomicAdd(t.nnz, c);
}
__device__ __forceinline__ bool mm_sparse(const MmTaps& t, int K, int Nc) { return 2L * *t.nnz <= (long)K * Nc; }

// persistent CTAs, `stages` operand tiles in flight: while the warps reduce tile i, the copy engine brings tiles
// i+1 .. i+stages-1.  Each of the K rows of a tile is ONE bulk asynchronous copy (cp.async.bulk global -> shared, 16-byte
// aligned middle of the row) counted on the stage's mbarrier, plus at most a head and a tail word moved by the issuing
// thread: rows are L = k*N+1 words apart, so every other row starts 8 bytes off a 16-byte boundary.  A row keeps its
// global 16-byte phase in shared memory (row stride MS_RS = MS_TC + 2 words, data at offset `phase`).
#define MS_RS (MS_TC + 2)
#define MM_NCW 8           // output columns per warp held in registers (Nc <= MM_NCW * 16 warps)
__device__ __forceinline__ unsigned mm_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mm_stage_tile(u64* xt, u64* bar, const u64* __restrict__ x, long tile, int nchunk, int K, int L) {
    const long m = tile / nchunk;
    const int t0 = (int)(tile - m * nchunk) * MS_TC;
    const int ncol = (L - t0 < MS_TC) ? L - t0 : MS_TC;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const u64* src = x + ((size_t)m * K + k) * L + t0;
        const int ph = (int)(((uintptr_t)src >> 3) & 1);
        u64* dst = xt + (size_t)k * MS_RS + ph;
        const int head = ph ? 1 : 0;                                   // words before the first 16-byte boundary
        const int mid = (ncol - head) & ~1;                            // words moved by the bulk copy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mm_smem(bar)), "r"((unsigned)(mid * 8)) : "memory");
        if (mid > 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(mm_smem(dst + head)), "l"(src + head), "r"((unsigned)(mid * 8)), "r"(mm_smem(bar)) : "memory");
        // head / tail words: 8-byte asynchronous copies, their completion is this thread's second arrival on the barrier
        // (nobody blocks on a global load while staging)
        if (head && ncol > 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mm_smem(dst)), "l"(src) : "memory");
        if (head + mid < ncol)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mm_smem(dst + head + mid)), "l"(src + head + mid) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mm_smem(bar)) : "memory");
        for (int c = ncol; c < MS_TC; c++) dst[c] = 0ULL;             // ragged last chunk of a row
    }
}

__global__ void __launch_bounds__(512)
k_matmul_taps(u64* __restrict__ out, const u64* __restrict__ x, int K, int Nc, int L, long Mr, int stages, MmTaps t) {
    extern __shared__ __align__(16) unsigned char ms_raw[];
    __shared__ __align__(8) u64 s_bar[4];
    if (!mm_sparse(t, K, Nc)) return;                     // dense weights: k_matmul takes this product
    u64* xs = reinterpret_cast<u64*>(ms_raw);             // [stages][K][MS_RS]
    const size_t stage_words = (size_t)K * MS_RS;
    const int nchunk = (L + MS_TC - 1) / MS_TC;
    const long tiles = Mr * nchunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mm_smem(&s_bar[s])), "r"(2 * K) : "memory");   // two arrivals per row
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // The CTA is persistent and a warp owns the same output columns (warp, warp + nwarps, ...) in every tile: the first
    // 32 taps of each of its columns stay in registers for the whole kernel (lane l holds tap l)
    int cnts[MM_NCW], kregs[MM_NCW];
    i64 wregs[MM_NCW];
#pragma unroll
    for (int ci = 0; ci < MM_NCW; ci++) {
        const int n = warp + ci * nwarps;
        cnts[ci] = (n < Nc && lane < 3) ? __ldg(t.cnt + 3 * n + lane) : 0;
        kregs[ci] = (n < Nc) ? (int)__ldg(t.kk + (size_t)n * K + (lane < K ? lane : 0)) : 0;
        wregs[ci] = (n < Nc) ? __ldg(t.ww + (size_t)n * K + (lane < K ? lane : 0)) : 0;
    }
    // prologue: the first stages-1 tiles of this CTA
    for (int s = 0; s < stages - 1; s++) {
        const long tile = blockIdx.x + (long)s * gridDim.x;
        if (tile < tiles) mm_stage_tile(xs + s * stage_words, &s_bar[s], x, tile, nchunk, K, L);
    }
    long it = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        {
            const long ahead = tile + (long)(stages - 1) * gridDim.x;
            const int sa = (int)((it + stages - 1) % stages);
            if (ahead < tiles || stages == 1) mm_stage_tile(xs + sa * stage_words, &s_bar[sa], x, stages == 1 ? tile : ahead, nchunk, K, L);
        }
        const int sc = (int)(it % stages);
        {
            const unsigned parity = (unsigned)((it / stages) & 1);
            asm volatile(
                "{\n.reg .pred p;\nMMW_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra MMD_%=;\nbra MMW_%=;\nMMD_%=:\n}\n" ::"r"(mm_smem(&s_bar[sc])), "r"(parity) : "memory");
        }
        __syncthreads();                                   // head / tail words were written by ordinary stores
        const u64* xt = xs + sc * stage_words;
        const long m = tile / nchunk;
        const int t0 = (int)(tile - m * nchunk) * MS_TC;
        const int ncol = (L - t0 < MS_TC) ? L - t0 : MS_TC;
        // 16-byte phase of row k of this tile: rows are L words apart
        const size_t row0 = (size_t)m * K * L + t0;
        const int ph0 = (int)((((uintptr_t)x >> 3) + row0) & 1), phstep = L & 1;
        // Tap lists live in global memory (L2): lane l of the warp holds tap l of the column (one coalesced load per 32
        // taps, requested one column ahead) and the loop broadcasts k / w with shuffles -- no dependent global load
        // sits between two shared-memory reads.
        // byte offset of row k inside the stage (row stride + its 16-byte phase), computed once per lane-held tap
#define MM_OFF(kk_) (((kk_) * MS_RS + ((ph0 + (kk_) * phstep) & 1)) << 3)
#define MM_ROW(off_) reinterpret_cast<const u64*>(xl + (off_))
        const unsigned char* xl = reinterpret_cast<const unsigned char*>(xt + lane);
#pragma unroll
        for (int ci = 0; ci < MM_NCW; ci++) {
            const int n = warp + ci * nwarps;
            if (n >= Nc) break;
            const int cp = __shfl_sync(0xffffffffu, cnts[ci], 0), cn = __shfl_sync(0xffffffffu, cnts[ci], 1),
                      cg = __shfl_sync(0xffffffffu, cnts[ci], 2);
            int oreg = MM_OFF(kregs[ci]);
            i64 wreg = wregs[ci];
            const unsigned short* kn = t.kk + (size_t)n * K;
            const i64* wn = t.ww + (size_t)n * K;
            const int total = cp + cn + cg;
            u64 a0 = 0, a1 = 0, b0 = 0, b1 = 0;            // two word columns, two independent chains each
            for (int base = 0; base < total; base += 32) {
                if (base) {                                 // columns with more than 32 taps: further rounds from L2
                    const int kk2 = (base + lane < total) ? (int)__ldg(kn + base + lane) : 0;
                    oreg = MM_OFF(kk2);
                    wreg = (base + lane < total) ? __ldg(wn + base + lane) : 0;
                }
                const int lim = (total - base < 32) ? total - base : 32;
                int jj = 0;
                const int e_pos = (cp - base < 0) ? 0 : (cp - base < lim ? cp - base : lim);            // +1 taps of this round
                const int e_neg = (cp + cn - base < 0) ? 0 : (cp + cn - base < lim ? cp + cn - base : lim);
                for (; jj + 1 < e_pos; jj += 2) {
                    const u64* r0 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj));
                    const u64* r1 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj + 1));
                    a0 += r0[0]; a1 += r0[32]; b0 += r1[0]; b1 += r1[32];
                }
                if (jj < e_pos) { const u64* r0 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj)); a0 += r0[0]; a1 += r0[32]; jj++; }
                for (; jj + 1 < e_neg; jj += 2) {
                    const u64* r0 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj));
                    const u64* r1 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj + 1));
                    a0 -= r0[0]; a1 -= r0[32]; b0 -= r1[0]; b1 -= r1[32];
                }
                if (jj < e_neg) { const u64* r0 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj)); a0 -= r0[0]; a1 -= r0[32]; jj++; }
                for (; jj < lim; jj++) {
                    const u64 w = (u64)__shfl_sync(0xffffffffu, wreg, jj);
                    const u64* r0 = MM_ROW(__shfl_sync(0xffffffffu, oreg, jj));
                    a0 += w * r0[0]; a1 += w * r0[32];
                }
            }
            u64* o = out + ((size_t)m * Nc + n) * L + t0;
            if (lane < ncol) o[lane] = a0 + b0;
            if (lane + 32 < ncol) o[lane + 32] = a1 + b1;
        }
#undef MM_ROW
#undef MM_OFF
        __syncthreads();                                   // the stage is overwritten by the copies of a later tile
    }
}

static int g_sm_count = 0;
static inline int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            g_sm_count = sms;
        else
            g_sm_count = 148;
    }
    return g_sm_count;
}

// column chunks of a row of L words: a remainder under a quarter of a block (L = k*N + 1: the body word) does not get a
// block of its own, the kernels' grid-stride loops give it to the first block as a second trip
static inline int col_chunks(int L, int per_block) {
    int c = L / per_block;
    if (L - c * per_block >= per_block / 4) c++;
    return c < 1 ? 1 : (c > 64 ? 64 : c);
}

extern "C" {

#define LIN_ARGS_OK(E, L) ((E) >= 0 && (L) > 0)

// the flat path moves 16-byte vectors: row slices of tensors with an odd row length may start 8 bytes off
#define LIN_ALIGNED16(o, x, y) ((((uintptr_t)(o) | (uintptr_t)(x) | (uintptr_t)(y)) & 15) == 0)

int fhe_lin_add(u64* out, const u64* a, const u64* b, const int* ia, const int* ib, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_add: bad arguments");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!ia && !ib && LIN_ALIGNED16(out, a, b)) {
        long words = E * (long)L;
        const long cap = (long)sm_count() * 8;
        int grid = (int)((words / 2 + 255) / 256 < cap ? (words / 2 + 255) / 256 : cap);
        k_flat<0><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, b, words);
    } else
        k_rows<0><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, st>>>(out, a, b, ia, ib, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_lin_sub(u64* out, const u64* a, const u64* b, const int* ia, const int* ib, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_sub: bad arguments");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!ia && !ib && LIN_ALIGNED16(out, a, b)) {
        long words = E * (long)L;
        const long cap = (long)sm_count() * 8;
        int grid = (int)((words / 2 + 255) / 256 < cap ? (words / 2 + 255) / 256 : cap);
        k_flat<1><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, b, words);
    } else
        k_rows<1><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, st>>>(out, a, b, ia, ib, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_lin_neg(u64* out, const u64* a, const int* ia, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_neg: bad arguments");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!ia && LIN_ALIGNED16(out, a, a)) {
        long words = E * (long)L;
        const long cap = (long)sm_count() * 8;
        int grid = (int)((words / 2 + 255) / 256 < cap ? (words / 2 + 255) / 256 : cap);
        k_flat<2><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, nullptr, words);
    } else
        k_rows<2><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, st>>>(out, a, nullptr, ia, nullptr, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = a[ia[e]] with encoded plaintext pt[e] added to the body (add_eint_int / sub_eint_int)
int fhe_lin_add_plain(u64* out, const u64* a, const int* ia, const u64* pt, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_add_plain: bad arguments");
    if (E == 0) return 0;
    k_rows<3><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, (cudaStream_t)stream>>>(out, a, nullptr, ia, nullptr, pt, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = pt[e] - a[ia[e]]   (sub_int_eint)
int fhe_lin_plain_sub(u64* out, const u64* a, const int* ia, const u64* pt, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_plain_sub: bad arguments");
    if (E == 0) return 0;
    k_rows<5><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, (cudaStream_t)stream>>>(out, a, nullptr, ia, nullptr, pt, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = a[ia[e]] * c[e]   (mul_eint_int; c signed, wrapping)
int fhe_lin_mul_clear(u64* out, const u64* a, const int* ia, const i64* c, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_mul_clear: bad arguments");
    if (E == 0) return 0;
    k_rows<4><<<dim3((unsigned)E, col_chunks(L, 2048)), 256, 0, (cudaStream_t)stream>>>(
        out, a, nullptr, ia, nullptr, reinterpret_cast<const u64*>(c), L);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_lin_gather_mac(u64* out, const u64* x, const int* idx, const i64* w, const u64* bias, long E,
                       int K, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L) || K <= 0) FHE_FAIL(2, "lin_gather_mac: bad arguments");
    if (E == 0) return 0;
    k_gather_mac<<<dim3((unsigned)E, col_chunks(L, 256 * GM_TPT)), 256, 0, (cudaStream_t)stream>>>(
        out, x, idx, w, bias, K, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// x: [Mr][K][L], W: [K][Nc] (device, i64), out: [Mr][Nc][L]
int fhe_matmul_ct_clear(u64* out, const u64* x, const i64* W, long Mr, int K, int Nc, int L, void* stream) {
    if (Mr < 0 || K <= 0 || Nc <= 0 || L <= 0) FHE_FAIL(2, "matmul_ct_clear: bad arguments");
    if (Mr == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_dense = (size_t)K * MM_TN * sizeof(i64);
    const size_t smem_taps = (size_t)K * MS_RS * 8;
    if (smem_dense > 48 * 1024 || smem_taps > 200 * 1024)
        FHE_FAIL(3, "matmul_ct_clear: K=%d too large for the tiled kernels (use gather_mac)", K);
    unsigned char* scratch = nullptr;
    if (Nc > MM_NCW * 16) {
        // more output columns than the tap-list kernel keeps in registers: the register-blocked kernel takes every W
        k_matmul<<<dim3((unsigned)Mr, (Nc + MM_TN - 1) / MM_TN, col_chunks(L, 256)), 256, smem_dense, st>>>(out, x, W, K, Nc, L, nullptr);
        FHE_CHECK_LAUNCH();
        return 0;
    }
    // which kernel runs is decided ON DEVICE by the density of W (no host synchronisation inside Server.run): both are
    // enqueued and the one the tap count rules out exits at once
    MmTaps t = {};
    const size_t b_cnt = ((size_t)Nc * 3 * sizeof(int) + 15) & ~(size_t)15;
    const size_t b_kk = ((size_t)Nc * K * sizeof(unsigned short) + 15) & ~(size_t)15;
    const size_t b_ww = (size_t)Nc * K * sizeof(i64);
    FHE_CUDA(fhe_malloc_async(reinterpret_cast<void**>(&scratch), b_cnt + b_kk + b_ww + 16, st));
    t.cnt = reinterpret_cast<int*>(scratch);
    t.kk = reinterpret_cast<unsigned short*>(scratch + b_cnt);
    t.ww = reinterpret_cast<i64*>(scratch + b_cnt + b_kk);
    t.nnz = reinterpret_cast<int*>(scratch + b_cnt + b_kk + b_ww);
    FHE_CUDA(cudaMemsetAsync(t.nnz, 0, sizeof(int), st));
    k_mm_taps<<<(unsigned)((Nc * 32 + 255) / 256), 256, 0, st>>>(W, K, Nc, t);
    FHE_CHECK_LAUNCH();
    {
        // as many operand tiles in flight per CTA as shared memory holds (up to 4), persistent CTAs of 16 warps
        const size_t stage = (size_t)K * MS_RS * 8;
        int stages = (int)((200 * 1024) / stage);
        stages = stages > 4 ? 4 : (stages < 1 ? 1 : stages);
        const long tiles = Mr * (long)((L + MS_TC - 1) / MS_TC);
        const int per_sm = (stage * stages <= 100 * 1024) ? 2 : 1;
        long grid = (long)sm_count() * per_sm;
        if (grid > tiles) grid = tiles;
        static size_t smem_allowed[64] = {};              // per device: raise the opt-in limit only when it grows
        int dev = 0;
        FHE_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || smem_allowed[dev] < stage * stages) {
            FHE_CUDA(cudaFuncSetAttribute(k_matmul_taps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(stage * stages)));
            if (dev >= 0 && dev < 64) smem_allowed[dev] = stage * stages;
        }
        k_matmul_taps<<<(unsigned)grid, 512, stage * stages, st>>>(out, x, K, Nc, L, Mr, stages, t);
        FHE_CHECK_LAUNCH();
    }
    k_matmul<<<dim3((unsigned)Mr, (Nc + MM_TN - 1) / MM_TN, col_chunks(L, 256)), 256, smem_dense, st>>>(out, x, W, K, Nc, L, t.nnz);
    FHE_CHECK_LAUNCH();
    if (scratch) FHE_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

}  // extern "C"
