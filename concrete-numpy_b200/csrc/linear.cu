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
         int L, const int* __restrict__ skip_if_set) {
    extern __shared__ i64 s_w[];   // [K][MM_TN]
    if (skip_if_set && *skip_if_set) return;          // the staged kernel took this matmul
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
    int* sparse;           // 1: k_matmul_taps runs, 0: k_matmul runs
};
__global__ void k_mm_taps(const i64* __restrict__ W, int K, int Nc, MmTaps t) {
    __shared__ int s_nnz;
    if (threadIdx.x == 0) s_nnz = 0;
    __syncthreads();
    int mine = 0;
    for (int n = threadIdx.x; n < Nc; n += blockDim.x) {
        unsigned short* kn = t.kk + (size_t)n * K;
        i64* wn = t.ww + (size_t)n * K;
        int c = 0, c0, c1;
        for (int k = 0; k < K; k++) if (W[(size_t)k * Nc + n] == 1) kn[c++] = (unsigned short)k;
        c0 = c;
        for (int k = 0; k < K; k++) if (W[(size_t)k * Nc + n] == -1) kn[c++] = (unsigned short)k;
        c1 = c;
        for (int k = 0; k < K; k++) {
            const i64 w = W[(size_t)k * Nc + n];
            if (w != 0 && w != 1 && w != -1) { kn[c] = (unsigned short)k; wn[c] = w; c++; }
        }
        t.cnt[3 * n] = c0; t.cnt[3 * n + 1] = c1 - c0; t.cnt[3 * n + 2] = c - c1;
        mine += c;
    }
    atomicAdd(&s_nnz, mine);
    __syncthreads();
    if (threadIdx.x == 0) *t.sparse = (2L * s_nnz <= (long)K * Nc) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_matmul_taps(u64* __restrict__ out, const u64* __restrict__ x, int K, int Nc, int L, MmTaps t) {
    extern __shared__ __align__(16) unsigned char ms_raw[];
    if (!*t.sparse) return;                               // dense weights: k_matmul takes this product
    u64* xt = reinterpret_cast<u64*>(ms_raw);             // [K][MS_TC]
    const long m = blockIdx.x;
    const int t0 = blockIdx.y * MS_TC;
    const u64* xm = x + (size_t)m * K * L + t0;
    const int ncol = (L - t0 < MS_TC) ? L - t0 : MS_TC;
    // stage the tile: consecutive threads copy consecutive words of a row (8-byte asynchronous copies, no registers)
    for (int i = threadIdx.x; i < K * MS_TC; i += blockDim.x) {
        const int k = i / MS_TC, c = i - k * MS_TC;
        if (c < ncol) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(xt + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(xm + (size_t)k * L + c) : "memory");
        } else
            xt[i] = 0ULL;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int n = warp; n < Nc; n += nwarps) {
        const int cp = __ldg(t.cnt + 3 * n), cn = __ldg(t.cnt + 3 * n + 1), cg = __ldg(t.cnt + 3 * n + 2);
        const unsigned short* kn = t.kk + (size_t)n * K;
        const i64* wn = t.ww + (size_t)n * K;
        u64 a0 = 0, a1 = 0, b0 = 0, b1 = 0;                // two word columns, two independent chains each
        int j = 0;
        for (; j + 1 < cp; j += 2) {
            const u64* r0 = xt + (int)__ldg(kn + j) * MS_TC + lane;
            const u64* r1 = xt + (int)__ldg(kn + j + 1) * MS_TC + lane;
            a0 += r0[0]; a1 += r0[32]; b0 += r1[0]; b1 += r1[32];
        }
        if (j < cp) { const u64* r0 = xt + (int)__ldg(kn + j) * MS_TC + lane; a0 += r0[0]; a1 += r0[32]; j++; }
        for (; j + 1 < cp + cn; j += 2) {
            const u64* r0 = xt + (int)__ldg(kn + j) * MS_TC + lane;
            const u64* r1 = xt + (int)__ldg(kn + j + 1) * MS_TC + lane;
            a0 -= r0[0]; a1 -= r0[32]; b0 -= r1[0]; b1 -= r1[32];
        }
        if (j < cp + cn) { const u64* r0 = xt + (int)__ldg(kn + j) * MS_TC + lane; a0 -= r0[0]; a1 -= r0[32]; j++; }
        for (; j < cp + cn + cg; j++) {
            const u64 w = (u64)__ldg(wn + j);
            const u64* r0 = xt + (int)__ldg(kn + j) * MS_TC + lane;
            a0 += w * r0[0]; a1 += w * r0[32];
        }
        u64* o = out + ((size_t)m * Nc + n) * L + t0;
        if (lane < ncol) o[lane] = a0 + b0;
        if (lane + 32 < ncol) o[lane + 32] = a1 + b1;
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
    const size_t smem_taps = (size_t)K * MS_TC * 8;
    if (smem_dense > 48 * 1024 || smem_taps > 200 * 1024)
        FHE_FAIL(3, "matmul_ct_clear: K=%d too large for the tiled kernels (use gather_mac)", K);
    // which kernel runs is decided ON DEVICE by the density of W (no host synchronisation inside Server.run): both are
    // enqueued and the one the flag rules out exits at once
    unsigned char* scratch = nullptr;
    MmTaps t = {};
    const size_t b_cnt = ((size_t)Nc * 3 * sizeof(int) + 15) & ~(size_t)15;
    const size_t b_kk = ((size_t)Nc * K * sizeof(unsigned short) + 15) & ~(size_t)15;
    const size_t b_ww = (size_t)Nc * K * sizeof(i64);
    FHE_CUDA(cudaMallocAsync(&scratch, b_cnt + b_kk + b_ww + 16, st));
    t.cnt = reinterpret_cast<int*>(scratch);
    t.kk = reinterpret_cast<unsigned short*>(scratch + b_cnt);
    t.ww = reinterpret_cast<i64*>(scratch + b_cnt + b_kk);
    t.sparse = reinterpret_cast<int*>(scratch + b_cnt + b_kk + b_ww);
    k_mm_taps<<<1, 256, 0, st>>>(W, K, Nc, t);
    FHE_CHECK_LAUNCH();
    FHE_CUDA(cudaFuncSetAttribute(k_matmul_taps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_taps));
    k_matmul_taps<<<dim3((unsigned)Mr, (unsigned)((L + MS_TC - 1) / MS_TC)), 256, smem_taps, st>>>(out, x, K, Nc, L, t);
    FHE_CHECK_LAUNCH();
    k_matmul<<<dim3((unsigned)Mr, (Nc + MM_TN - 1) / MM_TN, col_chunks(L, 256)), 256, smem_dense, st>>>(out, x, W, K, Nc, L, t.sparse);
    FHE_CHECK_LAUNCH();
    if (scratch) FHE_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

}  // extern "C"
