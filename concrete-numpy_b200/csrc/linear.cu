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
__global__ void __launch_bounds__(256)
k_rows(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b,
       const int* __restrict__ ia, const int* __restrict__ ib, const u64* __restrict__ scal,
       int L) {
    const long e = blockIdx.x;
    const u64* ra = a + (size_t)(ia ? ia[e] : e) * L;
    const u64* rb = (OP <= 1) ? b + (size_t)(ib ? ib[e] : e) * L : nullptr;
    u64* ro = out + (size_t)e * L;
    const u64 s = (OP >= 3) ? scal[e] : 0ULL;
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < L; t += gridDim.y * blockDim.x) {
        u64 x = ra[t], r;
        if (OP == 0) r = x + rb[t];
        else if (OP == 1) r = x - rb[t];
        else if (OP == 2) r = 0ULL - x;
        else if (OP == 3) r = x + ((t == L - 1) ? s : 0ULL);
        else if (OP == 4) r = x * s;
        else r = (0ULL - x) + ((t == L - 1) ? s : 0ULL);
        ro[t] = r;
    }
}

// ---- general clear-weighted gather: out[e] = sum_j w[e][j] * x[idx[e][j]] (+ bias on body)
// idx < 0 or w == 0 => no term (zero weights of ternary matrices are skipped warp-uniformly).
// Covers dot / matmul / conv2d / sum in one kernel; grid = (E, column chunks), 4 columns/thread.
#define GM_TPT 4
__global__ void __launch_bounds__(256)
k_gather_mac(u64* __restrict__ out, const u64* __restrict__ x, const int* __restrict__ idx,
             const i64* __restrict__ w, const u64* __restrict__ bias, int K, int L) {
    const long e = blockIdx.x;
    const int* ie = idx + (size_t)e * K;
    const i64* we = w + (size_t)e * K;
    for (int t0 = blockIdx.y * blockDim.x * GM_TPT + threadIdx.x; t0 < L;
         t0 += gridDim.y * blockDim.x * GM_TPT) {
        u64 acc[GM_TPT];
#pragma unroll
        for (int u = 0; u < GM_TPT; u++) acc[u] = 0;
        for (int j = 0; j < K; j++) {
            const int s = __ldg(ie + j);
            const u64 wj = (u64)__ldg(we + j);
            if (s < 0 || wj == 0) continue;
            const u64* src = x + (size_t)s * L;
#pragma unroll
            for (int u = 0; u < GM_TPT; u++) {
                int t = t0 + u * blockDim.x;
                if (t < L) acc[u] += wj * src[t];
            }
        }
#pragma unroll
        for (int u = 0; u < GM_TPT; u++) {
            int t = t0 + u * blockDim.x;
            if (t < L) out[(size_t)e * L + t] = acc[u] + ((bias && t == L - 1) ? bias[e] : 0ULL);
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
         int L) {
    extern __shared__ i64 s_w[];   // [K][MM_TN]
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

static inline int col_chunks(int L, int per_block) {
    int c = (L + per_block - 1) / per_block;
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
        int grid = (int)((words / 2 + 255) / 256 < 148L * 8 ? (words / 2 + 255) / 256 : 148L * 8);
        k_flat<0><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, b, words);
    } else
        k_rows<0><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, st>>>(out, a, b, ia, ib, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_lin_sub(u64* out, const u64* a, const u64* b, const int* ia, const int* ib, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_sub: bad arguments");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!ia && !ib && LIN_ALIGNED16(out, a, b)) {
        long words = E * (long)L;
        int grid = (int)((words / 2 + 255) / 256 < 148L * 8 ? (words / 2 + 255) / 256 : 148L * 8);
        k_flat<1><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, b, words);
    } else
        k_rows<1><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, st>>>(out, a, b, ia, ib, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

int fhe_lin_neg(u64* out, const u64* a, const int* ia, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_neg: bad arguments");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!ia && LIN_ALIGNED16(out, a, a)) {
        long words = E * (long)L;
        int grid = (int)((words / 2 + 255) / 256 < 148L * 8 ? (words / 2 + 255) / 256 : 148L * 8);
        k_flat<2><<<grid < 1 ? 1 : grid, 256, 0, st>>>(out, a, nullptr, words);
    } else
        k_rows<2><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, st>>>(out, a, nullptr, ia, nullptr, nullptr, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = a[ia[e]] with encoded plaintext pt[e] added to the body (add_eint_int / sub_eint_int)
int fhe_lin_add_plain(u64* out, const u64* a, const int* ia, const u64* pt, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_add_plain: bad arguments");
    if (E == 0) return 0;
    k_rows<3><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, (cudaStream_t)stream>>>(out, a, nullptr, ia, nullptr, pt, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = pt[e] - a[ia[e]]   (sub_int_eint)
int fhe_lin_plain_sub(u64* out, const u64* a, const int* ia, const u64* pt, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_plain_sub: bad arguments");
    if (E == 0) return 0;
    k_rows<5><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, (cudaStream_t)stream>>>(out, a, nullptr, ia, nullptr, pt, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

// out[e] = a[ia[e]] * c[e]   (mul_eint_int; c signed, wrapping)
int fhe_lin_mul_clear(u64* out, const u64* a, const int* ia, const i64* c, long E, int L, void* stream) {
    if (!LIN_ARGS_OK(E, L)) FHE_FAIL(2, "lin_mul_clear: bad arguments");
    if (E == 0) return 0;
    k_rows<4><<<dim3((unsigned)E, col_chunks(L, 1024)), 256, 0, (cudaStream_t)stream>>>(
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
    size_t smem = (size_t)K * MM_TN * sizeof(i64);
    if (smem > 48 * 1024) FHE_FAIL(3, "matmul_ct_clear: K=%d too large for the tiled kernel (use gather_mac)", K);
    dim3 grid((unsigned)Mr, (Nc + MM_TN - 1) / MM_TN, col_chunks(L, 256));
    k_matmul<<<grid, 256, smem, (cudaStream_t)stream>>>(out, x, W, K, Nc, L);
    FHE_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
