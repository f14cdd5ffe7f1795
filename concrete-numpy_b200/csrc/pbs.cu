// Host side of the batched programmable bootstrap: plan selection, twiddle tables, C-ABI entry points.
// The kernels live in pbs_kernels.cuh and are instantiated in pbs_inst_{a,b,c}.cu.
#include <math.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "pbs_kernels.cuh"

const PbsPlanEntry* pbs_plans_a(int* count);
const PbsPlanEntry* pbs_plans_b(int* count);
const PbsPlanEntry* pbs_plans_c(int* count);
const PbsPlanEntry* pbs_plans_d(int* count);
const PbsPlanEntry* pbs_plans_e(int* count);

static std::vector<PbsPlanEntry> all_plans() {
    std::vector<PbsPlanEntry> v;
    int n = 0;
    const PbsPlanEntry* p = pbs_plans_c(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_b(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_a(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_d(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_e(&n);
    v.insert(v.end(), p, p + n);
    return v;
}

static size_t plan_smem(const PbsPlanMeta& m, int k, int level) {
    const long N = 2L << m.logM;
    const long buf = (m.C > 1) ? ((1L << m.logM) / m.C) : (1L << m.logM);
    return (size_t)(k + 1) * (N / m.C) * 8 + (size_t)(k + 1) * level * buf * 16 + (size_t)m.smem_fixed;
}

// every plan whose on-chip layout fits (N, k, level), smallest cluster first: entry 0 is the throughput
// plan (most ciphertexts in flight per GPU), the later ones spread one ciphertext over more SMs and
// are used for small batches (latency plans).  Values are indices into all_plans().
static std::vector<int> fitting_plans(int N, int k, int level) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    // development knob, read once: FHE_PBS_ACCTM=0 keeps the accumulator in shared memory (A/B measurements)
    static const bool use_acctm = [] { const char* e = getenv("FHE_PBS_ACCTM"); return !(e && e[0] == '0'); }();
    static const bool no_256x2 = [] { const char* e = getenv("FHE_PBS_256X2"); return e && e[0] == '0'; }();
    std::vector<int> out;
    const int logN = ilog2_exact(N);
    if (logN < 8 || logN > 15 || k < 1 || k > 3 || level < 1 || level > 4) return out;
    const int logM = logN - 1;
    // one plan per cluster size: the one that keeps the most threads resident per SM for this (k, level)
    // (shared memory, the 64K-register file at 255 registers per thread, tensor-memory columns); ties go to the
    // earlier entry
    for (int C = 1; C <= 16; C *= 2) {
        int best = -1;
        long best_threads = 0;
        for (size_t i = 0; i < plans.size(); i++) {
            const PbsPlanMeta& m = plans[i].meta;
            if (m.C != C || m.logM != logM) continue;
            if (k > m.kmax || (k <= 1 && m.kmax > 1)) continue;
            const size_t smem = plan_smem(m, k, level);
            if (smem > PBS_SMEM_MAX) continue;
            long ctas = (long)(233472 / (smem + 1024));
            long by_regs = 256 / m.NT > 0 ? 256 / m.NT : 1;      // 255 registers per thread unless __launch_bounds__ says more CTAs
            if (by_regs < m.minb) by_regs = m.minb;
            if (ctas > by_regs) ctas = by_regs;
            if (m.minb == 1 && ctas > 1) ctas = 1;            // tensor-memory twiddle plans are built for one CTA per SM
            if (ctas < 1) ctas = 1;
            const long threads = ctas * m.NT;
            if (m.acctm && !use_acctm) continue;
            // pipelined plans (acctm == 2) exist for what k_pbs_pipe runs: k = 1, two levels
            if (m.acctm == 2 && !(k == 1 && level == 2)) continue;
            if (no_256x2 && m.C == 1 && m.NT == 256 && m.minb == 2) continue;
            // at equal residency the tensor-memory accumulator plan wins (profiles/r02_*: no remote loads in P1)
            const bool tie = (threads == best_threads) && best >= 0;
            // ties: more registers per thread first (lower minb), then the tensor-memory accumulator plan
            // a pipelined plan that applies (checked above) beats the generic plans of its cluster size: it removes the
            // cluster barriers, which matter most where a CTA has little work per CMUX
            const bool pipe_wins = best >= 0 && m.acctm == 2 && plans[best].meta.acctm != 2;
            const bool pipe_holds = best >= 0 && plans[best].meta.acctm == 2 && m.acctm != 2;
            if (pipe_holds) continue;
            if (pipe_wins || threads > best_threads || (tie && m.minb < plans[best].meta.minb) ||
                (tie && m.minb == plans[best].meta.minb && m.acctm && !plans[best].meta.acctm)) {
                best_threads = threads;
                best = (int)i;
            }
        }
        if (best >= 0) out.push_back(best);
    }
    return out;
}

struct PbsTables {
    double2* twT;
    double2* tw_mid;
};
static std::mutex g_plan_mu;
static std::map<std::pair<int, int>, PbsTables> g_tables;   // (device, plan index)

static int get_plan(int n, int k, int N, int level, int base_log, int variant, const PbsPlanEntry** entry,
                    PbsLaunchArgs* a) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const int logN = ilog2_exact(N);
    if (logN < 8 || logN > 15) FHE_FAIL(2, "pbs: polynomial size %d unsupported (256..32768)", N);
    if (k < 1 || k > 3) FHE_FAIL(2, "pbs: glwe dimension %d unsupported (1..3)", k);
    if (level < 1 || level > 4 || base_log < 1 || level * base_log > 52)
        FHE_FAIL(2, "pbs: decomposition level=%d base_log=%d unsupported", level, base_log);
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (fit.empty()) FHE_FAIL(3, "pbs: N=%d k=%d level=%d does not fit the on-chip layout", N, k, level);
    if (variant < 0 || variant >= (int)fit.size())
        FHE_FAIL(2, "pbs: plan variant %d out of range (0..%d)", variant, (int)fit.size() - 1);
    const int pi = fit[variant];
    const PbsPlanMeta& m = plans[pi].meta;
    int dev = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    const int R1 = 1 << m.logR1, logL1 = m.lra + m.lrb + m.lrf, L1 = 1 << logL1;
    PbsTables tb;
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        auto it = g_tables.find({dev, pi});
        if (it == g_tables.end()) {
            std::vector<double2> tT((size_t)R1 * L1), tm;
            for (int t1 = 0; t1 < R1; t1++)
                for (int j2 = 0; j2 < L1; j2++) {
                    // exp(i*pi*j2*(1-4*t1)/N); reduce the integer angle mod 2N first
                    long num = ((long)j2 * (1 - 4 * t1)) % (2L * N);
                    long double ang = M_PIl * (long double)num / (long double)N;
                    tT[(size_t)t1 * L1 + j2] = make_double2((double)cosl(ang), (double)sinl(ang));
                }
            auto add_pass = [&](int lr, int logLc) {
                if (!lr) return;
                const int R = 1 << lr, S = 1 << (logLc - lr);
                for (int mm = 1; mm < R; mm++)
                    for (int q = 0; q < S; q++) {
                        long double ang = -2.0L * M_PIl * (long double)((long)q * mm) / (long double)(1L << logLc);
                        tm.push_back(make_double2((double)cosl(ang), (double)sinl(ang)));
                    }
            };
            add_pass(m.lra, logL1);
            add_pass(m.lrb, logL1 - m.lra);
            if (tm.empty()) tm.push_back(make_double2(1.0, 0.0));
            tb.twT = tb.tw_mid = nullptr;
            cudaError_t ce = cudaMalloc(&tb.twT, sizeof(double2) * tT.size());
            if (ce == cudaSuccess) ce = cudaMalloc(&tb.tw_mid, sizeof(double2) * tm.size());
            if (ce == cudaSuccess) ce = cudaMemcpy(tb.twT, tT.data(), sizeof(double2) * tT.size(), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = cudaMemcpy(tb.tw_mid, tm.data(), sizeof(double2) * tm.size(), cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) {                       // nothing half-built stays behind
                if (tb.twT) cudaFree(tb.twT);
                if (tb.tw_mid) cudaFree(tb.tw_mid);
                FHE_FAIL(1000 + (int)ce, "pbs: twiddle tables: %s", cudaGetErrorString(ce));
            }
            g_tables[{dev, pi}] = tb;
        } else
            tb = it->second;
    }
    PbsGeom& g = a->g;
    g.n = n; g.k = k; g.N = N; g.level = level; g.base_log = base_log;
    g.logN = logN; g.logM = logN - 1;
    g.J = (k + 1) * level;
    g.inv_M = 1.0 / (double)(N / 2);
    g.theta = 0;
    g.nout = 1;
    {
        // development knobs (ablation bits, start-up stagger): read ONCE per process, 0 unless set before the first launch
        static const int env_stagger = [] { const char* e = getenv("FHE_PBS_STAGGER"); return e ? atoi(e) : 0; }();
        static const int env_dbg = [] { const char* e = getenv("FHE_PBS_DBG"); return e ? atoi(e) : 0; }();
        g.stagger = env_stagger;
        g.dbg = env_dbg;
    }
    for (int j1 = 0; j1 < 16; j1++) {
        long double ang = M_PIl * (long double)j1 / (2.0L * R1);
        g.cj[j1] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    a->twT = tb.twT;
    a->tw_mid = tb.tw_mid;
    *entry = &plans[pi];
    return 0;
}

// Wave-alignment counters of the pipelined bootstrap: a small per-device pool, one zeroed slot per launch (launches on
// different streams may overlap; a shared counter would only cost them their alignment, never correctness)
static int g_pbs_wave = [] { const char* e = getenv("FHE_PBS_WAVE"); return (e && e[0] == '0') ? 0 : 1; }();
int pbs_wave_counter(cudaStream_t st, unsigned** ctr) {
    *ctr = nullptr;
    if (!g_pbs_wave) return 0;
    constexpr int SLOTS = 64;
    static std::map<int, unsigned*> pools;
    static std::map<int, unsigned> next;
    int dev = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    unsigned* slot = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        auto it = pools.find(dev);
        if (it == pools.end()) {
            unsigned* p = nullptr;
            FHE_CUDA(cudaMalloc(&p, SLOTS * 64));            // one 64-byte line per slot
            it = pools.emplace(dev, p).first;
            next[dev] = 0;
        }
        slot = it->second + 16 * (next[dev]++ % SLOTS);
    }
    FHE_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned), st));
    *ctr = slot;
    return 0;
}

static long long* g_prof_buf = nullptr;
int g_pbs_pipeline = [] { const char* e = getenv("FHE_PBS_PIPE"); return (e && e[0] == '0') ? 0 : 1; }();
int g_pbs_group = [] { const char* e = getenv("FHE_PBS_GROUP"); return e ? atoi(e) : 1; }();

// position of natural frequency t of the M-point spectrum inside one (j, o) slice [C][RF][BUF/RF] of a plan
static long plan_position(const PbsPlanMeta& m, long t) {
    const int R1 = 1 << m.logR1, RA = 1 << m.lra, RB = 1 << m.lrb, RF = 1 << m.lrf;
    const long L1 = (long)RA * RB * RF;
    const long t1 = t % R1, t2 = t / R1;
    const long mA = t2 % RA, mB = (t2 / RA) % RB, mF = t2 / ((long)RA * RB);
    const long pl = mA * (L1 / RA) + mB * RF + mF;             // DIF: the first pass fixes the lowest digit
    const long buf = (m.C > 1) ? L1 : ((long)R1 * L1);
    const long rank = (m.C > 1) ? t1 : 0, local = (m.C > 1) ? pl : t1 * L1 + pl;
    const long w = local / RF, r = local % RF;
    return rank * buf + r * (buf / RF) + w;
}

__global__ void k_bsk_relayout(double2* __restrict__ dst, const double2* __restrict__ src,
                               const int* __restrict__ map /* dst position -> src position */, long nslices, int M) {
    for (long s = blockIdx.y; s < nslices; s += gridDim.y)
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < M; p += gridDim.x * blockDim.x)
            dst[s * M + p] = src[s * M + map[p]];
}

extern "C" {

// development aid: device buffer of >= 16 int64 that k_pbs accumulates per-phase clock64() deltas
// into (thread 0 of CTA 0); pass NULL to switch it off
int fhe_pbs_set_profile(long long* dev_buf) {
    g_prof_buf = dev_buf;
    return 0;
}

// test / measurement hook: 0 makes the pipelined plans run their generic kernel (k_pbs) instead of k_pbs_pipe; returns
// the previous setting.  The two kernels are bit-identical by construction (tests/test_gpu_kernels.py).
int fhe_pbs_set_pipeline(int enabled) {
    const int prev = g_pbs_pipeline;
    g_pbs_pipeline = enabled ? 1 : 0;
    return prev;
}

// test / measurement hook: 0 makes the grouped plans (k_pbs_group: several ciphertexts per CTA, GGSW staged once per SM)
// run their generic kernel (k_pbs) instead; returns the previous setting.  Bit-identical by construction.
int fhe_pbs_set_group(int enabled) {
    const int prev = g_pbs_group;
    g_pbs_group = enabled;        // 0: k_pbs, 1: grouped, GGSW from L2 (default), 2: grouped, GGSW staged once per SM
    return prev;
}

// cluster size (CTAs per ciphertext) the on-chip layout uses for these parameters, <0 if unsupported
int fhe_pbs_cluster_size(int N, int k, int level) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    return fit.empty() ? -1 : plans[fit[0]].meta.C;
}

// number of plan variants for these parameters (0 = unsupported); variant 0 is the throughput plan, higher
// variants use larger clusters (lower latency per ciphertext, fewer ciphertexts in flight)
int fhe_pbs_num_variants(int N, int k, int level) { return (int)fitting_plans(N, k, level).size(); }
int fhe_pbs_variant_cluster_size(int N, int k, int level, int variant) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (variant < 0 || variant >= (int)fit.size()) return -1;
    return plans[fit[variant]].meta.C;
}

// plan parameters of a variant: meta[0..7] = C, log2 R1, log2 RA, log2 RB, log2 RF, threads, min CTAs/SM, k max
int fhe_pbs_variant_plan(int N, int k, int level, int variant, int* meta) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (variant < 0 || variant >= (int)fit.size() || !meta) return -1;
    const PbsPlanMeta& m = plans[fit[variant]].meta;
    const int v[8] = {m.C, m.logR1, m.lra, m.lrb, m.lrf, m.NT, m.minb, m.kmax};
    for (int i = 0; i < 8; i++) meta[i] = v[i];
    return 0;
}

// 1 if two-level k = 1 bootstraps of this variant run the pipelined kernel (k_pbs_pipe: per-buffer mbarriers instead of
// cluster barriers), 0 if not, < 0 if the variant does not exist
int fhe_pbs_variant_pipelined(int N, int k, int level, int variant) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (variant < 0 || variant >= (int)fit.size()) return -1;
    return plans[fit[variant]].meta.acctm == 2 ? 1 : 0;
}

// 1 if k = 1 bootstraps of this variant run the grouped kernel (k_pbs_group: one CTA per SM with several ciphertext slots --
// also the lowest-latency plan of its size: a lone ciphertext has the SM and its tensor memory to itself), 0 if not
int fhe_pbs_variant_grouped(int N, int k, int level, int variant) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (variant < 0 || variant >= (int)fit.size()) return -1;
    const PbsPlanMeta& m = plans[fit[variant]].meta;
    if (m.group <= 1 || k != 1 || g_pbs_group == 0) return 0;
    const size_t M = (size_t)1 << m.logM;
    const size_t smem = (size_t)m.group * ((size_t)(k + 1) * 2 * M * 8 + (size_t)(k + 1) * level * M * 16);
    return smem + 2048 <= PBS_SMEM_MAX ? 1 : 0;
}

// how many ciphertexts variant `variant` bootstraps concurrently on the current device (one wave)
int fhe_pbs_variant_capacity(int N, int k, int level, int variant) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (variant < 0 || variant >= (int)fit.size()) return -1;
    PbsGeom g = {};
    g.k = k; g.level = level; g.N = N; g.J = (k + 1) * level;
    return plans[fit[variant]].capacity(g);
}

// Fourier BSK of variant `src_variant` -> layout of `dst_variant` (same spectrum, different order)
int fhe_bsk_relayout(double2* dst, const double2* src, int n, int k, int N, int level, int src_variant,
                     int dst_variant, void* stream) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const std::vector<int> fit = fitting_plans(N, k, level);
    if (src_variant < 0 || dst_variant < 0 || src_variant >= (int)fit.size() || dst_variant >= (int)fit.size())
        FHE_FAIL(2, "bsk_relayout: variant out of range");
    const int M = N / 2;
    std::vector<long> inv_src(M);
    std::vector<int> map(M);
    for (long t = 0; t < M; t++) inv_src[t] = plan_position(plans[fit[src_variant]].meta, t);
    for (long t = 0; t < M; t++) map[plan_position(plans[fit[dst_variant]].meta, t)] = (int)inv_src[t];
    int* d_map = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    FHE_CUDA(fhe_malloc_async(reinterpret_cast<void**>(&d_map), sizeof(int) * M, st));
    FHE_CUDA(cudaMemcpyAsync(d_map, map.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
    FHE_CUDA(cudaStreamSynchronize(st));     // `map` is a host temporary
    const long nslices = (long)n * (k + 1) * level * (k + 1);
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)(nslices < 4096 ? nslices : 4096));
    k_bsk_relayout<<<grid, 256, 0, st>>>(dst, src, d_map, nslices, M);
    FHE_CHECK_LAUNCH();
    FHE_CUDA(cudaFreeAsync(d_map, st));
    return 0;
}

// number of double2 elements of the Fourier BSK
long fhe_bsk_fourier_elems(int n, int k, int N, int level) {
    return (long)n * (k + 1) * level * (k + 1) * (N / 2);
}

int fhe_bsk_to_fourier_variant(double2* bsk_f, const u64* bsk_std, int n, int k, int N, int level, int base_log,
                               int variant, void* stream);
int fhe_bsk_to_fourier(double2* bsk_f, const u64* bsk_std, int n, int k, int N, int level, int base_log,
                       void* stream) {
    return fhe_bsk_to_fourier_variant(bsk_f, bsk_std, n, k, N, level, base_log, 0, stream);
}
int fhe_bsk_to_fourier_variant(double2* bsk_f, const u64* bsk_std, int n, int k, int N, int level, int base_log,
                               int variant, void* stream) {
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, variant, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.outF = bsk_f;
    a.bsk = bsk_std;
    a.npoly = (long)n * level * (k + 1) * (k + 1);
    return e->bsk(a);
}

// out [B][kN+1], in [B][n+1] (small key), luts [T][(k+1)N] accumulators, lut_idx [B] or null;
// bsk_f must be in the layout of `variant` (fhe_bsk_to_fourier produces variant 0)
int fhe_pbs_batch_variant(u64* out, const u64* in, long B, const double2* bsk_f, const u64* luts,
                          const int* lut_idx, int n, int k, int N, int level, int base_log, int variant,
                          void* stream) {
    if (B < 0) FHE_FAIL(2, "pbs_batch: negative batch");
    if (B == 0) return 0;
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, variant, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.out = out; a.in = in; a.B = B; a.bskF = bsk_f; a.luts = luts; a.lut_idx = lut_idx; a.prof = g_prof_buf;
    a.g.prof = g_prof_buf;
    return e->pbs(a);
}
// Blind rotation with one GGSW list per ciphertext (wide-integer lookup, wide.py): ciphertext e uses the n Fourier
// GGSWs at ggsw_f + e * ct_stride (double2 elements, variant-0 layout); `in` holds the rotation amounts as LWE masks.
int fhe_blind_rotate_batch(u64* out, const u64* in, long B, const double2* ggsw_f, long ct_stride, const u64* luts,
                           const int* lut_idx, int n, int k, int N, int level, int base_log, void* stream) {
    if (B < 0 || ct_stride <= 0) FHE_FAIL(2, "blind_rotate_batch: bad arguments");
    if (B == 0) return 0;
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, 0, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.out = out; a.in = in; a.B = B; a.bskF = ggsw_f; a.luts = luts; a.lut_idx = lut_idx;
    a.bsk_ct_stride = (size_t)ct_stride;
    return e->pbs(a);
}

// Multi-output bootstrap (Chillotti, Ligier, Orfila, Tap, "Improved programmable bootstrapping with larger precision
// and efficient arithmetic circuits for TFHE", 2021, PBSmanyLUT): the modulus switch keeps `theta` zero low bits, the
// accumulators interleave up to 2^theta functions (coefficient i belongs to function i mod 2^theta) and coefficients
// 0..nout-1 of the rotated accumulator are extracted: out [B][nout][kN+1].  One blind rotation for nout results on the
// same input, at the price of 2^theta times the modulus-switch noise.
int fhe_pbs_batch_multi(u64* out, const u64* in, long B, const double2* bsk_f, const u64* luts, const int* lut_idx,
                        int n, int k, int N, int level, int base_log, int theta, int nout, int variant, void* stream) {
    if (B < 0 || theta < 0 || theta > 4 || nout < 1 || nout > (1 << theta)) FHE_FAIL(2, "pbs_batch_multi: bad arguments");
    if (B == 0) return 0;
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, variant, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.out = out; a.in = in; a.B = B; a.bskF = bsk_f; a.luts = luts; a.lut_idx = lut_idx;
    a.g.theta = theta;
    a.g.nout = nout;
    return e->pbs(a);
}

// out[b] = pool[i0[b]] + GGSW[gi[b]] (x) (pool[i1[b]] - pool[i0[b]]); pool [P][(k+1)N], out [B][(k+1)N],
// ggsw_f: Fourier GGSWs in the variant-0 layout of (N, k, level) (fhe_bsk_to_fourier with n = their count)
// plain != 0: every pool entry used is a plaintext polynomial (zero mask), only the bodies are transformed
int fhe_cmux_batch(u64* out, const u64* pool, const int* i0, const int* i1, const double2* ggsw_f, const int* gi,
                   long B, int k, int N, int level, int base_log, int plain, void* stream) {
    if (B < 0) FHE_FAIL(2, "cmux_batch: negative batch");
    if (B == 0) return 0;
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(1, k, N, level, base_log, 0, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.out = out; a.B = B; a.bskF = ggsw_f; a.cm_pool = pool; a.cm_i0 = i0; a.cm_i1 = i1; a.cm_gi = gi; a.cm_plain = plain;
    return e->cmux(a);
}

int fhe_pbs_batch(u64* out, const u64* in, long B, const double2* bsk_f, const u64* luts,
                  const int* lut_idx, int n, int k, int N, int level, int base_log, void* stream) {
    return fhe_pbs_batch_variant(out, in, B, bsk_f, luts, lut_idx, n, k, N, level, base_log, 0, stream);
}

// test hook: out[p] = a_small[p] * b_torus[p] in Z_2^64[X]/(X^N+1) through the fp64 FFT path
int fhe_debug_polymul(u64* out, const i64* a_small, const u64* b_torus, long npoly, int N, int k,
                      int level, void* stream) {
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(1, k, N, level, 8, 0, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.pm_out = out; a.pm_a = a_small; a.pm_b = b_torus; a.npoly = npoly;
    return e->polymul(a);
}

}  // extern "C"
