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

static std::vector<PbsPlanEntry> all_plans() {
    std::vector<PbsPlanEntry> v;
    int n = 0;
    const PbsPlanEntry* p = pbs_plans_c(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_b(&n);
    v.insert(v.end(), p, p + n);
    p = pbs_plans_a(&n);
    v.insert(v.end(), p, p + n);
    return v;
}

static size_t plan_smem(const PbsPlanMeta& m, int k, int level) {
    const long N = 2L << m.logM;
    const long buf = (m.C > 1) ? ((1L << m.logM) / m.C) : (1L << m.logM);
    return (size_t)(k + 1) * (N / m.C) * 8 + (size_t)(k + 1) * level * buf * 16 + (size_t)m.smem_fixed;
}

// smallest cluster whose on-chip layout fits; index into all_plans() or -1
static int choose_plan(int N, int k, int level) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const int logM = ilog2_exact(N) - 1;
    // experiment knob: FHE_PBS_SMALL_L=1 prefers the plans with 1024-point local transforms (twice the
    // cluster size, two clusters resident per SM so that one's exchange phases overlap the other's math)
    const char* e = getenv("FHE_PBS_SMALL_L");
    const bool small_l = e && atoi(e) != 0;
    int best = -1;
    for (int C = 1; C <= 16; C *= 2)
        for (size_t i = 0; i < plans.size(); i++) {
            const PbsPlanMeta& m = plans[i].meta;
            if (m.C != C || m.logM != logM) continue;
            if (k > m.kmax || (k <= 1 && m.kmax > 1)) continue;
            if (plan_smem(m, k, level) > PBS_SMEM_MAX) continue;
            if (best < 0) best = (int)i;
            else if (small_l && plans[best].meta.C * 2 == C && plan_smem(m, k, level) <= 115712 &&
                     m.lra + m.lrb + m.lrf == 10)
                best = (int)i;
        }
    return best;
}

struct PbsTables {
    double2* twT;
    double2* tw_mid;
};
static std::mutex g_plan_mu;
static std::map<std::pair<int, int>, PbsTables> g_tables;   // (device, plan index)

static int get_plan(int n, int k, int N, int level, int base_log, const PbsPlanEntry** entry, PbsLaunchArgs* a) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    const int logN = ilog2_exact(N);
    if (logN < 8 || logN > 15) FHE_FAIL(2, "pbs: polynomial size %d unsupported (256..32768)", N);
    if (k < 1 || k > 3) FHE_FAIL(2, "pbs: glwe dimension %d unsupported (1..3)", k);
    if (level < 1 || level > 4 || base_log < 1 || level * base_log > 52)
        FHE_FAIL(2, "pbs: decomposition level=%d base_log=%d unsupported", level, base_log);
    const int pi = choose_plan(N, k, level);
    if (pi < 0) FHE_FAIL(3, "pbs: N=%d k=%d level=%d does not fit the on-chip layout", N, k, level);
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
            FHE_CUDA(cudaMalloc(&tb.twT, sizeof(double2) * tT.size()));
            FHE_CUDA(cudaMalloc(&tb.tw_mid, sizeof(double2) * tm.size()));
            FHE_CUDA(cudaMemcpy(tb.twT, tT.data(), sizeof(double2) * tT.size(), cudaMemcpyHostToDevice));
            FHE_CUDA(cudaMemcpy(tb.tw_mid, tm.data(), sizeof(double2) * tm.size(), cudaMemcpyHostToDevice));
            g_tables[{dev, pi}] = tb;
        } else
            tb = it->second;
    }
    PbsGeom& g = a->g;
    g.n = n; g.k = k; g.N = N; g.level = level; g.base_log = base_log;
    g.logN = logN; g.logM = logN - 1;
    g.J = (k + 1) * level;
    g.inv_M = 1.0 / (double)(N / 2);
    {
        const char* e = getenv("FHE_PBS_STAGGER");
        g.stagger = e ? atoi(e) : 0;
        e = getenv("FHE_PBS_DBG");
        g.dbg = e ? atoi(e) : 0;
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

static long long* g_prof_buf = nullptr;

extern "C" {

// development aid: device buffer of >= 16 int64 that k_pbs accumulates per-phase clock64() deltas
// into (thread 0 of CTA 0); pass NULL to switch it off
int fhe_pbs_set_profile(long long* dev_buf) {
    g_prof_buf = dev_buf;
    return 0;
}

// cluster size (CTAs per ciphertext) the on-chip layout uses for these parameters, <0 if unsupported
int fhe_pbs_cluster_size(int N, int k, int level) {
    static const std::vector<PbsPlanEntry> plans = all_plans();
    if (ilog2_exact(N) < 8 || ilog2_exact(N) > 15 || k < 1 || k > 3 || level < 1 || level > 4) return -1;
    const int pi = choose_plan(N, k, level);
    return pi < 0 ? -1 : plans[pi].meta.C;
}

// number of double2 elements of the Fourier BSK
long fhe_bsk_fourier_elems(int n, int k, int N, int level) {
    return (long)n * (k + 1) * level * (k + 1) * (N / 2);
}

int fhe_bsk_to_fourier(double2* bsk_f, const u64* bsk_std, int n, int k, int N, int level, int base_log,
                       void* stream) {
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.outF = bsk_f;
    a.bsk = bsk_std;
    a.npoly = (long)n * level * (k + 1) * (k + 1);
    return e->bsk(a);
}

// out [B][kN+1], in [B][n+1] (small key), luts [T][(k+1)N] accumulators, lut_idx [B] or null
int fhe_pbs_batch(u64* out, const u64* in, long B, const double2* bsk_f, const u64* luts,
                  const int* lut_idx, int n, int k, int N, int level, int base_log, void* stream) {
    if (B < 0) FHE_FAIL(2, "pbs_batch: negative batch");
    if (B == 0) return 0;
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(n, k, N, level, base_log, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.out = out; a.in = in; a.B = B; a.bskF = bsk_f; a.luts = luts; a.lut_idx = lut_idx; a.prof = g_prof_buf;
    a.g.prof = g_prof_buf;
    return e->pbs(a);
}

// test hook: out[p] = a_small[p] * b_torus[p] in Z_2^64[X]/(X^N+1) through the fp64 FFT path
int fhe_debug_polymul(u64* out, const i64* a_small, const u64* b_torus, long npoly, int N, int k,
                      int level, void* stream) {
    PbsLaunchArgs a = {};
    const PbsPlanEntry* e = nullptr;
    int rc = get_plan(1, k, N, level, 8, &e, &a);
    if (rc) return rc;
    a.st = (cudaStream_t)stream;
    a.pm_out = out; a.pm_a = a_small; a.pm_b = b_torus; a.npoly = npoly;
    return e->polymul(a);
}

}  // extern "C"
