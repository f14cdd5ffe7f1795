// Grouped bootstrap for the single-CTA plans (N = 2048, k = 1): one CTA per SM carries a GROUP of G ciphertexts ("slots").
//
// k_pbs runs G independent CTAs per SM for these sizes, which keeps its twiddle tables in shared memory (tensor memory is
// allocated per CTA) and lets every CTA stream its own copy of GGSW_i from L2.  Here
//
//   * the CTA has G * NT threads; slot `grp` is run by threads [grp*NT, (grp+1)*NT) with the phases, the FFT factorisation
//     and the Fourier key layout of k_pbs, synchronised by its own named barrier (bar.sync grp+1, NT), so the slots drift
//     out of phase like independent CTAs do and overlap each other's FP64- and shared-memory-bound phases;
//   * one CTA per SM owns all of tensor memory: the mid-pass twiddles, the stage-1 twiddles of a thread's item and the
//     item's own accumulator words (read in P1, updated in P5 by the same thread; the shared array remains the copy the
//     ROTATED operand is read from) live in its tensor-memory lane (twiddles fetched 4 at a time: the kernel runs at 128
//     registers), which takes a quarter of the traffic off the shared-memory pipe -- the most loaded resource of these
//     shapes (ncu: shared-memory wavefronts 56 %, issue slots 61 %, FP64 pipe 47 % of peak before the accumulator moved);
//   * RING = true (the north star's "each GGSW row staged once ... reused across the batch"): GGSW_i lives in a ring of
//     (k+1)*J slices of BUF double2, every slice ONE bulk asynchronous copy global -> shared (cp.async.bulk, completion on
//     the slice's mbarrier).  A warp that has taken its values of a slice counts itself off on the slice's counter; the
//     LAST warp of the last slot re-arms the barrier and issues the copy of the same slice of GGSW_{i+1} -- nobody blocks on
//     a refill.  MEASURED SLOWER than streaming (N = 2048, one level: 59.7 k vs 62.0 k PBS/s, k_pbs 48.5 k;
//     profiles/r02_group_ablation.log): the staged values are read through the shared-memory pipe, which is the scarce
//     resource here, while L2 -> SM bandwidth is not (47 MiB key, L2-resident, 8 B/clk/SM).  So RING = false is the
//     default (fhe_pbs_set_group(2) / FHE_PBS_GROUP=2 select the ring); it also serves the shapes whose GGSW_i does not
//     fit next to two ciphertexts (two levels) and the extended bootstraps of the wide-integer path (EXT).
//
// Same floating-point operations in the same order as k_pbs: the outputs are bit-identical (fhe_pbs_set_group switches at
// run time, tests/test_gpu_kernels.py).  N = 4096 cannot be grouped: one ciphertext already takes 192 KiB.
//
// Replaces memref_batched_bootstrap_lwe_u64 of the absent concrete-compiler runtime
// (concrete/numpy/compilation/server.py:286, concrete/numpy/mlir/node_converter.py:1129-1208).
#pragma once

__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned smem_atom_inc_acq_rel(unsigned* p) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
    return old;
}

template <class P, int G>
struct PbsGroup {
    static constexpr int NTC = P::NT * G;                       // threads per CTA
    // tensor-memory columns per thread: mid-pass twiddles + the R1 stage-1 twiddles of the thread's one item
    // ... and its 2 * R1 own accumulator words (the same thread reads them in P1 and updates them in P5)
    static constexpr int TM_T = P::TM_A + P::TM_B + 4 * P::R1 + 4 * P::R1;
    static constexpr int TM_NEED = TM_T * ((NTC + 127) / 128);
    static constexpr int TM_COLS = TM_NEED <= 32 ? 32 : TM_NEED <= 64 ? 64 : TM_NEED <= 128 ? 128 : TM_NEED <= 256 ? 256 : 512;
    static constexpr int MAX_SLICES = 8;                        // (k+1) * J with k = 1, level <= 4
    static_assert(P::C == 1 && P::KMAX == 1 && TM_NEED <= 512 && NTC <= 1024, "grouped bootstrap: single-CTA plan, k = 1");
    static_assert(P::NT == 2 * (P::BUF / P::RF), "fused pass: thread pair (w, w + NG) shares butterfly group w");
    static_assert((2 << P::LOGQ) == P::NT, "one stage-1 item per thread (its twiddles are per-thread constants)");
    static size_t smem(const PbsGeom& g, bool ring) {
        return (size_t)G * ((size_t)2 * P::S * 8 + (size_t)g.J * P::BUF * 16) + (ring ? (size_t)2 * g.J * P::BUF * 16 : 0);
    }
};

// RING = false: no staging (shapes whose GGSW_i does not fit next to G ciphertexts): every slot streams GGSW_i from L2 as
// k_pbs does; what is left of the design is one CTA per SM with G independent slots and the twiddles in tensor memory
// EXT (never staged): the extended bootstraps of the wide-integer lookup as in k_pbs<P, true> -- one GGSW list per ciphertext
// (bsk_ct_stride > 0) and / or a multi-output bootstrap (g.theta, g.nout)
template <class P, int G, bool RING, bool EXT>
__global__ void __launch_bounds__(P::NT * G, 1)
k_pbs_group(u64* __restrict__ out, const u64* __restrict__ in, long B, const double2* __restrict__ bskF_all,
            const u64* __restrict__ luts, const int* __restrict__ lut_idx, const PbsGeom g,
            const double2* __restrict__ twT, const double2* __restrict__ tw_mid, long long* prof, size_t bsk_ct_stride) {
    static_assert(!(RING && EXT), "extended bootstraps stream their GGSWs");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using GP = PbsGroup<P, G>;
    constexpr int NT = P::NT, R1 = P::R1, RF = P::RF, BUF = P::BUF, S = P::S, N = P::N;
    constexpr int NG = BUF / RF;                                  // butterfly groups of the fused pass
    constexpr unsigned BUF_BYTES = (unsigned)(sizeof(double2) << P::LOGBUF);
    constexpr unsigned WARPS_PER_SLICE = NG / 32;                 // warps of one ciphertext slot that read a given slice
    const int J = g.J, level = g.level, base_log = g.base_log;
    const int tid = (int)threadIdx.x % NT, grp = (int)threadIdx.x / NT;
    const int bar_id = 1 + grp;
    const int nslices = 2 * J;
    __shared__ __align__(8) u64 s_full[GP::MAX_SLICES];           // slice s of the current GGSW has landed
    __shared__ unsigned s_cnt[GP::MAX_SLICES];                    // warps that have taken their values of slice s
    __shared__ unsigned s_tmem_base;
    // shared memory: per slot [ACC 2*S u64][work J*BUF double2], then the GGSW ring [2J][BUF]
    const size_t slot_bytes = (size_t)2 * S * 8 + (size_t)J * BUF * 16;
    u64* acc = reinterpret_cast<u64*>(smem_raw + (size_t)grp * slot_bytes);
    double2* work = reinterpret_cast<double2*>(acc + 2 * (size_t)S);
    double2* ring = reinterpret_cast<double2*>(smem_raw + (size_t)G * slot_bytes);
    const size_t ggsw_stride = (size_t)J * 2 * BUF;               // double2 per CMUX iteration
    const double tscale = g.inv_M * 5.42101086242752217e-20;      // 2^-64 / M
    const bool do_prof = (prof != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_last = do_prof ? clock64() : 0;
#define GROUP_PROF(slot)                       \
    if (do_prof) {                             \
        const long long t_now = clock64();     \
        prof[slot] += t_now - t_last;          \
        t_last = t_now;                        \
    }

    // ciphertext of slot `grp` in round r: r * (grid * G) + grp * grid + blockIdx.x -- the first slots of all CTAs fill
    // before the second ones, so a small batch spreads over the SMs; the CTA is active in a round iff its slot 0 is
    const long grid = gridDim.x, per_round = grid * G;
    const long my_rounds = (B > (long)blockIdx.x + grp * grid) ? (B - blockIdx.x - grp * grid + per_round - 1) / per_round : 0;
    const long cta_rounds = (B > (long)blockIdx.x) ? (B - blockIdx.x + per_round - 1) / per_round : 0;
    const long total_iters = cta_rounds * g.n;                    // GGSWs this CTA stages

    // ---- set-up: tensor-memory twiddles, barriers, the first GGSW
    {
        const int warp = (int)threadIdx.x >> 5;
        if (warp == 0) tmem_alloc(&s_tmem_base, GP::TM_COLS);
        if (RING && threadIdx.x == 0) {          // the ring's barriers and counters (RING implies nslices <= MAX_SLICES)
            for (int s = 0; s < nslices && s < GP::MAX_SLICES; s++) {
                mbar_init(&s_full[s], 1);
                s_cnt[s] = 0;
            }
            mbar_fence_init();
        }
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        if (RING && threadIdx.x == 0 && total_iters > 0)
            for (int s = 0; s < nslices; s++) {
                mbar_arm(&s_full[s], BUF_BYTES);
                bulk_copy_from_global(ring + ((size_t)s << P::LOGBUF), bskF_all + ((size_t)s << P::LOGBUF), BUF_BYTES, &s_full[s]);
            }
    }
    TwCtx tw;
    tw.twT = twT;
    tw.sA = tw.sB = nullptr;
    tw.base = s_tmem_base;
    tw.tA = tw.base + ((unsigned)((((int)threadIdx.x >> 5) & 3) * 32) << 16) + (unsigned)(((int)threadIdx.x >> 7) * GP::TM_T);
    tw.tB = tw.tA + P::TM_A;
    tw.tT = tw.tB + P::TM_B;
    tw.tAcc = tw.tT + 4 * P::R1;
    {
        double2 T[R1];
        load_T<P>(T, tw, tid & (P::Q - 1));
        tmem_store_c<R1>(tw.tT, T);
    }
    if constexpr (P::LRA > 0) tw_init_pass<P::LRA, P::LOGL1, P::LOGBUF, NT>(tw.tA, tw_mid, tid);
    if constexpr (P::LRB > 0) tw_init_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT>(tw.tB, tw_mid + P::TWA, tid);
    tmem_wait_st();

    long it = 0;                                                  // GGSWs consumed so far by this slot (== by the CTA)
    for (long r = 0; r < my_rounds; r++) {
        const long ct = r * per_round + grp * grid + blockIdx.x;
        // ciphertext slots of this CTA that are active in this round (slot s is active iff its ciphertext exists)
        unsigned active = 0;
        for (int s = 0; s < G; s++) active += (r * per_round + s * grid + blockIdx.x < B) ? 1u : 0u;
        const unsigned expect = active * WARPS_PER_SLICE;
        const double2* bskF = EXT ? bskF_all + (size_t)ct * bsk_ct_stride : bskF_all;
        const u64* lwe = in + (size_t)ct * (g.n + 1);
        const u64* lut = luts + (size_t)(lut_idx ? lut_idx[ct] : 0) * 2 * N;
        int a_next = EXT ? modswitch_coarse(lwe[0], g.logN, g.theta) : modswitch(lwe[0], g.logN);
        {
            // ---- ACC = X^{-b~} * LUT
            const int a0 = (2 * N - (EXT ? modswitch_coarse(lwe[g.n], g.logN, g.theta) : modswitch(lwe[g.n], g.logN))) & (2 * N - 1);
            for (int w = tid; w < 2 * S; w += NT) {
                const int c = w >> P::LOGS, s = w & (S - 1);
                const int v = (slot_to_coef<P>(s, 0) - a0) & (2 * N - 1);
                const u64 x = lut[(size_t)c * N + (v & (N - 1))];
                acc[w] = (v >> (P::LOGM + 1)) ? (0ULL - x) : x;
            }
        }
        group_sync(bar_id, NT);
        {
            // the thread's own words (item tid: polynomial tid >> LOGQ, slot q) -> its tensor-memory lane; the shared array
            // stays the copy the ROTATED operand is read from
            const u64* accc = acc + ((size_t)(tid >> P::LOGQ) << P::LOGS) + (tid & (P::Q - 1));
            unsigned ow[4 * R1];
#pragma unroll
            for (int hj = 0; hj < 2 * R1; hj++) {
                const u64 v = accc[hj << P::LOGQ];
                ow[2 * hj] = (unsigned)v;
                ow[2 * hj + 1] = (unsigned)(v >> 32);
            }
            TmemIo<4 * R1>::st(tw.tAcc, ow);
            tmem_wait_st();
        }

        for (int i = 0; i < g.n; i++, it++) {
            const int a = a_next;
            a_next = (i + 1 < g.n) ? (EXT ? modswitch_coarse(lwe[i + 1], g.logN, g.theta) : modswitch(lwe[i + 1], g.logN)) : 0;
            if (!RING && i + 1 < g.n && tid < nslices)      // GGSW_{i+1} -> L2 one iteration ahead
                prefetch_l2_bulk(bskF + (size_t)(i + 1) * ggsw_stride + ((size_t)tid << P::LOGBUF), BUF_BYTES);
            // ---- P1: rotate, subtract, decompose, fold + twist, radix-R1 stage
            if (level * base_log <= 30)
                p1_items<P, unsigned, true, true>(acc, work, 0, a, g, tw, tid);
            else
                p1_items<P, u64, true, true>(acc, work, 0, a, g, tw, tid);
            group_sync(bar_id, NT);
            GROUP_PROF(0)
            // ---- P2: forward mid passes of the J digit polynomials
            if constexpr (P::LRA > 0) {
                mid_pass<P::LRA, P::LOGL1, P::LOGBUF, NT, false, true, 4>(work, J, tw.tA, nullptr, 1, tid);
                group_sync(bar_id, NT);
            }
            if constexpr (P::LRB > 0) {
                mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, false, true, 4>(work, J, tw.tB, nullptr, 1, tid);
                group_sync(bar_id, NT);
            }
            GROUP_PROF(2)
            // ---- P3: last forward pass + multiply-accumulate against the staged GGSW_i + first inverse pass.  Thread pair
            // (w, w + NG) shares butterfly group w; thread (w, o) accumulates output polynomial o and reads slices 2*j + o
            {
                const int w = tid % NG, o = tid / NG;
                const unsigned par = (unsigned)(it & 1);
                const long it_next = it + 1;
                const double2* gnext_global = bskF_all + (size_t)((i + 1 < g.n) ? i + 1 : 0) * ggsw_stride;
                double2 r[RF], gv[RF];
                // slice sl of GGSW_i as this thread reads it: the staged copy, or the key itself (L2)
                const double2* gsl = (RING ? ring : bskF + (size_t)i * ggsw_stride) + w;
                if (RING) mbar_wait(&s_full[o], par);
                {
                    const double2* gs = gsl + ((size_t)o << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) {
                        r[m] = make_double2(0.0, 0.0);
                        gv[m] = RING ? gs[m * NG] : __ldg(gs + m * NG);
                    }
                }
#pragma unroll 1
                for (int j = 0; j < J; j++) {
                    double2 x[RF];
                    const double2* base = work + ((size_t)j << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) x[m] = base[SW(w * RF + m)];
                    Dft<RF, false>::run(x);
                    const bool more = (j + 1 < J);
                    const int sl = 2 * j + o;
                    const double2* gn = gsl + ((size_t)(sl + 2) << P::LOGBUF);
                    if (RING && more) mbar_wait(&s_full[sl + 2], par);
#pragma unroll
                    for (int m = 0; m < RF; m++) {
                        const double2 b = gv[m];
                        r[m].x = fma(x[m].x, b.x, r[m].x);
                        r[m].x = fma(-x[m].y, b.y, r[m].x);
                        r[m].y = fma(x[m].x, b.y, r[m].y);
                        r[m].y = fma(x[m].y, b.x, r[m].y);
                        if (more) gv[m] = RING ? gn[m * NG] : __ldg(gn + m * NG);
                    }
                    if constexpr (RING) {
                        // this warp holds its values of slice sl in registers (they were requested one step ahead): count
                        // it off; the last warp of the last slot refills the slice with GGSW_{i+1} (next round: GGSW_0)
                        __syncwarp();
                        if ((tid & 31) == 0) {
                            const unsigned old = smem_atom_inc_acq_rel(&s_cnt[sl]);
                            if (old + 1 == expect) {
                                s_cnt[sl] = 0;
                                if (it_next < total_iters) {
                                    mbar_arm(&s_full[sl], BUF_BYTES);
                                    bulk_copy_from_global(ring + ((size_t)sl << P::LOGBUF), gnext_global + ((size_t)sl << P::LOGBUF),
                                                          BUF_BYTES, &s_full[sl]);
                                }
                            }
                        }
                    }
                }
                group_sync(bar_id, NT);      // the partner thread still reads the buffers the outputs overwrite
                Dft<RF, true>::run(r);
                double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                for (int m = 0; m < RF; m++) base[SW(w * RF + m)] = r[m];
            }
            group_sync(bar_id, NT);
            GROUP_PROF(3)
            // ---- P4: inverse mid passes of the two output polynomials
            if constexpr (P::LRB > 0) {
                mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, true, true, 4>(work, 2, tw.tB, nullptr, 1, tid);
                group_sync(bar_id, NT);
            }
            if constexpr (P::LRA > 0) {
                mid_pass<P::LRA, P::LOGL1, P::LOGBUF, NT, true, true, 4>(work, 2, tw.tA, nullptr, 1, tid);
                group_sync(bar_id, NT);
            }
            GROUP_PROF(4)
            // ---- P5: inverse radix-R1 stage, untwist, round, ACC +=
#pragma unroll 1
            for (int itm = tid; itm < (2 << P::LOGQ); itm += NT) {
                const int o = itm >> P::LOGQ, q = itm & (P::Q - 1);
                double2 T[R1];
                load_T<P, true>(T, tw, q);
                double2 z[R1];
                stage1_gather<P>(work + ((size_t)o << P::LOGBUF), q, T, g, z);
                u64* acco = acc + ((size_t)o << P::LOGS);
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    unsigned ow[2 * R1];
                    TmemIo<2 * R1>::ld(tw.tAcc + (unsigned)(half * 2 * R1), ow);
                    tmem_wait_ld();
#pragma unroll
                    for (int j1 = 0; j1 < R1; j1++) {
                        const u64 v = (((u64)ow[2 * j1 + 1] << 32) | (u64)ow[2 * j1]) + f2torus_frac(half ? z[j1].y : z[j1].x, tscale);
                        ow[2 * j1] = (unsigned)v;
                        ow[2 * j1 + 1] = (unsigned)(v >> 32);
                        acco[((half * R1 + j1) << P::LOGQ) + q] = v;          // the copy the next rotation reads
                    }
                    TmemIo<2 * R1>::st(tw.tAcc + (unsigned)(half * 2 * R1), ow);
                }
            }
            tmem_wait_st();
            group_sync(bar_id, NT);
            GROUP_PROF(6)
        }
        // ---- sample extract coefficient 0 (extended: coefficients 0..nout-1)
        {
            const int nout = EXT ? g.nout : 1;
            for (int u = 0; u < nout; u++) {
                // coefficient u of the product: mask word t is A[u - t] (t <= u), -A[N + u - t] (t > u)
                u64* o = out + ((size_t)ct * nout + u) * ((size_t)N + 1);
                for (int w = tid; w < 2 * S; w += NT) {
                    const int c = w >> P::LOGS, s = w & (S - 1);
                    const int j = slot_to_coef<P>(s, 0);
                    const u64 v = acc[w];
                    if (c < 1) {
                        if (j <= u) o[u - j] = v;
                        else o[N + u - j] = 0ULL - v;
                    } else if (j == u)
                        o[N] = v;
                }
            }
        }
        group_sync(bar_id, NT);              // the next round re-initialises the accumulator
        GROUP_PROF(8)
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (((int)threadIdx.x >> 5) == 0) tmem_dealloc(s_tmem_base, GP::TM_COLS);
#undef GROUP_PROF
}
