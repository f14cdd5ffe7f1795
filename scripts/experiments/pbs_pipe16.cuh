// EXPERIMENT, NOT BUILT INTO THE LIBRARY (kept for the record): 16-warp variant of the pipelined
// bootstrap (pbs_pipe.cuh) for N = 32768 on 8-CTA clusters.  Correct (all bootstrap parity tests and the cfg2 golden
// circuit pass) but SLOWER than the 8-warp kernel: 735 vs 901 PBS/s.  Its phase timers show why: the two halves run
// their forward mid passes concurrently in 5.9 k cycles each -- the same time the 8-warp kernel needs for all four
// buffers (5.2 k) -- i.e. the radix-16 passes are bound by shared-memory bandwidth (64 KiB read + written per buffer and
// pass), not by latency, so twice the warps buy nothing there, while the join before the fused pass adds the lag of the
// half that waits for the staging-buffer handshake (fused pass 11 k cycles as seen from half 0).
// To try it again: include it after pbs_pipe.cuh in pbs_kernels.cuh and launch k_pbs_pipe16<P> with 512 threads.
//
//
// k_pbs_pipe runs 8 warps per SM at 252 registers: ncu shows its FP64 pipe 33 % busy and two warps per scheduler that
// cannot hide the load -> butterfly -> store chains (profiles/r02_pbs_pipe_final_*).  Here the CTA has 512 threads at
// <= 128 registers, organised as TWO HALVES of 256 threads that run the per-polynomial pipeline of k_pbs_pipe
// concurrently and independently (half-local named barriers):
//
//   half c (c = 0, 1)        : wait rotated polynomial c -> decompose (ONE stage-1 item per thread) -> emit level 1,
//                              emit level 0 (staged locally, bulk-copied to the owners) -> wait its two work buffers
//                              -> forward mid passes of its two buffers (2 x 128 radix-16 butterflies = 256 threads)
//   both halves, full barrier: fused pass -- half o accumulates OUTPUT o over the four digit polynomials (the last
//                              forward radix-8 pass is computed by both halves: +4 % FP64 work, half the registers)
//   half o                   : inverse mid passes of output o -> bulk copies to the owners' spare buffers -> wait its
//                              inverse data -> inverse stage 1, rounding, accumulator update (tensor memory), rotated
//                              pushes for the next CMUX.
//
// Shared memory, mbarriers, buffer map, the three split cluster handshakes and their safety argument are those of
// k_pbs_pipe.  The four staged emits still share three staging buffers: half 0 uses X then its rotated-accumulator half,
// half 1 its rotated-accumulator half then X -- after the "X has been read everywhere" handshake; while half 1 waits
// for it, half 0 keeps the SM busy with its own pipeline.  Every thread executes the cluster-barrier instructions (the
// hardware barrier counts all threads of the cluster); half 0 arrives early and waits late, so it never blocks on them.
//
// Tensor memory (512 columns per lane, four thread slots per lane): mid-pass twiddles 64 + 64 (the same butterfly index
// for the four slots: shared), accumulator 4 x 32, stage-1 twiddles 2 x 32 (slots s and s + 2 own the same item index).
//
// Replaces memref_batched_bootstrap_lwe_u64 of the absent concrete-compiler runtime
// (concrete/numpy/compilation/server.py:286, concrete/numpy/mlir/node_converter.py:1129-1208).
#pragma once

__device__ __forceinline__ void half_sync(int half) {        // named barrier of one 256-thread half
    asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
}

template <class P>
__global__ void __launch_bounds__(512, 1)
k_pbs_pipe16(u64* __restrict__ out, const u64* __restrict__ in, long B, const double2* __restrict__ bskF,
             const u64* __restrict__ luts, const int* __restrict__ lut_idx, const PbsGeom g,
             const double2* __restrict__ twT, const double2* __restrict__ tw_mid, long long* prof) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, R1 = P::R1, RF = P::RF, BUF = P::BUF, Q = P::Q, S = P::S;
    constexpr int N = P::N;
    constexpr int HT = 256;                   // threads per half
    constexpr int NG = BUF / RF;
    static_assert(C >= 2 && Q == HT && NG == HT && P::LRA > 0 && P::LRB > 0, "one stage-1 item and one fused-pass group per thread");
    static_assert(S * 8 == BUF * 16, "a rotated-accumulator half doubles as a staging buffer");
    constexpr unsigned BUF_BYTES = (unsigned)(sizeof(double2) << P::LOGBUF);
    constexpr unsigned TM_TWA = 0, TM_TWB = 64, TM_ACC0 = 128, TM_T0 = 256;
    const bool do_prof = (prof != nullptr) && blockIdx.x < 16 && threadIdx.x == 0;
    long long t_last = do_prof ? clock64() : 0;
#define PIPE_PROF(slot)                                  \
    if (do_prof) {                                       \
        const long long t_now = clock64();               \
        prof[16 * blockIdx.x + slot] += t_now - t_last;  \
        t_last = t_now;                                  \
    }
    u64* ra = reinterpret_cast<u64*>(smem_raw);                        // [2][S] rotated accumulator of the coming CMUX
    double2* work = reinterpret_cast<double2*>(ra + 2 * (size_t)S);    // [4][BUF]
    double2* xstg = work + 4 * (size_t)BUF;                            // [BUF] spare staging buffer
    __shared__ __align__(8) u64 s_bar[8];     // 0..3: forward data of work buffer b; 4, 5: inverse data of output o; 6, 7: rotated polynomial c
    __shared__ unsigned s_tmem_base;
    const int tid = threadIdx.x;
    const int half = tid >> 8;                // polynomial (P1, P2) / output (fused pass, P4, P5) of this thread
    const int lt = tid & (HT - 1);            // index inside the half: stage-1 item q and fused-pass group w
    const int rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const size_t ggsw_stride = (size_t)4 * 2 * C * BUF;               // double2 per CMUX iteration (J = 4, k + 1 = 2)
    const int base_log = g.base_log;
    const int nonrep = 64 - 2 * base_log;
    const double tscale = g.inv_M * 5.42101086242752217e-20;           // 2^-64 / M

    // ---- tensor memory
    if ((tid >> 5) == 0) tmem_alloc(&s_tmem_base, 512);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const unsigned tmb = s_tmem_base + ((unsigned)(((tid >> 5) & 3) * 32) << 16);
    const unsigned tA = tmb + TM_TWA, tB = tmb + TM_TWB;
    const unsigned tAcc = tmb + TM_ACC0 + (unsigned)((tid >> 7) * 4 * R1);
    const unsigned tT = tmb + TM_T0 + (unsigned)(((tid >> 7) & 1) * 4 * R1);
    const int j2 = (rank << P::LOGQ) + lt;
    if (tid < 128) {          // the butterfly index of the mid passes is tid % 128: one set of twiddles per lane
        tw_init_pass<P::LRA, P::LOGL1, P::LOGBUF, HT>(tA, tw_mid);
        tw_init_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, HT>(tB, tw_mid + P::TWA);
    }
    if (tid < HT) {           // stage-1 twiddles of item lt (both halves own the same item indices)
        double2 T[R1];
#pragma unroll
        for (int t1 = 0; t1 < R1; t1++) T[t1] = __ldg(&twT[((size_t)t1 << P::LOGL1) + j2]);
        tmem_store_c<R1>(tT, T);
    }
    tmem_wait_st();
    if (tid == 0) {
        for (int b = 0; b < 8; b++) mbar_init(&s_bar[b], 1);
        mbar_fence_init();
        for (int b = 0; b < 6; b++) mbar_arm(&s_bar[b], BUF_BYTES);    // first CMUX: forward and inverse data
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    cg::this_cluster().sync();                // every CTA's barriers exist before the first remote access
    cluster_arrive();                         // pairs with the wait of the first CMUX ("previous CMUX finished everywhere")
    unsigned ph = 0;
    u64* rah = ra + (size_t)half * S;                                  // rotated polynomial of this half
    double2* stg_lv1 = half ? reinterpret_cast<double2*>(rah) : xstg;  // staging of the level-1 emit
    double2* stg_lv0 = half ? xstg : reinterpret_cast<double2*>(rah);  // ... and of the level-0 emit
    const int off = SW(j2) & (Q - 1);

    for (long ct = cluster_id; ct < B; ct += nclusters) {
        const u64* lwe = in + (size_t)ct * (g.n + 1);
        const u64* lut = luts + (size_t)(lut_idx ? lut_idx[ct] : 0) * 2 * N;
        int a_next = modswitch(lwe[0], g.logN);
        u64 lwe_ahead = lwe[1];               // mask word i + 1, loaded one CMUX before its rotation amount is needed
        {
            // tensor memory: ACC = X^{-b~} * LUT; shared array: X^{a_0} * ACC, the operand of the first CMUX
            const int a0 = (2 * N - modswitch(lwe[g.n], g.logN)) & (2 * N - 1);
            const int a0s = (a0 + a_next) & (2 * N - 1);
            for (int w = tid; w < 2 * S; w += 2 * HT) {
                const int c = w >> P::LOGS, s = w & (S - 1);
                const int v = (slot_to_coef<P>(s, rank) - a0s) & (2 * N - 1);
                const u64 x = lut[(size_t)c * N + (v & (N - 1))];
                ra[w] = (v >> (P::LOGM + 1)) ? (0ULL - x) : x;
            }
            unsigned ow[4 * R1];
#pragma unroll
            for (int hj = 0; hj < 2 * R1; hj++) {
                const int v = ((hj << P::LOGL1) + j2 - a0) & (2 * N - 1);
                u64 x = lut[(size_t)half * N + (v & (N - 1))];
                if (v >> (P::LOGM + 1)) x = 0ULL - x;
                ow[2 * hj] = (unsigned)x;
                ow[2 * hj + 1] = (unsigned)(x >> 32);
            }
            TmemIo<4 * R1>::st(tAcc, ow);
            tmem_wait_st();
        }
        __syncthreads();
        if (tid == 0) {                       // the first rotated accumulator was written locally: complete its phase
            mbar_arrive(&s_bar[6]);
            mbar_arrive(&s_bar[7]);
        }

        for (int i = 0; i < g.n; i++) {
            const bool push = (i + 1 < g.n);
            a_next = push ? modswitch(lwe_ahead, g.logN) : 0;
            if (i + 2 < g.n) lwe_ahead = lwe[i + 2];
            // GGSW_{i+1} slices of this rank -> L2 one iteration ahead
            if (push && tid < 8)
                prefetch_l2_bulk(bskF + (size_t)(i + 1) * ggsw_stride + (((size_t)tid * C + rank) << P::LOGBUF), BUF_BYTES);
            // ---------------- P1 of polynomial `half`
            {
                double2 T[R1];
                tmem_load_c<R1>(tT, T);
                unsigned st[2 * R1];
                double2 v[R1];
                mbar_wait(&s_bar[6 + half], ph);
                PIPE_PROF(7)
                if (lt == 0 && push) mbar_arm(&s_bar[6 + half], BUF_BYTES);
                pipe_read_state<P>(rah, tAcc, lt, nonrep, st);
                half_sync(half);                                   // rotated polynomial consumed: its half may stage
                pipe_emit_compute<P>(st, base_log, T, g, v);       // level index 1 (least significant digits)
#pragma unroll
                for (int t1 = 0; t1 < R1; t1++) stg_lv1[(t1 << P::LOGQ) + off] = v[t1];
                fence_proxy_async();
                half_sync(half);
                // nothing of this CMUX may reach a peer that has not finished the previous one
                cluster_wait();
                if (lt < C) pipe_send<P>(stg_lv1, work + (size_t)(2 + half) * BUF, &s_bar[2 + half], rank, lt, g.dbg & 2);
                if (half == 0) {
                    cluster_arrive();                              // (X handshake: half 0 has nothing to say, it only keeps the count)
                    pipe_emit_compute<P>(st, base_log, T, g, v);   // level index 0
                } else {
                    // X is used a second time: every CTA must have received the chunks half 0 staged in it
                    pipe_emit_compute<P>(st, base_log, T, g, v);
                    mbar_wait(&s_bar[2], ph);
                    cluster_arrive();
                    cluster_wait();
                }
#pragma unroll
                for (int t1 = 0; t1 < R1; t1++) stg_lv0[(t1 << P::LOGQ) + off] = v[t1];
                fence_proxy_async();
                half_sync(half);
                if (lt < C) pipe_send<P>(stg_lv0, work + (size_t)half * BUF, &s_bar[half], rank, lt, g.dbg & 2);
            }
            PIPE_PROF(0)
            // ---------------- P2: forward mid passes of this half's two buffers (half, 2 + half)
            mbar_wait(&s_bar[half], ph);
            mbar_wait(&s_bar[2 + half], ph);
            PIPE_PROF(1)
            if (lt == 0) {                                          // next CMUX
                mbar_arm(&s_bar[half], BUF_BYTES);
                mbar_arm(&s_bar[2 + half], BUF_BYTES);
            }
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, HT, false, true>(work + (size_t)half * BUF, 2, tA, nullptr, 2, lt);
            half_sync(half);
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, HT, false, true>(work + (size_t)half * BUF, 2, tB, nullptr, 2, lt);
            if (half == 0) cluster_wait();                          // (X handshake: long complete)
            PIPE_PROF(2)
            // ---------------- P3: last forward pass + multiply-accumulate of OUTPUT `half` + first inverse pass
            {
                const double2* gp = bskF + (size_t)i * ggsw_stride + ((size_t)rank << P::LOGBUF) + lt + (size_t)half * ((size_t)C << P::LOGBUF);
                const size_t j_stride = (size_t)2 * ((size_t)C << P::LOGBUF);
                double2 r[RF], gv[RF];
#pragma unroll
                for (int m = 0; m < RF; m++) {
                    r[m] = make_double2(0.0, 0.0);
                    gv[m] = __ldg(gp + m * NG);
                }
                __syncthreads();                                    // all four digit buffers are transformed
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    double2 x[RF];
                    const double2* base = work + ((size_t)(((j & 1) << 1) | (j >> 1)) << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) x[m] = base[SW(lt * RF + m)];
                    Dft<RF, false>::run(x);
                    const double2* gn = gp + (size_t)(j + 1) * j_stride;
                    const bool more = (j + 1 < 4);
#pragma unroll
                    for (int m = 0; m < RF; m++) {
                        const double2 b = gv[m];
                        r[m].x = fma(x[m].x, b.x, r[m].x);
                        r[m].x = fma(-x[m].y, b.y, r[m].x);
                        r[m].y = fma(x[m].x, b.y, r[m].y);
                        r[m].y = fma(x[m].y, b.x, r[m].y);
                        if (more) gv[m] = __ldg(gn + m * NG);
                    }
                }
                Dft<RF, true>::run(r);
                __syncthreads();                                    // the other half still reads the buffers the outputs overwrite
                double2* base = work + ((size_t)half << P::LOGBUF);
#pragma unroll
                for (int m = 0; m < RF; m++) base[SW(lt * RF + m)] = r[m];
            }
            __syncthreads();
            cluster_arrive();                 // "my buffers 2 and 3 are free": completes behind the inverse passes
            PIPE_PROF(3)
            // ---------------- P4: inverse mid passes of output `half`, bulk copies to the owners' spare buffers
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, HT, true, true>(work + (size_t)half * BUF, 1, tB, nullptr, 1, lt);
            half_sync(half);
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, HT, true, true>(work + (size_t)half * BUF, 1, tA, nullptr, 1, lt);
            fence_proxy_async();
            half_sync(half);
            PIPE_PROF(4)
            cluster_wait();
            if (lt < C)
                pipe_send<P>(work + ((size_t)half << P::LOGBUF), work + ((size_t)(2 + half) << P::LOGBUF), &s_bar[4 + half], rank, lt, g.dbg & 4);
            PIPE_PROF(5)
            // ---------------- P5 of output `half`
            {
                double2 T[R1];
                tmem_load_c<R1>(tT, T);
                mbar_wait(&s_bar[4 + half], ph);
                if (lt == 0) mbar_arm(&s_bar[4 + half], BUF_BYTES);            // next CMUX
                unsigned ow[4 * R1];
                TmemIo<4 * R1>::ld(tAcc, ow);
                tmem_wait_ld();
                double2 z[R1];
                stage1_gather_local<P>(work + ((size_t)(2 + half) << P::LOGBUF), j2, T, g, z);
                // coefficient p = hj*L1 + j2 lands on (p + a') mod N, negated when it wraps once
                const int u = j2 + (a_next & (P::L1 - 1));
                const int j2d = u & (P::L1 - 1);
                const int hbd = (a_next >> P::LOGL1) + (u >> P::LOGL1);
                const int dcta = (g.dbg & 1) ? rank : (j2d >> P::LOGQ);
                const unsigned dst = mapa_u32(smem_u32(rah + (j2d & (Q - 1))), dcta);
                const unsigned dbar = mapa_u32(smem_u32(&s_bar[6 + half]), dcta);
#pragma unroll
                for (int hj = 0; hj < 2 * R1; hj++) {
                    const int j1 = hj & (R1 - 1);
                    const u64 v = (((u64)ow[2 * hj + 1] << 32) | (u64)ow[2 * hj]) +
                                  f2torus_frac((hj >= R1) ? z[j1].y : z[j1].x, tscale);
                    ow[2 * hj] = (unsigned)v;
                    ow[2 * hj + 1] = (unsigned)(v >> 32);
                    if (push) {
                        const int bt = (hj + hbd) & (4 * R1 - 1);
                        st_async_u64(dst + ((unsigned)(bt & (2 * R1 - 1)) << (P::LOGQ + 3)),
                                     (bt >> (P::LOGR1 + 1)) ? (0ULL - v) : v, dbar);
                    }
                }
                TmemIo<4 * R1>::st(tAcc, ow);
                tmem_wait_st();
            }
            cluster_arrive();                 // "I have finished this CMUX": awaited before the first send of the next one
            ph ^= 1u;
            PIPE_PROF(6)
        }
        // ---- sample extract coefficient 0: the accumulator words are in the owners' tensor-memory lanes
        {
            u64* o = out + (size_t)ct * ((size_t)N + 1);
            unsigned ow[4 * R1];
            TmemIo<4 * R1>::ld(tAcc, ow);
            tmem_wait_ld();
#pragma unroll
            for (int hj = 0; hj < 2 * R1; hj++) {
                const int j = (hj << P::LOGL1) + j2;
                const u64 v = ((u64)ow[2 * hj + 1] << 32) | (u64)ow[2 * hj];
                if (half == 0) {
                    if (j == 0) o[0] = v;
                    else o[N - j] = 0ULL - v;
                } else if (j == 0)
                    o[N] = v;
            }
        }
        PIPE_PROF(8)
    }
    cluster_wait();                           // pairs with the arrive of the last CMUX
    cg::this_cluster().sync();                // nobody leaves while a peer could still address its shared memory
    tmem_fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(s_tmem_base, 512);
#undef PIPE_PROF
}
