// Pipelined bootstrap for the cluster plans with 2048-point local transforms (N = 32768 on 8 CTAs, 16384 on 4, 8192 on
// 2), k = 1, two decomposition levels.
//
// Same arithmetic, same Fourier key layout and same FFT factorisation as k_pbs with the accumulator in tensor memory
// (PbsPlan::ACC_TM); what changes is how the three exchanges of a CMUX cross the cluster:
//
//   * forward (stage-1 outputs, 112 KiB per CTA and CMUX): every thread writes its R1 outputs into a LOCAL staging
//     buffer ordered by destination CTA; 4 KiB chunks then travel as bulk asynchronous copies
//     (cp.async.bulk.shared::cluster, ~2-3x the rate of element-wise st.shared::cluster under load,
//     profiles/r01_ubench_dsmem_bulk.log) and complete on one mbarrier per work buffer in the destination CTA.  The
//     copy engine moves polynomial 0 while the threads decompose polynomial 1, and polynomial 1 while they run the
//     mid passes of polynomial 0.  Staging space: one spare 32 KiB buffer plus the two halves of the rotated-accumulator
//     array, each dead as soon as its polynomial has been read.
//   * inverse: bulk copies into the peers' spare work buffers, one mbarrier per output polynomial (as in k_pbs).
//   * rotated accumulator of the next CMUX: st.async (remote store + complete_tx) counted on one mbarrier per
//     polynomial in the destination, so the next CMUX starts on polynomial 0 while polynomial 1 still drains.
//
// No cluster-wide release/acquire barrier (MEMBAR.ALL.GPU + drain) is left inside a CMUX: a CTA waits only for the
// bytes it is about to read.  The hardware cluster barrier is used three times per CMUX in its relaxed split form
// (arrive early, wait late): "every CTA has finished the previous CMUX" (its work buffers may be overwritten: a CTA's
// rotated accumulator comes from two peers only, so its arrival says nothing about the other five), "staging buffer X
// has been read by every destination" and "my spare buffers are free".
//
// Buffer map (level = 2): digit polynomial (c, lv) lives in work buffer 2*lv + c.  The buffers polynomial 0 is sent to
// (0 and 2) are exactly the ones whose previous contents -- output 0 and the inverse exchange of output 0 -- are known
// to be consumed once the rotated polynomial 0 of the next CMUX has arrived from every CTA; likewise 1 and 3.
//
// Replaces memref_batched_bootstrap_lwe_u64 of the absent concrete-compiler runtime
// (concrete/numpy/compilation/server.py:286, concrete/numpy/mlir/node_converter.py:1129-1208).
#pragma once

__device__ __forceinline__ void st_async_u64(unsigned dst_cluster_addr, u64 v, unsigned mbar_cluster_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(dst_cluster_addr), "l"(v), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A thread owns IPT = Q / NT items of every polynomial (item ii: q = tid + ii * NT); an item is the 2*R1 coefficients of one
// stage-1 butterfly.  IPT * R1 = L1 / NT is the same for every plan with the same L1 and thread count (8 for L1 = 2048,
// 256 threads): R1 = 8 -> one radix-8 item, R1 = 2 -> four radix-2 items, and the per-thread register / tensor-memory
// footprint does not depend on the cluster size.
//
// 32-bit decomposition states of this thread's items of one polynomial: closest representable of (X^a * ACC - ACC); the
// rotated operand was pushed into `rac` by its producers, the thread's own words wait in its tensor-memory lane
template <class P>
__device__ __forceinline__ void pipe_read_state(const u64* __restrict__ rac, unsigned tacc, int tid, int nonrep,
                                                unsigned* st /* [IPT][2*R1] */) {
    constexpr int R1 = P::R1, IPT = P::Q / P::NT;
    unsigned ow[IPT * 4 * R1];
    TmemIo<IPT * 4 * R1>::ld(tacc, ow);
    u64 rv[IPT * 2 * R1];
#pragma unroll
    for (int ii = 0; ii < IPT; ii++)
#pragma unroll
        for (int hj = 0; hj < 2 * R1; hj++) rv[ii * 2 * R1 + hj] = rac[(hj << P::LOGQ) + tid + ii * P::NT];
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < IPT * 2 * R1; e++) {
        const u64 diff = rv[e] - (((u64)ow[2 * e + 1] << 32) | (u64)ow[2 * e]);
        const unsigned top = (unsigned)(diff >> 32);
        st[e] = ((top >> (unsigned)(nonrep - 33)) + 1u) >> 1;
    }
}

// next digit of every state (least significant level first), fold + twist, radix-R1 butterfly, stage twiddle
template <class P>
__device__ __forceinline__ void pipe_emit_compute(unsigned* st, int base_log, const double2* T /* [IPT][R1] */,
                                                  const PbsGeom& g, double2* v /* [IPT][R1] */) {
    constexpr int R1 = P::R1, IPT = P::Q / P::NT;
    const unsigned mask = (1u << base_log) - 1u;
#pragma unroll
    for (int ii = 0; ii < IPT; ii++) {
        double2 z[R1];
#pragma unroll
        for (int j1 = 0; j1 < R1; j1++) {
            double d[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int e = ii * 2 * R1 + h * R1 + j1;
                const unsigned res = st[e] & mask;
                const unsigned s2 = st[e] >> base_log;
                const unsigned carry = (((res - 1u) | s2) & res) >> (base_log - 1);
                st[e] = s2 + carry;
                d[h] = i32_to_f64((int)res - (int)(carry << base_log));
            }
            const double2 dd = make_double2(d[0], d[1]);
            z[j1] = (j1 == 0) ? dd : cmul(dd, g.cj[j1]);
        }
        Dft<R1, false>::run(z);
#pragma unroll
        for (int t1 = 0; t1 < R1; t1++) v[ii * R1 + t1] = cmul(z[t1], T[ii * R1 + t1]);
    }
}

// chunk t1 of the staging buffer goes to CTA t1; inside the chunk the element sits where the receiver's swizzle wants it
template <class P>
__device__ __forceinline__ void pipe_stage(double2* __restrict__ stg, int j2_0, const double2* v) {
    constexpr int R1 = P::R1, IPT = P::Q / P::NT;
#pragma unroll
    for (int ii = 0; ii < IPT; ii++) {
        const int off = SW(j2_0 + ii * P::NT) & (P::Q - 1);
#pragma unroll
        for (int t1 = 0; t1 < R1; t1++) stg[(t1 << P::LOGQ) + off] = v[ii * R1 + t1];
    }
}

// thread d < C: chunk d of `stg` -> chunk `rank` of work buffer `dst_buf` of CTA d, counted on its mbarrier `bar`
template <class P>
__device__ __forceinline__ void pipe_send(const double2* stg, double2* dst_buf, u64* bar, int rank, int d, bool self = false) {
    if (self) {         // development ablation (FHE_PBS_DBG): same bytes, same barrier, no SM-to-SM traffic; results are wrong
        bulk_copy_to_cluster(mapa_u32(smem_u32(dst_buf + (d << P::LOGQ)), rank), stg + (d << P::LOGQ),
                             (unsigned)(sizeof(double2) << P::LOGQ), mapa_u32(smem_u32(bar), rank));
        return;
    }
    bulk_copy_to_cluster(mapa_u32(smem_u32(dst_buf + (rank << P::LOGQ)), d), stg + (d << P::LOGQ),
                         (unsigned)(sizeof(double2) << P::LOGQ), mapa_u32(smem_u32(bar), d));
}

// GGSW staging of the pipelined kernel: on when two stages of the CTA's 8 slices fit next to its buffers
template <class P>
struct PipeStage {
    static constexpr size_t BASE = (size_t)2 * P::S * 8 + (size_t)5 * P::BUF * 16;
    static constexpr size_t RING = (size_t)2 * 8 * P::BUF * 16;
    static constexpr bool ON = BASE + RING + 4096 <= PBS_SMEM_MAX;
    static constexpr size_t SMEM = BASE + (ON ? RING : 0);
};
template <bool SMEM_SRC>
__device__ __forceinline__ double2 pipe_ld(const double2* p) {
    if constexpr (SMEM_SRC) return *p;
    else return __ldg(p);
}

template <class P>
__global__ void __launch_bounds__(P::NT, 1)
k_pbs_pipe(u64* __restrict__ out, const u64* __restrict__ in, long B, const double2* __restrict__ bskF,
           const u64* __restrict__ luts, const int* __restrict__ lut_idx, const PbsGeom g,
           const double2* __restrict__ twT, const double2* __restrict__ tw_mid, long long* prof,
           unsigned* __restrict__ wave_ctr /* zeroed per launch, or null: no wave alignment */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = P::C, R1 = P::R1, NT = P::NT, RF = P::RF, BUF = P::BUF, Q = P::Q, S = P::S;
    constexpr int N = P::N;
    constexpr int NG = BUF / RF;
    constexpr int IPT = Q / NT;               // items per thread and polynomial
    static_assert(P::ACC_TM && P::TMEM_TW && C >= 2 && Q % NT == 0 && (NG == NT || 2 * NG == NT) && IPT * R1 * 4 <= 64,
                  "pipelined bootstrap: whole items per thread; one fused-pass group per thread, or per thread pair (one output each)");
    static_assert(S * 8 == BUF * 16, "a rotated-accumulator half doubles as a staging buffer");
    constexpr unsigned BUF_BYTES = (unsigned)(sizeof(double2) << P::LOGBUF);
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
    // Plans with small local transforms (the latency plans: one ciphertext on several SMs, a handful of warps per SM) leave
    // most of the shared memory free and cannot hide the L2 latency of the key behind other warps: the fused pass spent
    // 3.3 k of 12 k cycles per CMUX waiting for its GGSW values (N = 4096 on 4 CTAs).  There the CTA's 8 slices of GGSW_{i+1}
    // are fetched by bulk copies into a two-stage ring while CMUX i runs, and the fused pass reads shared memory.
    constexpr bool STG = PipeStage<P>::ON;
    double2* gst = xstg + (size_t)BUF;                                 // [2][8][BUF] staged GGSW slices (STG plans)
    __shared__ __align__(8) u64 s_gbar[2];    // staged GGSW of ring stage s has landed
    __shared__ __align__(8) u64 s_bar[8];     // 0..3: forward data of work buffer b; 4, 5: inverse data of output o; 6, 7: rotated polynomial c
    const int tid = threadIdx.x;
    const int rank = (int)cg::this_cluster().block_rank();
    const long cluster_id = blockIdx.x / C, nclusters = gridDim.x / C;
    const size_t ggsw_stride = (size_t)4 * 2 * C * BUF;               // double2 per CMUX iteration (J = 4, k + 1 = 2)
    const int base_log = g.base_log;
    const int nonrep = 64 - 2 * base_log;
    const double tscale = g.inv_M * 5.42101086242752217e-20;           // 2^-64 / M

    const TwCtx tw = tw_setup<P>(nullptr, twT, tw_mid);
    if (tid == 0) {
        for (int b = 0; b < 8; b++) mbar_init(&s_bar[b], 1);
        for (int b = 0; b < 2; b++) mbar_init(&s_gbar[b], 1);
        mbar_fence_init();
        for (int b = 0; b < 6; b++) mbar_arm(&s_bar[b], BUF_BYTES);    // first CMUX: forward and inverse data
    }
    __syncthreads();
    cg::this_cluster().sync();                // every CTA's barriers exist before the first remote access
    cluster_arrive();                         // pairs with the wait of the first CMUX ("previous CMUX finished everywhere")
    unsigned ph = 0;
    unsigned gph = 0;                                  // bit s: parity the next wait on s_gbar[s] uses
    // thread t < 8 issues slice t = 2 * j + o of GGSW_i for this rank into ring stage `stage`
    auto stage_ggsw = [&](int i, int stage) {
        if (tid == 0) mbar_arm(&s_gbar[stage], 8 * BUF_BYTES);
        __syncwarp();
        if (tid < 8)
            bulk_copy_from_global(gst + ((size_t)(stage * 8 + tid) << P::LOGBUF),
                                  bskF + (size_t)i * ggsw_stride + (((size_t)tid * C + rank) << P::LOGBUF), BUF_BYTES, &s_gbar[stage]);
    };
    const int j2_0 = (rank << P::LOGQ) + tid;         // item ii of this thread: q = tid + ii * NT, j2 = j2_0 + ii * NT
    // the stage-1 twiddles of this thread's items never change: they live in the tensor-memory columns the plan leaves free
    constexpr int TM_SLOTS = (NT + 127) / 128;       // thread slots per tensor-memory lane
    static_assert(TM_SLOTS * (P::TM_THREAD + 4 * IPT * R1) <= P::TM_CTA, "room for the stage-1 twiddles");
    const unsigned tT = tw.base + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)(TM_SLOTS * P::TM_THREAD + (tid >> 7) * 4 * IPT * R1);
    {
        double2 T[IPT * R1];
#pragma unroll
        for (int ii = 0; ii < IPT; ii++) load_T<P>(T + ii * R1, tw, j2_0 + ii * NT);
        tmem_store_c<IPT * R1>(tT, T);
        tmem_wait_st();
    }
    for (long ct = cluster_id; ct < B; ct += nclusters) {
        if (wave_ctr != nullptr && rank == 0 && tid == 0) {
            // Wave alignment: the key (2 GiB at N = 32768) streams from HBM once per WAVE of resident clusters only while
            // they stay within an L2's worth of CMUX iterations of each other (2.1 MB per iteration, 126 MB of L2); left
            // alone they drift apart over a launch and re-fetch slices a faster cluster already evicted (measured 1.9x the
            // key bytes).  Every cluster counts itself in before a new ciphertext and waits for the rest of its wave; the
            // wait is bounded (clusters that are not co-resident -- another kernel on the device -- must not deadlock).
            const long wave = (ct - cluster_id) / nclusters;
            const long want = (wave + 1) * nclusters;
            const unsigned target = (unsigned)(want < B ? want : B);
            atomicAdd(wave_ctr, 1u);
            const long long t_end = clock64() + 4000000LL;
            while (*reinterpret_cast<volatile unsigned*>(wave_ctr) < target && clock64() < t_end) __nanosleep(256);
        }
        const u64* lwe = in + (size_t)ct * (g.n + 1);
        const u64* lut = luts + (size_t)(lut_idx ? lut_idx[ct] : 0) * 2 * N;
        int a_next = modswitch(lwe[0], g.logN);
        u64 lwe_ahead = lwe[1];               // mask word i + 1, loaded one CMUX before its rotation amount is needed
        {
            // tensor memory: ACC = X^{-b~} * LUT; shared array: X^{a_0} * ACC, the operand of the first CMUX
            const int a0 = (2 * N - modswitch(lwe[g.n], g.logN)) & (2 * N - 1);
            const int a0s = (a0 + a_next) & (2 * N - 1);
            for (int w = tid; w < 2 * S; w += NT) {
                const int c = w >> P::LOGS, s = w & (S - 1);
                const int v = (slot_to_coef<P>(s, rank) - a0s) & (2 * N - 1);
                const u64 x = lut[(size_t)c * N + (v & (N - 1))];
                ra[w] = (v >> (P::LOGM + 1)) ? (0ULL - x) : x;
            }
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                unsigned ow[IPT * 4 * R1];
#pragma unroll
                for (int ii = 0; ii < IPT; ii++)
#pragma unroll
                    for (int hj = 0; hj < 2 * R1; hj++) {
                        const int v = ((hj << P::LOGL1) + j2_0 + ii * NT - a0) & (2 * N - 1);
                        u64 x = lut[(size_t)c * N + (v & (N - 1))];
                        if (v >> (P::LOGM + 1)) x = 0ULL - x;
                        ow[(ii * 2 * R1 + hj) * 2] = (unsigned)x;
                        ow[(ii * 2 * R1 + hj) * 2 + 1] = (unsigned)(x >> 32);
                    }
                TmemIo<IPT * 4 * R1>::st(tw.tAcc + (unsigned)(c * IPT * 4 * R1), ow);
            }
            tmem_wait_st();
        }
        __syncthreads();
        if (tid == 0) {                       // the first rotated accumulator was written locally: complete its phase
            mbar_arrive(&s_bar[6]);
            mbar_arrive(&s_bar[7]);
        }
        if constexpr (STG) {
            if (tid < 32) stage_ggsw(0, 0);   // both stages are free: every fused pass of the previous ciphertext is over
        }

        for (int i = 0; i < g.n; i++) {
            const bool push = (i + 1 < g.n);
            a_next = push ? modswitch(lwe_ahead, g.logN) : 0;
            if (i + 2 < g.n) lwe_ahead = lwe[i + 2];
            // GGSW_{i+1} slices of this rank: into the other ring stage (its readers finished with the fused pass of CMUX
            // i - 1, a __syncthreads ago), or into L2 one iteration ahead
            if constexpr (!STG) {
                if (push && tid < 8)
                    prefetch_l2_bulk(bskF + (size_t)(i + 1) * ggsw_stride + (((size_t)tid * C + rank) << P::LOGBUF), BUF_BYTES);
            }
            // ---------------- P1: decompose X^a * ACC - ACC, stage 1, bulk copies to the owners
            {
                double2 T[IPT * R1];
                tmem_load_c<IPT * R1>(tT, T);
                unsigned st[IPT * 2 * R1];
                double2 v[IPT * R1];
                // polynomial 0
                mbar_wait(&s_bar[6], ph);
                PIPE_PROF(7)
                if (tid == 0 && push) mbar_arm(&s_bar[6], BUF_BYTES);
                pipe_read_state<P>(ra, tw.tAcc, tid, nonrep, st);
                __syncthreads();                                   // rotated polynomial 0 consumed: its half stages (0, lv 0)
                PIPE_PROF(9)
                pipe_emit_compute<P>(st, base_log, T, g, v);       // level index 1 (least significant digits)
                pipe_stage<P>(xstg, j2_0, v);
                fence_proxy_async();                 // every writer: its staged values must be visible to the copy engine
                __syncthreads();
                // nothing of this CMUX may reach a peer that has not finished the previous one (it still reads its spare
                // buffers and its inverse copies still read its output buffers); also covers the ciphertext boundary
                cluster_wait();
                if (tid < C) pipe_send<P>(xstg, work + 2 * (size_t)BUF, &s_bar[2], rank, tid, g.dbg & 2);
                PIPE_PROF(10)
                pipe_emit_compute<P>(st, base_log, T, g, v);       // level index 0
                pipe_stage<P>(reinterpret_cast<double2*>(ra), j2_0, v);
                fence_proxy_async();                 // every writer: its staged values must be visible to the copy engine
                __syncthreads();
                if (tid < C) pipe_send<P>(reinterpret_cast<double2*>(ra), work, &s_bar[0], rank, tid, g.dbg & 2);
                PIPE_PROF(11)
                // polynomial 1
                mbar_wait(&s_bar[7], ph);
                PIPE_PROF(12)
                if (tid == 0 && push) mbar_arm(&s_bar[7], BUF_BYTES);
                pipe_read_state<P>(ra + S, tw.tAcc + (unsigned)(IPT * 4 * R1), tid, nonrep, st);
                __syncthreads();
                pipe_emit_compute<P>(st, base_log, T, g, v);
                pipe_stage<P>(reinterpret_cast<double2*>(ra + S), j2_0, v);
                fence_proxy_async();                 // every writer: its staged values must be visible to the copy engine
                __syncthreads();
                if (tid < C) pipe_send<P>(reinterpret_cast<double2*>(ra + S), work + 3 * (size_t)BUF, &s_bar[3], rank, tid, g.dbg & 2);
                PIPE_PROF(13)
                // the spare staging buffer is used a second time: every CTA must have received the chunks staged in it
                mbar_wait(&s_bar[2], ph);
                cluster_arrive();
                PIPE_PROF(14)
                pipe_emit_compute<P>(st, base_log, T, g, v);
                cluster_wait();
                PIPE_PROF(15)
                pipe_stage<P>(xstg, j2_0, v);
                fence_proxy_async();                 // every writer: its staged values must be visible to the copy engine
                __syncthreads();
                if (tid < C) pipe_send<P>(xstg, work + (size_t)BUF, &s_bar[1], rank, tid, g.dbg & 2);
                // staged GGSW_{i+1}: issued BEHIND the last forward send -- the copy engine serves its queue in order, and
                // 64 KiB from L2 / HBM ahead of the sends delayed every one of them (measured: P1 +1.4 k cycles per CMUX);
                // the next latency-critical copies (the inverse exchange) are three phases away
                if constexpr (STG) {
                    if (push && tid < 32) stage_ggsw(i + 1, (i + 1) & 1);
                }
            }
            PIPE_PROF(0)
            // ---------------- P2: forward mid passes, polynomial 0 while polynomial 1 is still in flight
            mbar_wait(&s_bar[0], ph);
            mbar_wait(&s_bar[2], ph);
            PIPE_PROF(1)
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, NT, false, true>(work, 2, tw.tA, tw.sA, 2);
            __syncthreads();
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, false, true>(work, 2, tw.tB, tw.sB, 2);
            __syncthreads();
            mbar_wait(&s_bar[1], ph);
            mbar_wait(&s_bar[3], ph);
            if (tid == 0)
                for (int b = 0; b < 4; b++) mbar_arm(&s_bar[b], BUF_BYTES);      // next CMUX
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, NT, false, true>(work + BUF, 2, tw.tA, tw.sA, 2);
            __syncthreads();
            const double2* gp = STG ? gst + ((size_t)((i & 1) * 8) << P::LOGBUF)
                                    : bskF + (size_t)i * ggsw_stride + ((size_t)rank << P::LOGBUF);
            const size_t o_stride = STG ? (size_t)BUF : (size_t)C << P::LOGBUF;
            if constexpr (STG) {
                mbar_wait(&s_gbar[i & 1], (gph >> (i & 1)) & 1u);
                gph ^= 1u << (i & 1);
            }
            if constexpr (NG == NT) {
                // the GGSW values of the first digit polynomial are requested before the last mid pass: their L2 latency
                // hides behind it instead of opening the multiply-accumulate
                gp += tid;
                double2 gv[2][RF];
#pragma unroll
                for (int o = 0; o < 2; o++)
#pragma unroll
                    for (int m = 0; m < RF; m++) gv[o][m] = pipe_ld<STG>(gp + o * o_stride + m * NG);
                if constexpr (P::LRB > 0) {
                    mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, false, true>(work + BUF, 2, tw.tB, tw.sB, 2);
                    __syncthreads();
                }
                PIPE_PROF(2)
                // ---------------- P3: last forward pass + multiply-accumulate against GGSW_i + first inverse pass
                double2 r[2][RF];
#pragma unroll
                for (int o = 0; o < 2; o++)
#pragma unroll
                    for (int m = 0; m < RF; m++) r[o][m] = make_double2(0.0, 0.0);
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    double2 x[RF];
                    const double2* base = work + ((size_t)(((j & 1) << 1) | (j >> 1)) << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) x[m] = base[SW(tid * RF + m)];
                    Dft<RF, false>::run(x);
                    const double2* gn = gp + (size_t)(j + 1) * 2 * o_stride;
                    const bool more = (j + 1 < 4);
#pragma unroll
                    for (int o = 0; o < 2; o++)
#pragma unroll
                        for (int m = 0; m < RF; m++) {
                            const double2 b = gv[o][m];
                            r[o][m].x = fma(x[m].x, b.x, r[o][m].x);
                            r[o][m].x = fma(-x[m].y, b.y, r[o][m].x);
                            r[o][m].y = fma(x[m].x, b.y, r[o][m].y);
                            r[o][m].y = fma(x[m].y, b.x, r[o][m].y);
                            if (more) gv[o][m] = pipe_ld<STG>(gn + o * o_stride + m * NG);
                        }
                }
#pragma unroll
                for (int o = 0; o < 2; o++) {
                    Dft<RF, true>::run(r[o]);
                    double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) base[SW(tid * RF + m)] = r[o][m];
                }
            } else {
                // half as many fused-pass groups as threads (small local transforms): thread pair (w, w + NG) shares group
                // w, each thread accumulates ONE output polynomial (o = tid / NG) -- the last forward radix pass is computed
                // by both, the GGSW loads and multiply-adds are split
                const int w = tid % NG, o = tid / NG;
                gp += w + (size_t)o * o_stride;
                double2 gv[RF];
#pragma unroll
                for (int m = 0; m < RF; m++) gv[m] = pipe_ld<STG>(gp + m * NG);
                if constexpr (P::LRB > 0) {
                    mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, false, true>(work + BUF, 2, tw.tB, tw.sB, 2);
                    __syncthreads();
                }
                PIPE_PROF(2)
                double2 r[RF];
#pragma unroll
                for (int m = 0; m < RF; m++) r[m] = make_double2(0.0, 0.0);
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    double2 x[RF];
                    const double2* base = work + ((size_t)(((j & 1) << 1) | (j >> 1)) << P::LOGBUF);
#pragma unroll
                    for (int m = 0; m < RF; m++) x[m] = base[SW(w * RF + m)];
                    Dft<RF, false>::run(x);
                    const double2* gn = gp + (size_t)(j + 1) * 2 * o_stride;
                    const bool more = (j + 1 < 4);
#pragma unroll
                    for (int m = 0; m < RF; m++) {
                        const double2 b = gv[m];
                        r[m].x = fma(x[m].x, b.x, r[m].x);
                        r[m].x = fma(-x[m].y, b.y, r[m].x);
                        r[m].y = fma(x[m].x, b.y, r[m].y);
                        r[m].y = fma(x[m].y, b.x, r[m].y);
                        if (more) gv[m] = pipe_ld<STG>(gn + m * NG);
                    }
                }
                Dft<RF, true>::run(r);
                __syncthreads();              // the partner thread still reads the buffers the outputs overwrite
                double2* base = work + ((size_t)o << P::LOGBUF);
#pragma unroll
                for (int m = 0; m < RF; m++) base[SW(w * RF + m)] = r[m];
            }
            __syncthreads();
            cluster_arrive();                 // "my buffers 2 and 3 are free": completes behind the inverse passes
            PIPE_PROF(3)
            // ---------------- P4: inverse mid passes of the two outputs, bulk copies to the owners' spare buffers
            mid_pass<P::LRB, P::LOGL1 - P::LRA, P::LOGBUF, NT, true, true>(work, 2, tw.tB, tw.sB);
            __syncthreads();
            mid_pass<P::LRA, P::LOGL1, P::LOGBUF, NT, true, true>(work, 2, tw.tA, tw.sA);
            fence_proxy_async();
            __syncthreads();
            PIPE_PROF(4)
            cluster_wait();
            if (tid < 2 * C) {
                const int o = tid / C, d = tid % C;
                pipe_send<P>(work + ((size_t)o << P::LOGBUF), work + ((size_t)(2 + o) << P::LOGBUF), &s_bar[4 + o], rank, d, g.dbg & 4);
            }
            PIPE_PROF(5)
            // ---------------- P5: inverse radix-R1 stage, untwist, round, ACC += ; push X^{a'} * ACC to its readers
            {
                double2 T[IPT * R1];
                tmem_load_c<IPT * R1>(tT, T);
#pragma unroll 1
                for (int o = 0; o < 2; o++) {
                    mbar_wait(&s_bar[4 + o], ph);
                    if (tid == 0) mbar_arm(&s_bar[4 + o], BUF_BYTES);          // next CMUX
                    const unsigned ta = tw.tAcc + (unsigned)(o * IPT * 4 * R1);
                    unsigned ow[IPT * 4 * R1];
                    TmemIo<IPT * 4 * R1>::ld(ta, ow);
                    tmem_wait_ld();
#pragma unroll
                    for (int ii = 0; ii < IPT; ii++) {
                        const int j2 = j2_0 + ii * NT;
                        double2 z[R1];
                        stage1_gather_local<P>(work + ((size_t)(2 + o) << P::LOGBUF), j2, T + ii * R1, g, z);
                        // coefficient p = hj*L1 + j2 lands on (p + a') mod N, negated when it wraps once
                        const int u = j2 + (a_next & (P::L1 - 1));
                        const int j2d = u & (P::L1 - 1);
                        const int hbd = (a_next >> P::LOGL1) + (u >> P::LOGL1);
                        const int dcta = (g.dbg & 1) ? rank : (j2d >> P::LOGQ);
                        const unsigned dst = mapa_u32(smem_u32(ra + (size_t)o * S + (j2d & (Q - 1))), dcta);
                        const unsigned dbar = mapa_u32(smem_u32(&s_bar[6 + o]), dcta);
#pragma unroll
                        for (int hj = 0; hj < 2 * R1; hj++) {
                            const int e = ii * 2 * R1 + hj;
                            const int j1 = hj & (R1 - 1);
                            const u64 v = (((u64)ow[2 * e + 1] << 32) | (u64)ow[2 * e]) +
                                          f2torus_frac((hj >= R1) ? z[j1].y : z[j1].x, tscale);
                            ow[2 * e] = (unsigned)v;
                            ow[2 * e + 1] = (unsigned)(v >> 32);
                            if (push) {
                                const int bt = (hj + hbd) & (4 * R1 - 1);
                                st_async_u64(dst + ((unsigned)(bt & (2 * R1 - 1)) << (P::LOGQ + 3)),
                                             (bt >> (P::LOGR1 + 1)) ? (0ULL - v) : v, dbar);
                            }
                        }
                    }
                    TmemIo<IPT * 4 * R1>::st(ta, ow);
                }
                tmem_wait_st();
            }
            cluster_arrive();                 // "I have finished this CMUX": awaited before the first send of the next one
            ph ^= 1u;
            PIPE_PROF(6)
        }
        // ---- sample extract coefficient 0: the accumulator words are in the owners' tensor-memory lanes
        {
            u64* o = out + (size_t)ct * ((size_t)N + 1);
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                unsigned ow[IPT * 4 * R1];
                TmemIo<IPT * 4 * R1>::ld(tw.tAcc + (unsigned)(c * IPT * 4 * R1), ow);
                tmem_wait_ld();
#pragma unroll
                for (int ii = 0; ii < IPT; ii++)
#pragma unroll
                    for (int hj = 0; hj < 2 * R1; hj++) {
                        const int j = (hj << P::LOGL1) + j2_0 + ii * NT;
                        const int e = ii * 2 * R1 + hj;
                        const u64 v = ((u64)ow[2 * e + 1] << 32) | (u64)ow[2 * e];
                        if (c == 0) {
                            if (j == 0) o[0] = v;
                            else o[N - j] = 0ULL - v;
                        } else if (j == 0)
                            o[N] = v;
                    }
            }
        }
        PIPE_PROF(8)
    }
    cluster_wait();                           // pairs with the arrive of the last CMUX
    cg::this_cluster().sync();                // nobody leaves while a peer could still address its shared memory
    tw_teardown<P>(tw);
#undef PIPE_PROF
}
