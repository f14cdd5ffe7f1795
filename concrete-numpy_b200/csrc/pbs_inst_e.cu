// Explicit instantiations of the bootstrap kernels: pipelined plans (k_pbs_pipe, pbs_pipe.cuh) for the 4- and 2-CTA
// clusters; picked for k = 1, two-level bootstraps (BASELINE configs 3 and 4 run N = 8192, config 5 N = 4096).
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    //             C  logR1 RA RB RF  NT  minb kmax acctm
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 4, 3, 256, 1, 1, 2>),   // L1 = 2048 (N = 16384)
    PBS_PLAN_ENTRY(PbsPlan<2, 1, 4, 4, 3, 256, 1, 1, 2>),   // L1 = 2048 (N = 8192)
    // latency plan: one ciphertext on 4 SMs with 512-point local transforms (small batches of N = 4096 lookups: the
    // sequential chains of the key-value database circuits); thread pairs share a fused-pass group
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 3, 3, 3, 128, 1, 1, 2>),   // L1 = 512 (N = 4096)
};
const PbsPlanEntry* pbs_plans_e(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
