// Explicit instantiations of the bootstrap kernels: 16- and 8-CTA cluster plans (N = 32768 / 16384).
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    //             C  logR1 RA RB RF  NT  minb kmax
    PBS_PLAN_ENTRY(PbsPlan<16, 4, 4, 3, 3, 128, 2, 1>),  // L1 = 1024, two clusters share every SM (N = 32768)
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 4, 4, 3, 256, 1, 1>),   // L1 = 2048 (N = 32768)
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 4, 4, 3, 256, 1, 1, 2>),   // same, accumulator in tensor memory + pushed rotated copy; two-level bootstraps run k_pbs_pipe
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 4, 3, 3, 128, 2, 1>),   // L1 = 1024 (N = 16384)
};
const PbsPlanEntry* pbs_plans_a(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
