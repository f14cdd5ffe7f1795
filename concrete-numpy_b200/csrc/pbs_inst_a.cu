// Explicit instantiations of the bootstrap kernels: 8- and 4-CTA cluster plans (N = 32768 / 16384 / 8192).
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    //             C  logR1 RA RB RF  NT  minb kmax
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 4, 4, 3, 256, 1, 1>),   // L1 = 2048
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 4, 3, 3, 256, 1, 1>),   // L1 = 1024
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 4, 3, 256, 1, 1>),
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 3, 3, 256, 1, 1>),
};
const PbsPlanEntry* pbs_plans_a(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
