// Explicit instantiations of the bootstrap kernels: small single-CTA plans (N <= 1024, and k >= 2).
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 3, 0, 3, 128, 2, 1>),   // M = 512
    PBS_PLAN_ENTRY(PbsPlan<1, 2, 3, 0, 3, 128, 2, 1>),   // M = 256
    PBS_PLAN_ENTRY(PbsPlan<1, 1, 3, 0, 3, 64, 1, 1>),    // M = 128
    // k = 2, 3: radix-4 fused pass keeps the (k+1) x RF accumulators in registers
    PBS_PLAN_ENTRY(PbsPlan<1, 2, 3, 0, 2, 64, 1, 3>),    // M = 128
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 3, 0, 2, 128, 1, 3>),   // M = 256
    PBS_PLAN_ENTRY(PbsPlan<1, 4, 3, 0, 2, 128, 1, 3>),   // M = 512
    PBS_PLAN_ENTRY(PbsPlan<1, 4, 4, 0, 2, 128, 1, 3>),   // M = 1024
};
const PbsPlanEntry* pbs_plans_c(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
