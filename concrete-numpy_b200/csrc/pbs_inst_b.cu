// Explicit instantiations of the bootstrap kernels: 4- and 2-CTA cluster plans and the large single-CTA plans.
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 4, 3, 256, 1, 1>),   // L1 = 2048 (N = 16384)
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 3, 3, 128, 2, 1>),   // L1 = 1024 (N = 8192)
    PBS_PLAN_ENTRY(PbsPlan<2, 1, 4, 4, 3, 256, 1, 1>),   // L1 = 2048 (N = 8192)
    PBS_PLAN_ENTRY(PbsPlan<2, 1, 4, 3, 3, 128, 2, 1>),   // L1 = 1024 (N = 4096)
    PBS_PLAN_ENTRY(PbsPlan<1, 4, 4, 0, 3, 256, 1, 1>),   // M = 2048 (N = 4096)
    // M = 2048 on 512 threads at 128 registers (one item per thread, radix 8 x 8 x 4 x 8): k = 1 bootstraps run k_pbs_group
    // with a single slot -- 16 warps on one ciphertext, every twiddle in tensor memory
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 3, 2, 3, 512, 1, 1, 0, 1>),
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 4, 0, 3, 128, 2, 1>),   // M = 1024 (N = 2048): 2 CTAs/SM at 255 registers beat 3 with spills (+25 %)
    // M = 1024, 2 x 256 threads per SM at 128 registers (16 warps per SM); one- and two-level bootstraps run k_pbs_group: ONE
    // CTA per SM with two ciphertext slots, twiddles in tensor memory
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 4, 0, 3, 256, 2, 1, 0, 2>),
    // M = 1024 with >= 3 levels (one CTA per SM by shared memory): 8 warps instead of 4; k = 1 runs k_pbs_group with one slot
    PBS_PLAN_ENTRY(PbsPlan<1, 3, 4, 0, 3, 256, 1, 1, 0, 1>),
};
const PbsPlanEntry* pbs_plans_b(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
