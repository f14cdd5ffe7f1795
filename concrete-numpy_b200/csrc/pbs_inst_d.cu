// Explicit instantiations of the bootstrap kernels: latency plans -- larger clusters for polynomial sizes the
// throughput plans run on 1..4 CTAs; picked for small batches (sequential circuits such as the key-value
// database insert chain), where spreading one ciphertext over more SMs shortens the blind rotation.
#include "pbs_kernels.cuh"

static const PbsPlanEntry g_entries[] = {
    //             C  logR1 RA RB RF  NT  minb kmax
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 3, 3, 3, 128, 2, 1>),   // N = 8192,  L1 = 512
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 3, 3, 3, 128, 2, 1>),   // N = 4096,  L1 = 512
    PBS_PLAN_ENTRY(PbsPlan<8, 3, 3, 2, 3, 64, 2, 1>),    // N = 4096,  L1 = 256
    PBS_PLAN_ENTRY(PbsPlan<2, 1, 3, 3, 3, 128, 2, 1>),   // N = 2048,  L1 = 512
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 3, 2, 3, 64, 2, 1>),    // N = 2048,  L1 = 256
    PBS_PLAN_ENTRY(PbsPlan<2, 1, 3, 2, 3, 64, 2, 1>),    // N = 1024,  L1 = 256
    PBS_PLAN_ENTRY(PbsPlan<4, 2, 4, 0, 3, 64, 2, 1>),    // N = 1024,  L1 = 128
};
const PbsPlanEntry* pbs_plans_d(int* count) {
    *count = (int)(sizeof(g_entries) / sizeof(g_entries[0]));
    return g_entries;
}
