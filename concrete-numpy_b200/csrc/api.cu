// C-ABI housekeeping: error text, version, device probes, fp64 peak micro-benchmark.
#include "common.cuh"

thread_local char g_fhe_err[512] = {0};

// DFMA throughput probe: the roofline denominator for the PBS (MEASURED_PEAKS.json has
// no fp64 figure).  Each thread runs 8 independent FMA chains.
__global__ void __launch_bounds__(256) k_dfma_probe(double* sink, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

extern "C" {

const char* fhe_last_error(void) { return g_fhe_err; }
int fhe_abi_version(void) { return 1; }

int fhe_device_info(int* sm_count, int* cc_major, int* cc_minor, long* total_mem) {
    int dev = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    FHE_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = (long)p.totalGlobalMem;
    return 0;
}

// returns measured fp64 FMA throughput in TFLOP/s (2 flops per FMA) through *tflops
static int measure_fp64_peak_impl(double* tflops, cudaStream_t st, double* sink, cudaEvent_t e0, cudaEvent_t e1, int sms) {
    const int iters = 1 << 16, blocks = sms * 8;
    k_dfma_probe<<<blocks, 256, 0, st>>>(sink, iters, 1.0);   // warm-up
    FHE_CHECK_LAUNCH();
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        FHE_CUDA(cudaEventRecord(e0, st));
        k_dfma_probe<<<blocks, 256, 0, st>>>(sink, iters, 1.0 + rep);
        FHE_CUDA(cudaEventRecord(e1, st));
        FHE_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    *tflops = best;
    return 0;
}

int fhe_measure_fp64_peak(double* tflops, void* stream) {
    if (!tflops) FHE_FAIL(2, "measure_fp64_peak: null output");
    int dev = 0, sms = 0;
    FHE_CUDA(cudaGetDevice(&dev));
    FHE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    cudaError_t ce = cudaMalloc(&sink, 8);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce != cudaSuccess) {
        snprintf(g_fhe_err, sizeof(g_fhe_err), "measure_fp64_peak: %s", cudaGetErrorString(ce));
        rc = 1000 + (int)ce;
    } else
        rc = measure_fp64_peak_impl(tflops, (cudaStream_t)stream, sink, e0, e1, sms);
    if (e0) cudaEventDestroy(e0);          // released on every path
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    return rc;
}

}  // extern "C"
