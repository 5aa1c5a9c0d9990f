// fp64_peak.cu -- DFMA-chain microbenchmark: the FP64 (non-tensor) peak of the device, which MEASURED_PEAKS.json does not
// hold (SURVEY.md section 8d asks for it to show that the FP64 roofline is not the binding one for this path).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += x[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chain alive
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* d;
    cudaMalloc(&d, sizeof(double) * 1 << 20);
    const int iters = 4096, blocks = p.multiProcessorCount * 8;
    constexpr int CH = 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<CH><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double dfma = (double)blocks * 256 * CH * iters;
        const double rate = dfma / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.4g, \"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f}\n", p.name,
           p.multiProcessorCount, best, 2 * best / 1e12, best / (p.multiProcessorCount * (double)p.clockRate * 1e3));
    return 0;
}
