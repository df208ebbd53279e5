// fp64_probe.cu -- what the FP64 pipe and the shared-memory data pipe of one B200 SM sustain (denominators for the
// "which pipe bounds sweep A" question in profiles/r1b_*).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_probe.cu -o fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 1e-9 + q;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) x[q] = fma(x[q], a, b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += x[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every thread reads 8-byte words of a 16 KB shared array, conflict-free (lane-contiguous), 8 independent loads in flight
__global__ void lds_kernel(double* out, int iters) {
    __shared__ double sm[2048];
    for (int q = threadIdx.x; q < 2048; q += blockDim.x) sm[q] = q;
    __syncthreads();
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int base = threadIdx.x & 255;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) s[q] += sm[(base + q * 256 + i) & 2047];
    }
    double t = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) t += s[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out;
    const int threads = 256, blocks = sms * 8;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)blocks * threads * 8.0 * iters;
        std::printf("dfma: %.3f ms  %.2f TFLOP/s  %.1f DFMA lanes/clk/SM at %d MHz (nominal)\n", ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (khz * 1e3),
                    khz / 1000);
    }
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        lds_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double lds = (double)blocks * threads * 8.0 * iters;      // 8-byte loads (each also feeds one DADD)
        std::printf("lds.64: %.3f ms  %.1f bytes/clk/SM  (%.2f warp-loads/clk/SM)\n", ms, 8 * lds / (ms * 1e-3) / sms / (khz * 1e3),
                    lds / 32 / (ms * 1e-3) / sms / (khz * 1e3));
    }
    std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
