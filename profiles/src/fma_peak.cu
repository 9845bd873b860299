// fma_peak.cu -- measured FP32 / FP64 FMA peak of the CUDA cores on this GPU (SURVEY.md section 8d asks for a micro-benchmark next to the HBM figure;
// bench.py's roofline_flops.peak reads the result from profiles/r2/fma_peak.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/fma_peak profiles/src/fma_peak.cu   (make -C 3drvo-marl-collisionavoidance_b200/csrc fma_peak)
// Every thread runs ILP independent FMA chains; the grid covers every SM with the maximum number of resident warps.  Timed with CUDA events, best of 5.
#include <cuda_runtime.h>
#include <cstdio>

template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_kernel(T *out, int iters, T a, T b) {
    T x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (T)(threadIdx.x + i) * (T)1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = x[i] * a + b;      // contracted to one FMA per element (-fmad=true, the default here)
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == (T)123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

template <typename T, int ILP>
double run(int sms, int iters) {
    T *out; cudaMalloc(&out, sizeof(T) * sms * 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        fma_kernel<T, ILP><<<sms * 8, 256>>>(out, iters, (T)0.999, (T)1e-4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * ILP * (double)iters * sms * 8 * 256 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double f32 = run<float, 8>(p.multiProcessorCount, 1 << 16);
    const double f64 = run<double, 8>(p.multiProcessorCount, 1 << 14);
    if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"fp32_tflops\": %.3f, \"fp64_tflops\": %.3f, "
                "\"how\": \"8 independent FMA chains per thread, 8 x 256 threads per SM, CUDA events, best of 5 (profiles/src/fma_peak.cu)\"}\n",
                p.name, p.multiProcessorCount, p.clockRate / 1e3, f32, f64);
    return 0;
}
