// fp32_peak.cu -- micro-benchmark: FP32 issue peak of one B200 with scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// Gives the compute roofline for update_e_b_dynamic (an FMA-issue-bound kernel), since MEASURED_PEAKS.json only has
// HBM and BF16 tensor numbers.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
    // 16 independent accumulator pairs per thread
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) {
                acc[i].x = fmaf(acc[i].x, x.x, y.x);
                acc[i].y = fmaf(acc[i].y, x.y, y.y);
            } else {
                acc[i] = __ffma2_rn(acc[i], x, y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
            else k<1><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double fma = (double)blocks * threads * iters * 32.0;
        printf("{\"mode\": \"%s\", \"ms\": %.4f, \"fma_per_s\": %.4e, \"tflops\": %.2f, \"sms\": %d}\n", mode ? "ffma2" : "ffma", best,
               fma / (best * 1e-3), 2.0 * fma / (best * 1e-3) / 1e12, sms);
    }
    return cudaGetLastError() != cudaSuccess;
}
