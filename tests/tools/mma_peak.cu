// mma_peak.cu -- micro-benchmark: issue rate of the legacy warp-level tensor path (mma.sync) on one B200 for TF32
// (m16n8k8) and BF16 (m16n8k16) with FP32 accumulators, and the rounding of the FP32 accumulation inside the MMA.
// Question it answers: could update_e_b_dynamic's Toeplitz tiles run as error-compensated
// 3xTF32 products on the tensor pipe faster than the 9 FFMA per pair on the CUDA cores?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_peak mma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE, int NACC> __global__ void __launch_bounds__(256) k(float* out, uint32_t av, uint32_t bv, int iters) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    uint32_t a[4] = {av, av + threadIdx.x, av, av}, b[2] = {bv, bv};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) mma_tf32(acc[i], a, b);
            else mma_bf16(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D = A*B + C with A*B = 1.5 in element (0,0) and C = 2^24: the exact sum 16777217.5 rounds to 16777218 (nearest) or
// 16777216 (toward zero).  Also a chain of 4096 accumulations of 1 + 2^-12 ... into a running sum, vs FP32 FMA.
__global__ void k_round(float* out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    if (g == 0 && t == 0) a[0] = __float_as_uint(1.5f);  // A[0][0]
    if (g == 0 && t == 0) b[0] = __float_as_uint(1.0f);  // B[0][0]
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane == 0) d[0] = 16777216.0f;
    mma_tf32(d, a, b);
    if (lane == 0) out[0] = d[0];
    // negative side
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane == 0) e[0] = -16777216.0f;
    uint32_t a2[4] = {0, 0, 0, 0};
    if (g == 0 && t == 0) a2[0] = __float_as_uint(-1.5f);
    mma_tf32(e, a2, b);
    if (lane == 0) out[1] = e[0];
    // chain: sum of 4096 terms (1 + k*2^-10 style values that need all 24 bits of the running sum)
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    float ref = 0.f;
    for (int k = 0; k < 4096; k++) {
        const float v = 1.0f + (float)(k % 7) * 0.125f;  // exactly representable in TF32
        uint32_t ak[4] = {0, 0, 0, 0};
        if (g == 0 && t == 0) ak[0] = __float_as_uint(v);
        uint32_t bk[2] = {0, 0};
        if (g == 0 && t == 0) bk[0] = __float_as_uint(0.3330078125f);  // 341/1024, exact in TF32
        mma_tf32(c, ak, bk);
        ref = fmaf(v, 0.3330078125f, ref);
    }
    if (lane == 0) { out[2] = c[0]; out[3] = ref; }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    const int blocks = sms * 4, threads = 256, iters = 2048;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int nacc = 4; nacc <= 16; nacc *= 2) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; rep++) {
                cudaEventRecord(e0);
                if (mode == 0 && nacc == 4) k<0, 4><<<blocks, threads>>>(out, 0x3f800000u, 0x3f000000u, iters);
                if (mode == 0 && nacc == 8) k<0, 8><<<blocks, threads>>>(out, 0x3f800000u, 0x3f000000u, iters);
                if (mode == 0 && nacc == 16) k<0, 16><<<blocks, threads>>>(out, 0x3f800000u, 0x3f000000u, iters);
                if (mode == 1 && nacc == 4) k<1, 4><<<blocks, threads>>>(out, 0x3f803f80u, 0x3f003f00u, iters);
                if (mode == 1 && nacc == 8) k<1, 8><<<blocks, threads>>>(out, 0x3f803f80u, 0x3f003f00u, iters);
                if (mode == 1 && nacc == 16) k<1, 16><<<blocks, threads>>>(out, 0x3f803f80u, 0x3f003f00u, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double mmas = (double)blocks * (threads / 32) * iters * nacc;
            const double flop = mmas * (mode == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16);
            printf("{\"mode\": \"%s\", \"independent_accumulators\": %d, \"ms\": %.4f, \"mma_per_s\": %.4e, \"tflops\": %.1f, "
                   "\"cycles_per_mma_per_subpartition_at_1965MHz\": %.2f, \"sms\": %d}\n",
                   mode ? "mma.sync m16n8k16 bf16" : "mma.sync m16n8k8 tf32", nacc, best, mmas / (best * 1e-3), flop / (best * 1e-3) / 1e12,
                   1.965e9 * (best * 1e-3) * sms * 4 / mmas, sms);
        }
    }
    k_round<<<1, 32>>>(out);
    float h[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"rounding\": {\"2^24 + 1.5\": %.1f, \"-2^24 - 1.5\": %.1f, \"nearest_would_be\": 16777218.0, \"chain_4096_mma\": %.3f, "
           "\"chain_4096_fmaf\": %.3f}}\n", h[0], h[1], h[2], h[3]);
    return cudaGetLastError() != cudaSuccess;
}
