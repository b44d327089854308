// packed_ops.cu -- are add/mul/fma.rn.f32x2 bit-identical, lane by lane, to the scalar IEEE operations?  (incl. denormals)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(const float* in, unsigned* bad, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float a = in[6 * t], b = in[6 * t + 1], c = in[6 * t + 2], d = in[6 * t + 3], e = in[6 * t + 4], f = in[6 * t + 5];
    const float2 A = make_float2(a, d), B = make_float2(b, e), C = make_float2(c, f);
    const float2 m = __fmul2_rn(A, B), s = __fadd2_rn(A, B), q = __ffma2_rn(A, B, C);
    const float2 nq = __ffma2_rn(make_float2(-A.x, -A.y), B, C);
    if (__float_as_uint(m.x) != __float_as_uint(a * b) || __float_as_uint(m.y) != __float_as_uint(d * e)) atomicAdd(&bad[0], 1u);
    if (__float_as_uint(s.x) != __float_as_uint(a + b) || __float_as_uint(s.y) != __float_as_uint(d + e)) atomicAdd(&bad[1], 1u);
    if (__float_as_uint(q.x) != __float_as_uint(fmaf(a, b, c)) || __float_as_uint(q.y) != __float_as_uint(fmaf(d, e, f))) atomicAdd(&bad[2], 1u);
    if (__float_as_uint(nq.x) != __float_as_uint(fmaf(-a, b, c)) || __float_as_uint(nq.y) != __float_as_uint(fmaf(-d, e, f))) atomicAdd(&bad[3], 1u);
}
int main() {
    const int n = 1 << 22;
    float* h = (float*)malloc(sizeof(float) * 6 * n);
    srand(7);
    for (int i = 0; i < 6 * n; i++) {
        const int cls = (i / 6) & 7;
        float v = (rand() / (float)RAND_MAX) - 0.5f;
        if (cls == 1) v *= 1e-20f;
        if (cls == 2) v *= 1e-38f;             // denormal operands / results
        if (cls == 3 && (i % 6) == 2) v *= 1e-7f;
        if (cls == 4) v = (float)((rand() % 7) - 3);
        h[i] = v;
    }
    float* d; unsigned* bad;
    cudaMalloc(&d, sizeof(float) * 6 * n); cudaMalloc(&bad, 16);
    cudaMemcpy(d, h, sizeof(float) * 6 * n, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 16);
    k<<<(n + 255) / 256, 256>>>(d, bad, n);
    unsigned hb[4];
    cudaMemcpy(hb, bad, 16, cudaMemcpyDeviceToHost);
    printf("{\"samples\": %d, \"fmul2_mismatch\": %u, \"fadd2_mismatch\": %u, \"ffma2_mismatch\": %u, \"ffma2_neg_mismatch\": %u}\n", n, hb[0], hb[1], hb[2], hb[3]);
    return 0;
}
