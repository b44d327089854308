// packed_check.cu -- device check: collide_two_species / rho_m2 (packed FP32, lattice.cuh) against the scalar functions
// (rho_u, f_eq, forcing_terms + SRT/TRT relaxation as written in stream_collide.cuh) on random inputs, bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -fmad=false -o packed_check tests/tools/packed_check.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../../ionsolver_b200/csrc/lattice.cuh"
using namespace ion;

template <int VS> __device__ void scalar_species(float* f, float rho, float ux, float uy, float uz, float fx, float fy, float fz, float w, bool vf,
                                                 bool force_always, bool trt, bool is_e) {
    constexpr int QQ = VSet<VS>::Q;
    float Fin[QQ], feq[QQ];
    const float c_tau = fmaf(w, -0.5f, 1.0f);
    if (vf || force_always) forcing_terms<VS>(ux, uy, uz, fx, fy, fz, Fin);
    else for (int i = 0; i < QQ; i++) Fin[i] = 0.0f;
    f_eq<VS>(rho, ux, uy, uz, feq);
    if (!trt) {
        for (int i = 0; i < QQ; i++) {
            const float Fi = (vf || force_always) ? Fin[i] * c_tau : Fin[i];
            f[i] = is_e ? feq[i] : fmaf(1.0f - w, f[i], fmaf(w, feq[i], Fi));
        }
    } else {
        const float wp = w, wm = 1.0f / (0.1875f / (1.0f / w - 0.5f) + 0.5f);
        if (vf) {
            const float c_taup = fmaf(wp, -0.25f, 0.5f), c_taum = fmaf(wm, -0.25f, 0.5f);
            float Fib[QQ];
            Fib[0] = Fin[0];
            for (int i = 1; i < QQ; i += 2) { Fib[i] = Fin[i + 1]; Fib[i + 1] = Fin[i]; }
            for (int i = 0; i < QQ; i++) Fin[i] = fmaf(c_taup, Fin[i] + Fib[i], c_taum * (Fin[i] - Fib[i]));
        }
        float fhb[QQ], feb[QQ];
        fhb[0] = f[0]; feb[0] = feq[0];
        for (int i = 1; i < QQ; i += 2) { fhb[i] = f[i + 1]; fhb[i + 1] = f[i]; feb[i] = feq[i + 1]; feb[i + 1] = feq[i]; }
        for (int i = 0; i < QQ; i++)
            f[i] = is_e ? feq[i] : fmaf(0.5f * wp, feq[i] - f[i] + feb[i] - fhb[i], fmaf(0.5f * wm, feq[i] - feb[i] - f[i] + fhb[i], f[i] + Fin[i]));
    }
}

template <int VS, bool TRT> __global__ void k(const float* in, unsigned* bad, float* dump, int n, bool vf) {
    constexpr int QQ = VSet<VS>::Q;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* p = in + (size_t)t * (2 * QQ + 16);
    float f[QQ], e[QQ], f2[QQ], e2[QQ];
    for (int i = 0; i < QQ; i++) { f[i] = f2[i] = p[i]; e[i] = e2[i] = p[QQ + i]; }
    const float* q = p + 2 * QQ;
    const float w = 1.0f + 0.9f * q[14];
    // moments
    float r, ux, uy, uz, re, uxe, uye, uze;
    rho_u<VS>(f, r, ux, uy, uz);
    rho_u<VS>(e, re, uxe, uye, uze);
    float2 rho2, m2[3];
    rho_m2<VS>(f, e, rho2, m2);
    const float pr[8] = {rho2.x, m2[0].x / rho2.x, m2[1].x / rho2.x, VS == ION_D2Q9 ? 0.0f / rho2.x : m2[2].x / rho2.x,
                         rho2.y, m2[0].y / rho2.y, m2[1].y / rho2.y, VS == ION_D2Q9 ? 0.0f / rho2.y : m2[2].y / rho2.y};
    const float sr[8] = {r, ux, uy, uz, re, uxe, uye, uze};
    for (int i = 0; i < 8; i++)
        if (__float_as_uint(pr[i]) != __float_as_uint(sr[i])) atomicAdd(&bad[0], 1u);
    const bool is_e = false;
    scalar_species<VS>(f, r, q[0], q[1], q[2], q[3], q[4], q[5], w, vf, false, TRT, is_e);
    scalar_species<VS>(e, re, q[6], q[7], q[8], q[9], q[10], q[11], w, vf, true, false, is_e);
    collide_two_species<VS, TRT>(f2, e2, make_float2(r, re), make_float2(q[0], q[6]), make_float2(q[1], q[7]), make_float2(q[2], q[8]),
                                 make_float2(q[3], q[9]), make_float2(q[4], q[10]), make_float2(q[5], q[11]), w, vf, is_e);
    for (int i = 0; i < QQ; i++) {
        if (__float_as_uint(f[i]) != __float_as_uint(f2[i])) { if (atomicAdd(&bad[1], 1u) == 0u) { dump[0] = (float)i; dump[1] = f[i]; dump[2] = f2[i]; dump[3] = (float)t; } }
        if (__float_as_uint(e[i]) != __float_as_uint(e2[i])) { if (atomicAdd(&bad[2], 1u) == 0u) { dump[4] = (float)i; dump[5] = e[i]; dump[6] = e2[i]; dump[7] = (float)t; } }
    }
}

template <int VS, bool TRT> int run(const char* name, bool vf) {
    constexpr int QQ = VSet<VS>::Q;
    const int n = 1 << 18, stride = 2 * QQ + 16;
    float* h = (float*)malloc(sizeof(float) * n * stride);
    srand(1234);
    for (int t = 0; t < n; t++) {
        float* p = h + (size_t)t * stride;
        for (int i = 0; i < 2 * QQ; i++) p[i] = ((rand() / (float)RAND_MAX) - 0.5f) * ((t & 3) == 0 ? 1e-3f : (t & 3) == 1 ? 1e-6f : 0.05f);
        for (int i = 0; i < 16; i++) p[2 * QQ + i] = ((rand() / (float)RAND_MAX) - 0.5f) * (i < 3 || (i >= 6 && i < 9) ? 0.3f : 1e-3f);
        if ((t & 15) == 7) { p[2 * QQ + 0] = 0.1f; p[2 * QQ + 1] = -0.1f; p[2 * QQ + 3] = 1e-4f; p[2 * QQ + 4] = -1e-4f; }  // cancelling sums
        if ((t & 15) == 9) { for (int i = 0; i < 12; i++) p[2 * QQ + i] = 0.0f; }
        p[2 * QQ + 14] = rand() / (float)RAND_MAX;
    }
    float *d, *dump; unsigned* bad;
    cudaMalloc(&d, sizeof(float) * n * stride); cudaMalloc(&bad, 64); cudaMalloc(&dump, 32);
    cudaMemcpy(d, h, sizeof(float) * n * stride, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 64); cudaMemset(dump, 0, 32);
    k<VS, TRT><<<(n + 127) / 128, 128>>>(d, bad, dump, n, vf);
    unsigned hb[16]; float hd[8];
    cudaMemcpy(hb, bad, 64, cudaMemcpyDeviceToHost); cudaMemcpy(hd, dump, 32, cudaMemcpyDeviceToHost);
    printf("%s vf=%d: moments mismatches %u, gas %u, electron %u", name, (int)vf, hb[0], hb[1], hb[2]);
    if (hb[1]) printf("  first gas: i=%d scalar %.9g packed %.9g (sample %d)", (int)hd[0], hd[1], hd[2], (int)hd[3]);
    if (hb[2]) printf("  first electron: i=%d scalar %.9g packed %.9g (sample %d)", (int)hd[4], hd[5], hd[6], (int)hd[7]);
    printf("\n");
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err));
    free(h); cudaFree(d); cudaFree(bad); cudaFree(dump);
    return (hb[0] || hb[1] || hb[2]) ? 1 : 0;
}

int main() {
    int rc = 0;
    rc |= run<ION_D3Q19, false>("D3Q19 SRT", true);
    rc |= run<ION_D3Q19, false>("D3Q19 SRT", false);
    rc |= run<ION_D3Q19, true>("D3Q19 TRT", true);
    rc |= run<ION_D3Q19, true>("D3Q19 TRT", false);
    rc |= run<ION_D3Q27, false>("D3Q27 SRT", true);
    rc |= run<ION_D3Q27, true>("D3Q27 TRT", true);
    rc |= run<ION_D3Q15, false>("D3Q15 SRT", true);
    rc |= run<ION_D2Q9, true>("D2Q9 TRT", true);
    return rc;
}
