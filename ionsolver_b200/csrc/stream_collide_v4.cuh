// stream_collide_v4.cuh -- plain (non-MHD) stream_collide with FOUR consecutive x cells per thread and 128-bit (FP32) / 64-bit
// (FP16S, FP16C) vector accesses.  Same arithmetic, same order, same results as k_stream_collide (sim.cl:465-529,680-758); what
// changes is how the DDFs travel:
//   * every "self" slot of the four cells is ONE aligned vector load and ONE aligned vector store;
//   * a neighbour slot with c_x = 0 is one aligned vector load; with c_x = +-1 the four values are the aligned vector of the
//     neighbour row shifted by one element, i.e. that vector plus one scalar (the periodic wrap only ever affects the scalar);
//   * pushes to neighbours with c_x = 0 are vector stores, with c_x = +-1 four scalar stores (their targets straddle two
//     threads' vectors; solid cells must not store at all, so no shuffle-merged vector store is attempted).
// Per four cells of D3Q19 that is 29 load and 42 store instructions instead of 76 + 76, a quarter of the address arithmetic, and
// four times the bytes in flight per warp.  Requires nx % 4 == 0 (rows then start 16-byte aligned); anything else, and all MHD
// configurations (2Q+7+6 live values per cell), stay on the one-cell kernel.
#pragma once
#include "stream_collide.cuh"

namespace ion {

// ---- four consecutive storage elements ----
template <typename S> struct VecIO;
template <> struct VecIO<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecIO<uint16_t> {
    static __device__ __forceinline__ void ld(const uint16_t* p, uint16_t (&v)[4]) {
        asm volatile("ld.volatile.global.v4.u16 {%0,%1,%2,%3}, [%4];" : "=h"(v[0]), "=h"(v[1]), "=h"(v[2]), "=h"(v[3]) : "l"(p));
    }
    static __device__ __forceinline__ void st(uint16_t* p, const uint16_t (&v)[4]) {
        *reinterpret_cast<uint2*>(p) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
    }
};
template <int FP> __device__ __forceinline__ void vload4(const typename Codec<FP>::store_t* p, float (&out)[4]) {
    typename Codec<FP>::store_t raw[4];
    VecIO<typename Codec<FP>::store_t>::ld(p, raw);
#pragma unroll
    for (int c = 0; c < 4; c++) out[c] = Codec<FP>::dec(raw[c]);
}
template <int FP> __device__ __forceinline__ void vstore4(typename Codec<FP>::store_t* p, const float (&in)[4]) {
    typename Codec<FP>::store_t raw[4];
#pragma unroll
    for (int c = 0; c < 4; c++) raw[c] = Codec<FP>::enc(in[c]);
    VecIO<typename Codec<FP>::store_t>::st(p, raw);
}

// One fluid cell between load and store, plain path: the body of k_stream_collide for MHD = false (sim.cl:496-529,680-755).
template <int VS, bool TRT>
__device__ __forceinline__ void plain_cell(const KArgs& a, uint32_t n, uint8_t flagsn, float* fhn, float fx, float fy, float fz) {
    constexpr int QQ = VSet<VS>::Q;
    const uint64_t N = a.N;
    const uint8_t bo = flagsn & ION_TYPE_BO;
    const bool eqb = (a.ext & ION_EXT_EQUILIBRIUM_BOUNDARIES) != 0u;
    const bool vf = (a.ext & ION_EXT_VOLUME_FORCE) != 0u;
    const bool is_e = eqb && bo == ION_TYPE_E;
    float rhon, uxn, uyn, uzn;
    if (is_e) {  // sim.cl:503-507
        rhon = a.rho[n];
        uxn = a.u[n];
        uyn = a.u[N + n];
        uzn = a.u[2ull * N + n];
    } else {
        rho_u<VS>(fhn, rhon, uxn, uyn, uzn);
    }
    float fxn = fx, fyn = fy, fzn = fz;  // sim.cl:513
    float Fin[QQ], feq[QQ];
    const float w = a.w;
    const float c_tau = fmaf(w, -0.5f, 1.0f);  // sim.cl:519
    if (a.ext & ION_EXT_FORCE_FIELD) {          // sim.cl:522-528
        fxn += a.F[n];
        fyn += a.F[N + n];
        fzn += a.F[2ull * N + n];
    }
    if (vf) {  // sim.cl:680-685
        const float rho2 = 0.5f / rhon;
        uxn = clampf(fmaf(fxn, rho2, uxn), -ION_DEF_C, ION_DEF_C);
        uyn = clampf(fmaf(fyn, rho2, uyn), -ION_DEF_C, ION_DEF_C);
        uzn = clampf(fmaf(fzn, rho2, uzn), -ION_DEF_C, ION_DEF_C);
        forcing_terms<VS>(uxn, uyn, uzn, fxn, fyn, fzn, Fin);
    } else {  // sim.cl:687-690
        uxn = clampf(uxn, -ION_DEF_C, ION_DEF_C);
        uyn = clampf(uyn, -ION_DEF_C, ION_DEF_C);
        uzn = clampf(uzn, -ION_DEF_C, ION_DEF_C);
#pragma unroll
        for (int i = 0; i < QQ; i++) Fin[i] = 0.0f;
    }
    if ((a.ext & ION_EXT_UPDATE_FIELDS) && !is_e) {  // sim.cl:694-710
        a.rho[n] = rhon;
        a.u[n] = uxn;
        a.u[N + n] = uyn;
        a.u[2ull * N + n] = uzn;
    }
    f_eq<VS>(rhon, uxn, uyn, uzn, feq);  // sim.cl:712
    if (!TRT) {                          // sim.cl:714-723
#pragma unroll
        for (int i = 0; i < QQ; i++) {
            const float Fi = vf ? Fin[i] * c_tau : Fin[i];
            fhn[i] = is_e ? feq[i] : fmaf(1.0f - w, fhn[i], fmaf(w, feq[i], Fi));
        }
    } else {  // sim.cl:725-755
        const float wm = 1.0f / (0.1875f / (1.0f / w - 0.5f) + 0.5f);
        {   // direction 0 is its own opposite (fhb[0] = fhn[0], feb[0] = feq[0], Fib[0] = Fin[0])
            float F0 = Fin[0];
            if (vf) F0 = fmaf(fmaf(w, -0.25f, 0.5f), Fin[0] + Fin[0], fmaf(wm, -0.25f, 0.5f) * (Fin[0] - Fin[0]));
            const float g = fmaf(0.5f * w, feq[0] - fhn[0] + feq[0] - fhn[0], fmaf(0.5f * wm, feq[0] - feq[0] - fhn[0] + fhn[0], fhn[0] + F0));
            fhn[0] = is_e ? feq[0] : g;
        }
#pragma unroll
        for (int i = 1; i < QQ; i += 2) {
            float ga = fhn[i], gb = fhn[i + 1];
            trt_pair(ga, gb, feq[i], feq[i + 1], Fin[i], Fin[i + 1], w, wm, vf);
            fhn[i] = is_e ? feq[i] : ga;
            fhn[i + 1] = is_e ? feq[i + 1] : gb;
        }
    }
}

// 6 resident blocks of 64 threads per SM (168 registers, no spills for Q <= 19).  Also measured: D3Q27 at 4 blocks / 254 registers
// 0.759 (one-cell kernel 0.823), FP16S / FP16C at 8 blocks / 128 registers 0.503 / 0.393 (6 blocks: 0.507 / 0.381) -- neither beats
// the one-cell kernels, so only FP32 with Q <= 19 uses this kernel by default.
template <int VS, int FP, bool TRT, bool ODD>
__global__ void __launch_bounds__(64, 6)
k_stream_collide_v4(const __grid_constant__ KArgs a, const float fx, const float fy, const float fz) {
    constexpr int QQ = VSet<VS>::Q;
    typedef typename Codec<FP>::store_t S;
    const uint32_t nx = a.nx, ny = a.ny, nz = a.nz;
    const uint32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u, y = blockIdx.y, z = blockIdx.z + a.z_off;
    if (x0 >= nx) return;
    if (((a.dy > 1u) & (y == 0u || y >= ny - 1u)) || ((a.dz > 1u) & (z == 0u || z >= nz - 1u))) return;  // halo rows, sim.cl:145-148
    const uint64_t N = a.N;
    const uint32_t nxy = nx * ny;
    const uint32_t y0 = y * nx, yp = ((y + 1u == ny) ? 0u : y + 1u) * nx, ym = ((y == 0u) ? ny - 1u : y - 1u) * nx;
    const uint32_t z0 = z * nxy, zp = ((z + 1u == nz) ? 0u : z + 1u) * nxy, zm = ((z == 0u) ? nz - 1u : z - 1u) * nxy;
    const uint32_t xl = x0 == 0u ? nx - 1u : x0 - 1u, xr = x0 + 4u == nx ? 0u : x0 + 4u;  // periodic wrap, sim.cl:250-252
    const uint32_t n0 = x0 + y0 + z0;
    S* const p = reinterpret_cast<S*>(a.fi);
    auto row = [&](int i) -> uint32_t {  // row of the neighbour in direction i (D2Q9 ignores z like the reference, sim.cl:265-268)
        const int cy = cvel<VS>(1, i), cz = VS == ION_D2Q9 ? 0 : cvel<VS>(2, i);
        return (cy > 0 ? yp : cy < 0 ? ym : y0) + (cz > 0 ? zp : cz < 0 ? zm : z0);
    };

    // ---- streaming part 2: esoteric-pull loads (load_f, sim.cl:234-240) ----
    float f[4][QQ];
    {
        float v[4];
        vload4<FP>(p + n0, v);
#pragma unroll
        for (int c = 0; c < 4; c++) f[c][0] = v[c];
    }
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        const int sa = ODD ? i : i + 1, sb = ODD ? i + 1 : i;
        float v[4];
        vload4<FP>(p + (uint64_t)sa * N + n0, v);
#pragma unroll
        for (int c = 0; c < 4; c++) f[c][i] = v[c];
        const S* q = p + (uint64_t)sb * N + row(i);
        const int cx = cvel<VS>(0, i);
        vload4<FP>(q + x0, v);
        if (cx == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) f[c][i + 1] = v[c];
        } else if (cx > 0) {
            f[0][i + 1] = v[1]; f[1][i + 1] = v[2]; f[2][i + 1] = v[3];
            f[3][i + 1] = Codec<FP>::dec(ddf_ld(q + xr));
        } else {
            f[1][i + 1] = v[0]; f[2][i + 1] = v[1]; f[3][i + 1] = v[2];
            f[0][i + 1] = Codec<FP>::dec(ddf_ld(q + xl));
        }
    }
    uint32_t fl4;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(fl4) : "l"(a.flags + n0));  // requested last, see k_stream_collide

    // ---- collide ----
    bool act[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t x = x0 + (uint32_t)c;
        const uint8_t fl = (uint8_t)(fl4 >> (8 * c));
        act[c] = !((a.dx > 1u) & (x == 0u || x >= nx - 1u)) && (fl & ION_TYPE_BO) != ION_TYPE_S;  // sim.cl:485-488 (quirk Q1)
        if (act[c]) plain_cell<VS, TRT>(a, n0 + (uint32_t)c, fl, f[c], fx, fy, fz);
    }
    const bool all4 = act[0] && act[1] && act[2] && act[3];

    // ---- streaming part 1: esoteric-pull stores (store_f, sim.cl:241-247); inactive cells store nothing ----
    auto store_self = [&](S* dst, int slot) {
        if (all4) {
            const float v[4] = {f[0][slot], f[1][slot], f[2][slot], f[3][slot]};
            vstore4<FP>(dst, v);
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (act[c]) ddf_st(dst + c, Codec<FP>::enc(f[c][slot]));
        }
    };
    store_self(p + n0, 0);
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        const int sa = ODD ? i : i + 1, sb = ODD ? i + 1 : i;
        S* q = p + (uint64_t)sb * N + row(i);
        const int cx = cvel<VS>(0, i);
        if (cx == 0) {
            store_self(q + x0, i);  // same column: an aligned vector of the neighbour row
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t xt = cx > 0 ? (c < 3 ? x0 + (uint32_t)c + 1u : xr) : (c > 0 ? x0 + (uint32_t)c - 1u : xl);
                if (act[c]) ddf_st(q + xt, Codec<FP>::enc(f[c][i]));
            }
        }
        store_self(p + (uint64_t)sa * N + n0, i + 1);
    }
}

// Where it is used.  Measured at 256^3 / 512^3 (profiles/r1_stream_collide_ab.md, fraction of the HBM copy peak, one-cell -> four-cell):
// D3Q19 FP32 SRT 0.869 -> 0.972 (512^3: 0.867 -> 0.969), TRT 0.723 -> 0.761; FP16S unchanged (0.51), FP16C slower (0.447 -> 0.381:
// four encoders per thread), D3Q27 FP32 slower (0.823 -> 0.675: 4 x 27 live values against the 168-register cap), 64^3 slower (8.5 us
// -> 14.9 us with 16 threads per x row; 12.3 us with 16 x 4 thread blocks: 65 536 threads are less than one wave of the GPU).  Default: FP32, Q <= 19, nx >= 256.  ION_SC_VEC=1 forces it wherever nx % 4 == 0 (the parity tests use this to
// run every configuration through it), ION_SC_VEC=0 turns it off.
inline int sc_vec_mode() {
    static const int mode = getenv("ION_SC_VEC") ? (atoi(getenv("ION_SC_VEC")) != 0 ? 1 : 0) : -1;
    return mode;
}

template <int VS>
inline bool launch_stream_collide_v4(const KArgs& a, int fp, bool trt, uint64_t t, float fx, float fy, float fz, cudaStream_t s) {
    const int mode = sc_vec_mode();
    if (mode == 0 || (a.nx % 4u) != 0u) return false;
    if (mode < 0 && !(fp == ION_FP32 && VSet<VS>::Q <= 19 && a.nx >= 256u)) return false;
    const unsigned threads_x = a.nx / 4u;
    unsigned block = ((threads_x + 31u) / 32u) * 32u;
    if (block > 64u) block = 64u;
    const dim3 grid((threads_x + block - 1u) / block, a.ny, a.z_cnt ? a.z_cnt : a.nz);
    const bool odd = (t & 1ull) != 0ull;
#define ION_V4_CASE(FPV, TRTV)                                                                          \
    if (fp == FPV && trt == TRTV) {                                                                      \
        if (odd) k_stream_collide_v4<VS, FPV, TRTV, true><<<grid, block, 0, s>>>(a, fx, fy, fz);        \
        else k_stream_collide_v4<VS, FPV, TRTV, false><<<grid, block, 0, s>>>(a, fx, fy, fz);           \
        return true;                                                                                     \
    }
    ION_V4_CASE(ION_FP32, false)
    ION_V4_CASE(ION_FP32, true)
    ION_V4_CASE(ION_FP16S, false)
    ION_V4_CASE(ION_FP16S, true)
    ION_V4_CASE(ION_FP16C, false)
    ION_V4_CASE(ION_FP16C, true)
#undef ION_V4_CASE
    return false;
}

}  // namespace ion
