// lattice.cuh -- device-side building blocks shared by all IonSolver-B200 kernels.
//
// Hand-written for sm_100a.  Written from the reference's *behaviour* (file:line citations into
// /root/reference/src/kernels/sim_kernels.cl = "sim.cl"), not from its text: the velocity sets, weights and
// operation ORDER of every floating-point expression are reproduced so that results are bit-identical to the
// reference kernels compiled without contraction (the library is built with -fmad=false; every fused
// multiply-add below is an explicit fmaf, exactly where the reference calls fma()).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/ionsolver_b200.h"

namespace ion {

// ------------------------------------------------------------------------------------------------------
// kernel argument block (plain pointers + the IonParams subset the device needs)
// ------------------------------------------------------------------------------------------------------
struct KArgs {
    void* fi;
    float* rho;
    float* u;
    uint8_t* flags;
    const float* F;
    float* E_stat;
    float* B_stat;
    float* E_dyn;
    float* B_dyn;
    void* fqi;
    void* ei;
    float* Q;
    float* QU_lod;
    float* lod_u;  // deterministic mode: per-cell scaled velocity deposit (3N floats), summed in order by k_lod_deposit_ordered
    float* lod_rep;             // privatised LOD deposit: lod_rep_count replicas of the own finest level (8^depth float4 each)
    uint32_t lod_rep_mask;      // lod_rep_count - 1 (power of two)
    uint32_t lod_rep_entries;   // 8^depth
    const float* E_var;
    void* eti;
    float* Et;
    uint8_t* transfer_p;
    uint8_t* transfer_m;
    uint64_t N;  // DEF_N
    uint32_t nx, ny, nz;
    uint32_t dx, dy, dz, di;
    int32_t ox, oy, oz;
    uint32_t ext;
    float w;
    float ke, kmu, kmu0, kkge, kme, wq, kkbme, keabs;
    uint32_t lod_depth, n_lod, n_lod_own;
    float ecrf;
    // stream_collide on a range of z layers (boundary-layer-first scheduling, ion_enqueue_stream_collide_range): the kernel's z is
    // blockIdx.z + z_off and the launcher's grid is z_cnt layers high (0 = the whole domain)
    uint32_t z_off, z_cnt;
};

__device__ __forceinline__ float sq(float x) { return x * x; }
__device__ __forceinline__ float cb(float x) { return x * x * x; }
// OpenCL clamp(x,lo,hi) = fmin(fmax(x,lo),hi)
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

#define ION_DEF_C 0.57735027f  // lattice speed of sound, domain.rs:807

// ------------------------------------------------------------------------------------------------------
// velocity sets: direction table sim.cl:326-347, weights domain.rs:757-768, transfers types.rs:28-46
// ------------------------------------------------------------------------------------------------------
template <int VS> struct VSet;
template <> struct VSet<ION_D2Q9> { static constexpr int Q = 9, DIM = 2, T = 3; };
template <> struct VSet<ION_D3Q15> { static constexpr int Q = 15, DIM = 3, T = 5; };
template <> struct VSet<ION_D3Q19> { static constexpr int Q = 19, DIM = 3, T = 5; };
template <> struct VSet<ION_D3Q27> { static constexpr int Q = 27, DIM = 3, T = 9; };

// c(axis, i): lattice velocity component; folded to an immediate once the callers' loops are unrolled
template <int VS> __host__ __device__ __forceinline__ constexpr int cvel(int axis, int i) {
    if (VS == ION_D2Q9) {
        constexpr int c[3][9] = {{0, 1, -1, 0, 0, 1, -1, 1, -1}, {0, 0, 0, 1, -1, 1, -1, -1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0}};
        return c[axis][i < 9 ? i : 0];
    } else if (VS == ION_D3Q15) {
        constexpr int c[3][15] = {{0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1},
                                  {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, -1, 1, 1, -1},
                                  {0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, 1, -1, 1, -1}};
        return c[axis][i < 15 ? i : 0];
    } else if (VS == ION_D3Q19) {
        constexpr int c[3][19] = {{0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0},
                                  {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1},
                                  {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1}};
        return c[axis][i < 19 ? i : 0];
    } else {
        constexpr int c[3][27] = {
            {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1},
            {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 1, -1},
            {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1, 1, -1, -1, 1, 1, -1, 1, -1}};
        return c[axis][i < 27 ? i : 0];
    }
}

// weight of direction i.  Class by number of non-zero components: 0 -> W0, 1 -> WS, 2 -> WE, 3 -> WC.
// D3Q27 uses the canonical 8/27, 2/27, 1/54, 1/216 (the reference cannot build D3Q27).
template <int VS> __host__ __device__ __forceinline__ constexpr float wclass(int nonzero) {
    if (VS == ION_D2Q9) return nonzero == 0 ? (1.0f / 2.25f) : nonzero == 1 ? (1.0f / 9.0f) : (1.0f / 36.0f);
    if (VS == ION_D3Q15) return nonzero == 0 ? (1.0f / 4.5f) : nonzero == 1 ? (1.0f / 9.0f) : (1.0f / 72.0f);
    if (VS == ION_D3Q19) return nonzero == 0 ? (1.0f / 3.0f) : nonzero == 1 ? (1.0f / 18.0f) : (1.0f / 36.0f);
    return nonzero == 0 ? (1.0f / 3.375f) : nonzero == 1 ? (1.0f / 13.5f) : nonzero == 2 ? (1.0f / 54.0f) : (1.0f / 216.0f);
}
template <int VS> __host__ __device__ __forceinline__ constexpr int nonzero(int i) {
    return (cvel<VS>(0, i) != 0) + (cvel<VS>(1, i) != 0) + (cvel<VS>(2, i) != 0);
}
template <int VS> __host__ __device__ __forceinline__ constexpr float wdir(int i) { return wclass<VS>(nonzero<VS>(i)); }

// sum_k c_k(i) * a_k evaluated left to right over the non-zero components (bit-equal to the reference's
// `c(i)*ax+c(Q+i)*ay+c(2Q+i)*az` because multiplications by +-1 are exact and adding the +-0 products is an identity)
template <int VS> __device__ __forceinline__ float cdot(int i, float ax, float ay, float az) {
    const int cx = cvel<VS>(0, i), cy = cvel<VS>(1, i), cz = cvel<VS>(2, i);
    float s = 0.0f;
    bool first = true;
    if (cx != 0) { s = cx > 0 ? ax : -ax; first = false; }
    if (cy != 0) { const float t = cy > 0 ? ay : -ay; s = first ? t : s + t; first = false; }
    if (cz != 0) { const float t = cz > 0 ? az : -az; s = first ? t : s + t; first = false; }
    return s;
}

// ------------------------------------------------------------------------------------------------------
// DDF storage codecs: FP32 plain, FP16S (domain.rs:773-776), FP16C (sim.cl:79-90)
// ------------------------------------------------------------------------------------------------------
// float_to_half_custom, sim.cl:79-84 (1-4-11 format).  The reference rounds the mantissa by adding 0x800 to the float's bits,
// then assembles exponent and mantissa fields, with a separate shift-and-round path for results below 2^-14 (~20 integer
// operations).  The same bits come out of ONE multiplication: |x| * 2^-112 re-biases the exponent (127 -> 15) exactly for
// normal results and, rounded TOWARD ZERO, lands results below 2^-14 on the float-denormal grid, which is the reference's
// truncating shift; adding 0x800 to those bits and dropping 12 bits is then its round-half-up in both ranges (mantissa
// carries ripple into the exponent field like in the reference, the 4-bit exponent wraps the same way, +-inf included).
// Checked against the reference formula for every non-NaN float on the CPU (tests/tools/fp16c_encode_check.c: 2^32 - 2^24 inputs,
// host FPU in round-toward-zero mode; a strided run is part of the CPU test suite) and
// on the device in tests/test_gpu_parity.py::test_codecs_exhaustive.  NaN inputs (a broken simulation) give a different
// finite code than the reference's bit shuffle.
__device__ __forceinline__ uint16_t fp16c_encode(float x) {
    const uint32_t a = __float_as_uint(__fmul_rz(fabsf(x), 1.925929944387236e-34f)) + 0x00000800u;  // 2^-112
    return (uint16_t)(((a >> 12) & 0x7FFFu) | ((__float_as_uint(x) >> 16) & 0x8000u));
}
__device__ __forceinline__ uint16_t fp16c_encode_ref(float x) {  // the reference's formula, kept for the codec test hook
    const uint32_t b = __float_as_uint(x) + 0x00000800u;
    const uint32_t e = (b & 0x7F800000u) >> 23;
    const uint32_t m = b & 0x007FFFFFu;
    return (uint16_t)((b & 0x80000000u) >> 16 | (uint32_t)(e > 112u) * ((((e - 112u) << 11) & 0x7800u) | m >> 12) |
                      (uint32_t)((e < 113u) & (e > 100u)) * ((((0x007FF800u + m) >> (124u - e)) + 1u) >> 1));
}
// half_to_float_custom, sim.cl:85-90.  The reference assembles the float from exponent and mantissa fields and normalises
// denormals through an int->float conversion (~14 integer operations).  Identical bits come from one multiplication: placed at
// bit 12, the 15 value bits ARE an IEEE float with the 4-bit exponent in the low exponent bits (e = 0: a float denormal), and
// scaling by 2^112 re-biases it (15 -> 127) exactly -- 11 mantissa bits never round, and the FMUL is not flush-to-zero.
// Checked against the reference formula for all 65 536 codes (tests/test_gpu_parity.py::test_codecs_exhaustive).
__device__ __forceinline__ float fp16c_decode(uint16_t x) {
    return __uint_as_float((((uint32_t)x & 0x8000u) << 16) | (((uint32_t)x & 0x7FFFu) << 12)) * 5.192296858534828e33f;  // 2^112
}

template <int FP> struct Codec;
template <> struct Codec<ION_FP32> {
    typedef float store_t;
    static __device__ __forceinline__ float dec(float v) { return v; }
    static __device__ __forceinline__ float enc(float v) { return v; }
};
template <> struct Codec<ION_FP16S> {
    typedef uint16_t store_t;
    static __device__ __forceinline__ float dec(uint16_t v) { return __half2float(__ushort_as_half(v)) * 3.0517578E-5f; }
    static __device__ __forceinline__ uint16_t enc(float v) { return __half_as_ushort(__float2half_rn(v * 32768.0f)); }
};
template <> struct Codec<ION_FP16C> {
    typedef uint16_t store_t;
    static __device__ __forceinline__ float dec(uint16_t v) { return fp16c_decode(v); }
    static __device__ __forceinline__ uint16_t enc(float v) { return fp16c_encode(v); }
};

// ------------------------------------------------------------------------------------------------------
// cell coordinates and periodic neighbours (sim.cl:134-148,248-302); one thread per cell, x fastest
// ------------------------------------------------------------------------------------------------------
struct Cell {
    uint32_t x, y, z;
    uint32_t n;                       // x+(y+z*ny)*nx
    uint32_t x0, xp, xm, y0, yp, ym;  // sim.cl:250-255
    uint32_t z0, zp, zm;              // sim.cl:256-258 (fits 32 bit: N <= 2^32)
};

__device__ __forceinline__ Cell make_cell(const KArgs& a, uint32_t x, uint32_t y, uint32_t z) {
    Cell c;
    c.x = x; c.y = y; c.z = z;
    const uint32_t nx = a.nx, ny = a.ny, nz = a.nz;
    c.x0 = x;
    c.xp = (x + 1u == nx) ? 0u : x + 1u;
    c.xm = (x == 0u) ? nx - 1u : x - 1u;
    c.y0 = y * nx;
    c.yp = ((y + 1u == ny) ? 0u : y + 1u) * nx;
    c.ym = ((y == 0u) ? ny - 1u : y - 1u) * nx;
    const uint32_t nxy = nx * ny;
    c.z0 = z * nxy;
    c.zp = ((z + 1u == nz) ? 0u : z + 1u) * nxy;
    c.zm = ((z == 0u) ? nz - 1u : z - 1u) * nxy;
    c.n = c.x0 + c.y0 + c.z0;
    return c;
}
__device__ __forceinline__ Cell make_cell_n(const KArgs& a, uint32_t n) {  // coordinates(n), sim.cl:134-137
    const uint32_t nxy = a.nx * a.ny;
    const uint32_t t = n % nxy;
    return make_cell(a, t % a.nx, t / a.nx, n / nxy);
}
__device__ __forceinline__ bool is_halo(const KArgs& a, uint32_t x, uint32_t y, uint32_t z) {  // sim.cl:145-148
    return ((a.dx > 1u) & (x == 0u || x >= a.nx - 1u)) || ((a.dy > 1u) & (y == 0u || y >= a.ny - 1u)) ||
           ((a.dz > 1u) & (z == 0u || z >= a.nz - 1u));
}
// j[i] of sim.cl:260-302.  D2Q9 ignores z exactly like the reference (sim.cl:265-268).
template <int VS> __device__ __forceinline__ uint32_t neighbor(const Cell& c, int i) {
    const int cx = cvel<VS>(0, i), cy = cvel<VS>(1, i), cz = cvel<VS>(2, i);
    uint32_t j = (cx > 0 ? c.xp : cx < 0 ? c.xm : c.x0) + (cy > 0 ? c.yp : cy < 0 ? c.ym : c.y0);
    if (VS != ION_D2Q9) j += (cz > 0 ? c.zp : cz < 0 ? c.zm : c.z0);
    return i == 0 ? c.n : j;
}
// D3Q7 sub-lattice neighbours for the charge / temperature DDFs (neighbors_a, sim.cl:382-389)
__device__ __forceinline__ uint32_t neighbor7(const Cell& c, int i) {
    switch (i) {
        case 1: return c.xp + c.y0 + c.z0;
        case 2: return c.xm + c.y0 + c.z0;
        case 3: return c.x0 + c.yp + c.z0;
        case 4: return c.x0 + c.ym + c.z0;
        case 5: return c.x0 + c.y0 + c.zp;
        case 6: return c.x0 + c.y0 + c.zm;
        default: return c.n;
    }
}

// ------------------------------------------------------------------------------------------------------
// Esoteric-Pull in-place streaming (load_f/store_f sim.cl:234-247, load_a/store_a sim.cl:397-410).
// SoA layout i*N+n (index_f, sim.cl:152-154).  Every (cell, slot) address is read and written by exactly one
// thread per step, so the update is race-free without a second DDF copy.
// ------------------------------------------------------------------------------------------------------
// DDF accesses.  ION_DDF_HINT selects the cache operator (build-time experiment switch):
// 0 = default (ld.global / st.global), 1 = streaming (ld.global.cs / st.global.cs: evict-first in L1/L2)
#ifndef ION_DDF_HINT
#define ION_DDF_HINT 0
#endif
// ION_SPEC_LOADS = 1: DDF / field loads of stream_collide are `asm volatile`, which pins them in program order BEFORE the
// branch on the cell's flag byte.  With plain C++ loads nvcc sinks all of them below that branch (their values are dead on
// the solid path), so a warp first waits one full HBM round trip for one byte per cell with nothing else in flight.
#ifndef ION_SPEC_LOADS
#define ION_SPEC_LOADS 1
#endif
#if ION_SPEC_LOADS == 2
#define ION_LD_PIN "ld.relaxed.gpu.global"
#elif ION_SPEC_LOADS == 3
#define ION_LD_PIN "ld.global"
#else
#define ION_LD_PIN "ld.volatile.global"
#endif
__device__ __forceinline__ float ld_f32_pinned(const float* p) {
#if ION_SPEC_LOADS
    float v;
    asm volatile(ION_LD_PIN ".f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ uint8_t ld_u8_pinned(const uint8_t* p) {
#if ION_SPEC_LOADS
    uint16_t v;
    asm volatile(ION_LD_PIN ".u8 %0, [%1];" : "=h"(v) : "l"(p));
    return (uint8_t)v;
#else
    return *p;
#endif
}
template <typename S> __device__ __forceinline__ S ddf_ld(const S* p) {
#if ION_DDF_HINT == 1
    return __ldcs(p);
#else
    return *p;
#endif
}
#if ION_SPEC_LOADS && ION_DDF_HINT == 0
template <> __device__ __forceinline__ float ddf_ld<float>(const float* p) { return ld_f32_pinned(p); }
template <> __device__ __forceinline__ uint16_t ddf_ld<uint16_t>(const uint16_t* p) {
    uint16_t v;
    asm volatile(ION_LD_PIN ".u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
#endif
template <typename S> __device__ __forceinline__ void ddf_st(S* p, S v) {
#if ION_DDF_HINT == 1
    __stcs(p, v);
#else
    *p = v;
#endif
}
template <int FP, int QQ, typename NB>
__device__ __forceinline__ void ep_load(float* f, const void* buf, uint64_t N, uint32_t n, uint64_t todd, NB nb) {
    typedef typename Codec<FP>::store_t S;
    const S* p = reinterpret_cast<const S*>(buf);
    f[0] = Codec<FP>::dec(ddf_ld(p + n));
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        f[i] = Codec<FP>::dec(ddf_ld(p + ((uint64_t)(i + 1 - (int)todd) * N + n)));          // t odd ? i : i+1
        f[i + 1] = Codec<FP>::dec(ddf_ld(p + ((uint64_t)(i + (int)todd) * N + nb(i))));      // t odd ? i+1 : i
    }
}
template <int FP, int QQ, typename NB>
__device__ __forceinline__ void ep_store(const float* f, void* buf, uint64_t N, uint32_t n, uint64_t todd, NB nb) {
    typedef typename Codec<FP>::store_t S;
    S* p = reinterpret_cast<S*>(buf);
    ddf_st(p + n, Codec<FP>::enc(f[0]));
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        ddf_st(p + ((uint64_t)(i + (int)todd) * N + nb(i)), Codec<FP>::enc(f[i]));          // t odd ? i+1 : i
        ddf_st(p + ((uint64_t)(i + 1 - (int)todd) * N + n), Codec<FP>::enc(f[i + 1]));      // t odd ? i : i+1
    }
}

// Compile-time parity variants (ODD = t & 1) used by stream_collide: every slot index is a constant, and each address is
// formed as (64-bit pointer of the cell or of its neighbour) + (warp-uniform slot offset), i.e. two integer instructions per
// access.  With a run-time parity the slot offsets (i + 1 - todd) * N are 64-bit run-time products and the address arithmetic
// was ~40 % of the kernel's instructions (SASS of the FP32 MHD kernel: 620 integer/uniform instructions for 90 accesses).
template <int FP, int QQ, bool ODD, typename NB>
__device__ __forceinline__ void ep_load_p(float* f, const void* buf, uint64_t N, uint32_t n, NB nb) {
    typedef typename Codec<FP>::store_t S;
    const S* p = reinterpret_cast<const S*>(buf);
    const S* pn = p + n;
    f[0] = Codec<FP>::dec(ddf_ld(pn));
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        f[i] = Codec<FP>::dec(ddf_ld(pn + (uint64_t)(ODD ? i : i + 1) * N));               // t odd ? i : i+1
        f[i + 1] = Codec<FP>::dec(ddf_ld((p + nb(i)) + (uint64_t)(ODD ? i + 1 : i) * N));   // t odd ? i+1 : i
    }
}
template <int FP, int QQ, bool ODD, typename NB>
__device__ __forceinline__ void ep_store_p(const float* f, void* buf, uint64_t N, uint32_t n, NB nb) {
    typedef typename Codec<FP>::store_t S;
    S* p = reinterpret_cast<S*>(buf);
    S* pn = p + n;
    ddf_st(pn, Codec<FP>::enc(f[0]));
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        ddf_st((p + nb(i)) + (uint64_t)(ODD ? i + 1 : i) * N, Codec<FP>::enc(f[i]));        // t odd ? i+1 : i
        ddf_st(pn + (uint64_t)(ODD ? i : i + 1) * N, Codec<FP>::enc(f[i + 1]));             // t odd ? i : i+1
    }
}

// ------------------------------------------------------------------------------------------------------
// moments, equilibrium, forcing
// ------------------------------------------------------------------------------------------------------
// calculate_rho_u, sim.cl:208-233: rho = (f0+f1+...)+1; momentum sums run over direction pairs in index
// order, "+ then -" inside each pair, strictly left to right.
template <int VS> __device__ __forceinline__ void rho_u(const float* f, float& rho, float& ux, float& uy, float& uz) {
    constexpr int QQ = VSet<VS>::Q;
    float r = f[0];
#pragma unroll
    for (int i = 1; i < QQ; i++) r += f[i];
    r += 1.0f;
    float m[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        bool first = true;
#pragma unroll
        for (int i = 1; i < QQ; i += 2) {
            const int c = cvel<VS>(ax, i);
            if (c != 0) {
                const float plus = c > 0 ? f[i] : f[i + 1], minus = c > 0 ? f[i + 1] : f[i];
                m[ax] = first ? plus : m[ax] + plus;
                m[ax] = m[ax] - minus;
                first = false;
            }
        }
    }
    rho = r;
    ux = m[0] / r;
    uy = m[1] / r;
    uz = VS == ION_D2Q9 ? 0.0f / r : m[2] / r;
}

// calculate_f_eq, sim.cl:155-207 (DDF-shifted equilibrium)
template <int VS> __device__ __forceinline__ void f_eq(float rho, float ux, float uy, float uz, float* feq) {
    constexpr int QQ = VSet<VS>::Q;
    const float c3 = -3.0f * (sq(ux) + sq(uy) + sq(uz)), rhom1 = rho - 1.0f;
    ux *= 3.0f;
    uy *= 3.0f;
    uz *= 3.0f;
    feq[0] = wclass<VS>(0) * fmaf(rho, 0.5f * c3, rhom1);
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        const float wi = wdir<VS>(i);
        const float rhow = wi * rho, rhom1w = wi * rhom1;
        const float ui = cdot<VS>(i, ux, uy, uz);
        const float t = fmaf(ui, ui, c3);
        feq[i] = fmaf(rhow, fmaf(0.5f, t, ui), rhom1w);
        feq[i + 1] = fmaf(rhow, fmaf(0.5f, t, -ui), rhom1w);
    }
}

// calculate_forcing_terms (Guo forcing), sim.cl:367-377
template <int VS>
__device__ __forceinline__ void forcing_terms(float ux, float uy, float uz, float fx, float fy, float fz, float* Fin) {
    constexpr int QQ = VSet<VS>::Q;
    const float uF = VS == ION_D2Q9 ? -0.33333334f * fmaf(ux, fx, uy * fy) : -0.33333334f * fmaf(ux, fx, fmaf(uy, fy, uz * fz));
    Fin[0] = 9.0f * wclass<VS>(0) * uF;
#pragma unroll
    for (int i = 1; i < QQ; i++) {
        Fin[i] = (9.0f * wdir<VS>(i)) * fmaf(cdot<VS>(i, fx, fy, fz), cdot<VS>(i, ux, uy, uz) + 0.33333334f, uF);
    }
}

// ------------------------------------------------------------------------------------------------------
// Two-species packed arithmetic for the MHD kernel.  The neutral gas (fi) and the electron gas (ei) go through the SAME
// moment / equilibrium / Guo-forcing / SRT-relaxation formulas on the same lattice (sim.cl:208-233,155-207,367-377,649-655
// and 712-723), so they are evaluated together in the two lanes of sm_100's packed FP32 instructions (add/mul/fma.rn.f32x2):
// lane .x = gas, lane .y = electrons.  Each lane is an independent IEEE round-to-nearest operation, so every value is
// bit-identical to the scalar sequence; the instruction count of these blocks halves.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int VS> __device__ __forceinline__ float2 cdot2(int i, float2 ax, float2 ay, float2 az) {  // cdot, both lanes
    const int cx = cvel<VS>(0, i), cy = cvel<VS>(1, i), cz = cvel<VS>(2, i);
    float2 s = make_float2(0.0f, 0.0f);
    bool first = true;
    if (cx != 0) { s = cx > 0 ? ax : neg2(ax); first = false; }
    if (cy != 0) { const float2 t = cy > 0 ? ay : neg2(ay); s = first ? t : add2(s, t); first = false; }
    if (cz != 0) { const float2 t = cz > 0 ? az : neg2(az); s = first ? t : add2(s, t); first = false; }
    return s;
}

// calculate_rho_u for both species: density and MOMENTUM sums (the three divisions stay scalar at the caller)
template <int VS> __device__ __forceinline__ void rho_m2(const float* f, const float* e, float2& rho, float2 (&m)[3]) {
    constexpr int QQ = VSet<VS>::Q;
    float2 r = make_float2(f[0], e[0]);
#pragma unroll
    for (int i = 1; i < QQ; i++) r = add2(r, make_float2(f[i], e[i]));
    rho = add2(r, splat2(1.0f));
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        m[ax] = make_float2(0.0f, 0.0f);
        bool first = true;
#pragma unroll
        for (int i = 1; i < QQ; i += 2) {
            const int c = cvel<VS>(ax, i);
            if (c != 0) {
                const float2 plus = c > 0 ? make_float2(f[i], e[i]) : make_float2(f[i + 1], e[i + 1]);
                const float2 minus = c > 0 ? make_float2(f[i + 1], e[i + 1]) : make_float2(f[i], e[i]);
                m[ax] = first ? plus : add2(m[ax], plus);
                m[ax] = add2(m[ax], neg2(minus));  // a - b == a + (-b) exactly
                first = false;
            }
        }
    }
}

// TRT relaxation of one direction pair (i, i+1) of the gas, sim.cl:725-755 restricted to the pair (the scheme only couples
// a direction with its opposite).  vf: Guo terms are split into symmetric / antisymmetric parts first (sim.cl:731-741).
__device__ __forceinline__ void trt_pair(float& fa, float& fb, float qa, float qb, float Fa, float Fb, float wp, float wm, bool vf) {
    if (vf) {
        const float c_taup = fmaf(wp, -0.25f, 0.5f), c_taum = fmaf(wm, -0.25f, 0.5f);
        const float Fa2 = fmaf(c_taup, Fa + Fb, c_taum * (Fa - Fb)), Fb2 = fmaf(c_taup, Fb + Fa, c_taum * (Fb - Fa));
        Fa = Fa2;
        Fb = Fb2;
    }
    const float na = fmaf(0.5f * wp, qa - fa + qb - fb, fmaf(0.5f * wm, qa - qb - fa + fb, fa + Fa));
    const float nb = fmaf(0.5f * wp, qb - fb + qa - fa, fmaf(0.5f * wm, qb - qa - fb + fa, fb + Fb));
    fa = na;
    fb = nb;
}

// Equilibrium + Guo forcing + relaxation of both species, one direction pair at a time so that neither feq[] nor Fin[] is
// ever materialised (the scalar code keeps 2*Q of them live).  rho/u/F: lane .x gas, lane .y electrons; u is the
// force-corrected, clamped velocity.  Gas: SRT or TRT (template), forcing only if vf; electrons: always SRT with forcing
// (sim.cl:641-656).  is_e: TYPE_E cell under EQUILIBRIUM_BOUNDARIES -> both populations are set to their equilibrium.
template <int VS, bool TRT>
__device__ __forceinline__ void collide_two_species(float* f, float* e, float2 rho, float2 ux, float2 uy, float2 uz, float2 Fx, float2 Fy, float2 Fz,
                                                    float w, bool vf, bool is_e) {
    constexpr int QQ = VSet<VS>::Q;
    const float c_tau = fmaf(w, -0.5f, 1.0f);  // sim.cl:519
    const float wm = TRT ? 1.0f / (0.1875f / (1.0f / w - 0.5f) + 0.5f) : 0.0f;
    const float2 W = splat2(w), W1 = splat2(1.0f - w), CT = splat2(c_tau), HALF = splat2(0.5f), THIRD = splat2(0.33333334f);
    // forcing_terms, sim.cl:367-377
    const float2 uF = VS == ION_D2Q9 ? mul2(splat2(-0.33333334f), fma2(ux, Fx, mul2(uy, Fy)))
                                     : mul2(splat2(-0.33333334f), fma2(ux, Fx, fma2(uy, Fy, mul2(uz, Fz))));
    // calculate_f_eq, sim.cl:155-207
    // c3 stays scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (seen in SASS; the
    // scalar forms are left alone), which would change the rounding of this sum of squares.  Nowhere else in this function
    // does a packed product feed a packed ADD (products only enter FMAs, as multiplicand or addend, which cannot be re-fused).
    const float2 c3 = make_float2(-3.0f * (sq(ux.x) + sq(uy.x) + sq(uz.x)), -3.0f * (sq(ux.y) + sq(uy.y) + sq(uz.y)));
    const float2 rhom1 = add2(rho, splat2(-1.0f));
    const float2 u3x = mul2(ux, splat2(3.0f)), u3y = mul2(uy, splat2(3.0f)), u3z = mul2(uz, splat2(3.0f));
    auto relax = [&](float2 old, float2 feq, float2 Fin) {  // SRT, sim.cl:720 / 653
        return fma2(W1, old, fma2(W, feq, mul2(Fin, CT)));
    };
    {
        const float2 feq0 = mul2(splat2(wclass<VS>(0)), fma2(rho, mul2(HALF, c3), rhom1));
        float2 Fin0 = mul2(splat2(9.0f * wclass<VS>(0)), uF);
        if (!vf) Fin0.x = 0.0f;
        const float2 o = relax(make_float2(f[0], e[0]), feq0, Fin0);
        float g = o.x;
        if (TRT) {  // direction 0 is its own opposite (fhb[0] = fhn[0], feb[0] = feq[0], Fib[0] = Fin[0])
            float F0 = Fin0.x;
            if (vf) F0 = fmaf(fmaf(w, -0.25f, 0.5f), F0 + F0, fmaf(wm, -0.25f, 0.5f) * (F0 - F0));
            g = fmaf(0.5f * w, feq0.x - f[0] + feq0.x - f[0], fmaf(0.5f * wm, feq0.x - feq0.x - f[0] + f[0], f[0] + F0));
        }
        f[0] = is_e ? feq0.x : g;
        e[0] = is_e ? feq0.y : o.y;
    }
#pragma unroll
    for (int i = 1; i < QQ; i += 2) {
        const float wi = wdir<VS>(i);
        const float2 cf = cdot2<VS>(i, Fx, Fy, Fz), cu = cdot2<VS>(i, ux, uy, uz);
        float2 Fa = mul2(splat2(9.0f * wi), fma2(cf, add2(cu, THIRD), uF));
        float2 Fb = mul2(splat2(9.0f * wi), fma2(neg2(cf), add2(neg2(cu), THIRD), uF));
        if (!vf) { Fa.x = 0.0f; Fb.x = 0.0f; }
        const float2 rhow = mul2(splat2(wi), rho), rhom1w = mul2(splat2(wi), rhom1);
        const float2 ui = cdot2<VS>(i, u3x, u3y, u3z);
        const float2 t = fma2(ui, ui, c3);
        const float2 qa = fma2(rhow, fma2(HALF, t, ui), rhom1w);
        const float2 qb = fma2(rhow, fma2(HALF, t, neg2(ui)), rhom1w);
        const float2 oa = relax(make_float2(f[i], e[i]), qa, Fa);
        const float2 ob = relax(make_float2(f[i + 1], e[i + 1]), qb, Fb);
        float ga = oa.x, gb = ob.x;
        if (TRT) {
            ga = f[i];
            gb = f[i + 1];
            trt_pair(ga, gb, qa.x, qb.x, Fa.x, Fb.x, w, wm, vf);
        }
        f[i] = is_e ? qa.x : ga;
        f[i + 1] = is_e ? qb.x : gb;
        e[i] = is_e ? qa.y : oa.y;
        e[i + 1] = is_e ? qb.y : ob.y;
    }
}

// calculate_a_eq (D3Q7 advected scalar), sim.cl:390-396
__device__ __forceinline__ void a_eq(float Q, float ux, float uy, float uz, float* qeq) {
    const float wsT4 = 0.5f * Q, wsTm1 = 0.125f * (Q - 1.0f);
    qeq[0] = fmaf(0.25f, Q, -0.25f);
    qeq[1] = fmaf(wsT4, ux, wsTm1); qeq[2] = fmaf(wsT4, -ux, wsTm1);
    qeq[3] = fmaf(wsT4, uy, wsTm1); qeq[4] = fmaf(wsT4, -uy, wsTm1);
    qeq[5] = fmaf(wsT4, uz, wsTm1); qeq[6] = fmaf(wsT4, -uz, wsTm1);
}

// ------------------------------------------------------------------------------------------------------
// LOD pyramid helpers (sim.cl:425-447)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lod_index(const KArgs& a, uint32_t x, uint32_t y, uint32_t z, uint32_t d) {
    const uint32_t nd = 1u << d;
    return x / (a.nx / nd) + (y / (a.ny / nd) + z / (a.nz / nd) * nd) * nd;
}
__device__ __forceinline__ float lod_s(const KArgs& a, uint32_t d) {
    const uint32_t nd = 1u << d;
    return (float)((a.nx / nd) * (a.ny / nd) * (a.nz / nd));
}

}  // namespace ion
