// eb_fft_core.cuh -- update_e_b_dynamic as a polyphase FFT convolution (the bodies of the kernels in eb_fft.cu).
//
// Behaviour: /root/reference/src/kernels/sim_kernels.cl ("sim.cl") update_e_b_dynamic :897-993, own-LOD loop :940-955.
//
// Why: the reference sums, for every cell, over the 8^D sources of its own finest LOD level (4096 at the default depth 4) --
// 9 multiply-adds per (cell, source) pair even after the Toeplitz tiling of fields.cu, 1.06e12 FMA per step at 256^3, 95 % of
// the whole time step.  But the sources sit on a regular ND^3 grid (ND = 2^D) and a cell is (block b, in-block offset o):
//     r = cell - centre = (b - c) * ds + (o - ds/2)            (sim.cl:945-946, lod_coordinates :440-447)
// so for a FIXED offset o the ND^3 cells that share it see the sources through a kernel that depends on b - c only:
//     E_o(b) = sum_c K_o(b - c) q(c),    B_o(b) = sum_c w(c) x K_o(b - c),    K_o(d) = r / |r|^3,  w = q v
// -- a 3-D linear convolution of ND^3 sources with a (2 ND)^3 kernel, one per offset ("polyphase").  With M = 2 ND points per
// axis the circular convolution has no wrap-around, and the convolution theorem turns 8^D terms per cell into
// O(log M) work: ~4.5 MFLOP per offset instead of 300 MFLOP.
//   * K^_o = FFT3(K_o) (Hermitian half spectrum, H = ND + 1 planes along x) is static per geometry: computed once
//     (eb_khat_*), kept in HBM, 102 B per cell.  The normalisation 1/M^3 and the factor 2 of the Hermitian fold are baked in.
//   * s^_j = FFT3(q, q vx, q vy, q vz) once per step (eb_src_*), 4 x H x M x M complex, L2 resident.
//   * per offset (eb_main_*): for each kx plane  E^_c = K^_c s^_0,  B^ = w^ x K^  (9 complex products per frequency),
//     2-D inverse FFT over (ky, kz) in shared memory, pruned to the ND x ND outputs that exist, and the last axis as a
//     direct Hermitian DFT accumulation  E(x) += Re e(kx) cos(2 pi kx x / M) - Im e(kx) sin(2 pi kx x / M)  into registers
//     (17 planes x 16 outputs: cheap, and it removes the third transpose and the 192 KB buffer it would need).
// Quirks kept: the window [NUM_LOD_OWN - 8^D, NUM_LOD_OWN) with positions taken from the table index (Q5), the self-skip
// d == lod_index(cell) (sim.cl:944) as K_o(0) = 0, lod_index overflow planes of halo-inclusive slabs (Q7) as extra
// z windows.  Summation order differs from the reference (as does every parallel sum): E/B agree to ~2e-7 relative L2
// (tests/test_gpu_parity.py), the deterministic mode keeps the reference's order and arithmetic.
//
// Every phase is a plain function of (thread id, block id): eb_fft.cu calls them between __syncthreads(), and
// tests/tools/eb_fft_emul.cpp runs the same functions thread by thread on the CPU against a direct double-precision sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define ION_HD __host__ __device__ __forceinline__
#else
#define ION_HD inline
#endif

namespace ion {
namespace ebfft {

struct Geom {
    uint32_t nx, ny, nz;
    uint64_t N;
    uint32_t dsx, dsy, dsz;  // cells per LOD block
    uint32_t lo, n_lod_own;  // source window [lo, n_lod_own) of sim.cl:943
    uint32_t cz0;            // z row of the first window entry: lo / ND^2
    uint32_t dx, dy, dz;     // domain grid (halo test, sim.cl:145-148)
    float ke, kmu;
};
// A set of sources that sits on the block lattice and therefore convolves: r = ds * (block difference) + (o + off).
//   kind 0: the own finest level, window [lo, n_lod_own) of sim.cl:943, positions from the table index (quirk Q5): centre
//           (c + 1/2) ds, i.e. off = -ds/2; z rows start at cz0, i.e. zadd = -cz0; the cell's own block is skipped (sim.cl:944).
//   kind 1: the level D-1 pyramid of the slab below (sim.cl:957-983).  Its blocks are two own blocks wide, centre (2i + 1) ds:
//           the source lands on the ODD positions of the ND-grid with off = 0 in x, y; in z the centre is shifted by the slab
//           height INCLUDING halos (quirk Q8), nz = A dsz + B: zadd = A, off_z = B.  No self-skip.
struct SourceSet {
    uint32_t kind;
    uint32_t entry0;  // kind 1: first QU_lod entry of the neighbour's pyramid level
    int32_t zadd;     // true z block difference = circular difference + ND * wz + zadd
    float offx, offy, offz;
};
// one polyphase problem: the cells (b*ds + o) with z blocks [ND*wz, ND*wz + ND)
// `kslot`: which entry of the kernel-spectrum buffer the task's products read (its own, or -- mirror != 0 -- the one of the task
// with x offset dsx - ox, see main_phase_product_mirror)
struct Task {
    uint16_t ox, oy, oz, wz;
    uint32_t kslot;
    uint32_t mirror;
};

template <int ND> struct Cfg {
    static constexpr int M = 2 * ND;      // FFT length per axis
    static constexpr int H = ND + 1;      // Hermitian half: kx = 0..ND
    static constexpr int ROW = M + 1;     // padded row of the shared-memory planes (bank-conflict-free strided access)
    static constexpr int P = ND == 16 ? 2 : 3;  // kx planes per iteration; the kernel holds two such groups (double buffer)
    static constexpr int NIT = (H + P - 1) / P; // iterations per task
    // threads of the main kernel.  ND = 16: 384 = P * 6 spectra * 32 columns, so the z pass is exactly one FFT per thread (12 warps,
    // one (plane, spectrum) each); the y pass uses half of them and the x accumulation the first 256 (one per (y, z) point)
    static constexpr int T = ND == 16 ? 384 : 256;
    static constexpr int XG = ND == 16 ? 1 : 4; // thread groups along x in the accumulation phase (ND^2 * XG threads take part)
    static constexpr int XPT = ND / XG;         // x outputs per thread (they live in registers: 6 * XPT accumulators)
    static constexpr int TACC = ND * ND * XG;   // threads that own accumulators
    static constexpr int PLANE = 6 * M * ROW;   // float2 per kx plane: E^x E^y E^z B^x B^y B^z
    static constexpr int SLOT = M * ROW;        // float2 per (spectrum, kx) slot; K^ and s^ use the same padded rows in HBM
    static constexpr size_t khat_per_task = (size_t)3 * H * SLOT;  // float2
    static constexpr size_t shat_count = (size_t)4 * H * SLOT;     // float2
    // work buffers [2][P][6 slots] + s^_0 staging [P slots] + x twiddle pairs [H][ND/2] float4
    static constexpr size_t main_smem = (size_t)2 * P * PLANE * sizeof(float2) + (size_t)P * SLOT * sizeof(float2) + (size_t)H * (ND / 2) * sizeof(float4);
    // compact spectrum of a source set that lives on the ODD grid positions only (kind 1): s^(k + ND e_a) = -s^(k) on every axis, so
    // ND x ND values per (component, kx) describe the whole M x M plane
    static constexpr int CSLOT = ND * ND;
    static constexpr size_t shatc_count = (size_t)4 * H * CSLOT;
    static constexpr size_t khat_smem = (size_t)H * M * ROW * sizeof(float2);
    static constexpr int KT = ((H * M + 31) / 32) * 32;  // threads of k_eb_khat: one (kx, line) FFT each in its y and z passes (544 / 160)
    static constexpr size_t src_smem = (size_t)M * ROW * sizeof(float2);
};

// cos(2 pi k / 32).  A chain of selects, not a table: with k a compile-time constant after unrolling it folds to an immediate
// (a local constexpr array is materialised on the stack by nvcc and read back with LDL).
ION_HD constexpr float tw_cos32_q(int k) {  // k = 0..8
    return k == 0 ? 1.0f
         : k == 1 ? 9.807852804e-01f
         : k == 2 ? 9.238795325e-01f
         : k == 3 ? 8.314696123e-01f
         : k == 4 ? 7.071067812e-01f
         : k == 5 ? 5.555702330e-01f
         : k == 6 ? 3.826834324e-01f
         : k == 7 ? 1.950903220e-01f
                  : 0.0f;
}
ION_HD constexpr float tw_cos32(int k) {
    k &= 31;
    if (k > 16) k = 32 - k;
    return k > 8 ? -tw_cos32_q(16 - k) : tw_cos32_q(k);
}
ION_HD constexpr float tw_sin32(int k) { return tw_cos32(k + 24); }  // sin(t) = cos(t - pi/2)

ION_HD constexpr int bitrev(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}
ION_HD constexpr int ilog2(int m) { return m <= 1 ? 0 : 1 + ilog2(m >> 1); }

// In-register radix-2 decimation-in-time FFT of M = 16 or 32 complex points.  INV = false: e^{-2 pi i k n / M};
// INV = true: e^{+...}, unnormalised.  Everything is unrolled, so all indices and twiddles are compile-time constants;
// outputs the caller never reads (pruned transforms) and inputs that are literal zeros fold away.
// complex add / subtract as ONE packed FP32 instruction on the device (add.f32x2 / fma.f32x2 with -1): each lane is an IEEE
// operation, so the results equal the scalar form bit for bit; the butterflies' add/sub are 70 % of an FFT's instructions
#ifndef ION_FFT_PACKED
#define ION_FFT_PACKED 1
#endif
ION_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && ION_FFT_PACKED
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
ION_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && ION_FFT_PACKED
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
template <int M, int LEN, bool INV> struct FftStage {  // butterflies of span LEN, after all shorter spans
    static ION_HD void run(float2 (&a)[M]) {
        FftStage<M, LEN / 2, INV>::run(a);
        constexpr int half = LEN / 2, step = 32 / LEN;
#pragma unroll
        for (int i = 0; i < M; i += LEN) {
#pragma unroll
            for (int k = 0; k < half; k++) {
                const int ti = k * step;  // twiddle angle in units of 2 pi / 32, 0..15
                const float2 u = a[i + k], v = a[i + k + half];
                float2 w;
                if (ti == 0) {
                    w = v;
                } else if (ti == 8) {
                    w = INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
                } else {
                    const float c = tw_cos32(ti), sn = INV ? tw_sin32(ti) : -tw_sin32(ti);
                    w = make_float2(fmaf(v.x, c, -(v.y * sn)), fmaf(v.x, sn, v.y * c));
                }
                a[i + k] = cadd(u, w);
                a[i + k + half] = csub(u, w);
            }
        }
    }
};
template <int M, bool INV> struct FftStage<M, 1, INV> {
    static ION_HD void run(float2 (&)[M]) {}
};
template <int M, bool INV> ION_HD void fft_reg(float2 (&a)[M]) {
    constexpr int BITS = ilog2(M);
#pragma unroll
    for (int i = 0; i < M; i++) {
        const int j = bitrev(i, BITS);
        if (i < j) {
            const float2 t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
    FftStage<M, M, INV>::run(a);
}

ION_HD float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)); }
// a*b - c*d
ION_HD float2 cmul_sub(float2 a, float2 b, float2 c, float2 d) {
    const float2 p = cmul(a, b), q = cmul(c, d);
    return make_float2(p.x - q.x, p.y - q.y);
}

// ------------------------------------------------------------------------------------------------------
// K^ of one (task, component): grid (ntasks, 3), Cfg::T threads, shared S[H][M][ROW]
// ------------------------------------------------------------------------------------------------------
// signed block difference of circular index m
template <int ND> ION_HD int circ_diff(int m) { return m < ND ? m : m - 2 * ND; }

template <int ND> ION_HD float khat_value(const Geom& g, const SourceSet& ss, const Task t, int comp, int ddx, int ddy, int ddz, float ry, float rz) {
    // r = (float)cell - ((float)c * ds + 0.5 ds): integers and half-integers below 2^12, exact in FP32 (sim.cl:945, :440-447);
    // r^2 < 2^24 is exact too up to 1024-cell blocks.  1 / (r^2 sqrt(r^2)) in IEEE FP32 (three roundings, <= 1.5 ulp): below the
    // 2e-7 the FP32 transforms carry, and several times cheaper than FP64 -- it matters in streamed mode, where K^ is rebuilt every step
    const float rx = (float)ddx * (float)g.dsx + ((float)t.ox + ss.offx);
    const float r2 = rx * rx + (ry * ry + rz * rz);
    if ((ss.kind == 0u && ddx == 0 && ddy == 0 && ddz == 0) || !(r2 > 0.0f)) return 0.0f;  // the cell's own block contributes nothing (sim.cl:944)
    const float inv = 1.0f / (r2 * sqrtf(r2));
    return (comp == 0 ? rx : comp == 1 ? ry : rz) * inv;
}
// x transform of the real kernel, two y lines per complex FFT: z = K(my) + i K(my+1), K^(my)[k] = (z[k] + conj z[M-k]) / 2,
// K^(my+1)[k] = -i (z[k] - conj z[M-k]) / 2; the factor 1/2 is applied with the normalisation in khat_phase_z (a power of two: exact)
template <int ND> ION_HD void khat_phase_x(int tid, int nthreads, const Geom& g, const SourceSet& ss, const Task t, int comp, float2* S) {
    typedef Cfg<ND> C;
    for (int pair = tid; pair < C::M * C::M / 2; pair += nthreads) {
        const int mz = pair / (C::M / 2), my = 2 * (pair % (C::M / 2));
        const int ddy0 = circ_diff<ND>(my), ddy1 = circ_diff<ND>(my + 1);
        const int ddz = circ_diff<ND>(mz) + ND * (int)t.wz + ss.zadd;  // true block difference along z
        const float ry0 = (float)ddy0 * (float)g.dsy + ((float)t.oy + ss.offy);
        const float ry1 = (float)ddy1 * (float)g.dsy + ((float)t.oy + ss.offy);
        const float rz = (float)ddz * (float)g.dsz + ((float)t.oz + ss.offz);
        float2 a[C::M];
#pragma unroll
        for (int mx = 0; mx < C::M; mx++) {
            const int ddx = circ_diff<ND>(mx);
            a[mx] = make_float2(khat_value<ND>(g, ss, t, comp, ddx, ddy0, ddz, ry0, rz), khat_value<ND>(g, ss, t, comp, ddx, ddy1, ddz, ry1, rz));
        }
        fft_reg<C::M, false>(a);
        float2* row = S + (size_t)mz * C::ROW + my;
#pragma unroll
        for (int kx = 0; kx < C::H; kx++) {
            const float2 z = a[kx], m = a[(C::M - kx) % C::M];
            row[(size_t)kx * C::M * C::ROW] = make_float2(z.x + m.x, z.y - m.y);
            row[(size_t)kx * C::M * C::ROW + 1] = make_float2(z.y + m.y, m.x - z.x);
        }
    }
}
template <int ND> ION_HD void khat_phase_y(int tid, int nthreads, float2* S) {
    typedef Cfg<ND> C;
    for (int line = tid; line < C::H * C::M; line += nthreads) {
        float2* row = S + (size_t)line * C::ROW;  // line = kx * M + mz
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = row[i];
        fft_reg<C::M, false>(a);
#pragma unroll
        for (int i = 0; i < C::M; i++) row[i] = a[i];
    }
}
// layout of K^: [task][comp][kx][kz][ROW] -- rows padded exactly like the shared-memory planes, so that one slot is ONE
// contiguous bulk copy (cp.async.bulk) into the place where the products are formed
template <int ND> ION_HD void khat_phase_z(int tid, int nthreads, int task, int comp, const float2* S, float2* khat) {
    typedef Cfg<ND> C;
    for (int line = tid; line < C::H * C::M; line += nthreads) {
        const int kx = line / C::M, ky = line % C::M;
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = S[(kx * C::M + i) * C::ROW + ky];
        fft_reg<C::M, false>(a);
        // 1 / M^3 of the inverse transform; planes 1..ND-1 stand for themselves and their conjugates (Hermitian fold)
        const float scale = (kx == 0 || kx == ND ? 0.5f : 1.0f) / (float)(C::M * C::M * C::M);  // incl. the 1/2 of khat_phase_x
        float2* out = khat + ((size_t)task * 3 + comp) * C::H * C::SLOT + (size_t)kx * C::SLOT + ky;
#pragma unroll
        for (int kz = 0; kz < C::M; kz++) out[(size_t)kz * C::ROW] = make_float2(a[kz].x * scale, a[kz].y * scale);
        if (ky == 0) {  // the padding column is copied along with the rows: keep it defined
#pragma unroll
            for (int kz = 0; kz < C::M; kz++) out[(size_t)kz * C::ROW + C::M] = make_float2(0.0f, 0.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// s^ of the source table: grid (H, 4), any thread count, shared plane[M][ROW].  Component j: q, q vx, q vy, q vz
// ------------------------------------------------------------------------------------------------------
template <int ND> ION_HD float src_value(const Geom& g, const SourceSet& ss, const float* QU_lod, int j, int cx, int cy, int czp) {
    uint32_t d;
    if (ss.kind == 0u) {
        d = (uint32_t)cx + (uint32_t)ND * ((uint32_t)cy + (uint32_t)ND * ((uint32_t)czp + g.cz0));
        if (d < g.lo || d >= g.n_lod_own) return 0.0f;  // outside the window of sim.cl:943
    } else {  // entry (i, j, k) of the neighbour's level D-1 sits at (2i+1, 2j+1, 2k+1)
        constexpr int NF = ND / 2;
        if (!(cx & 1) || !(cy & 1) || !(czp & 1) || czp >= ND) return 0.0f;
        d = ss.entry0 + (uint32_t)(cx >> 1) + (uint32_t)NF * ((uint32_t)(cy >> 1) + (uint32_t)NF * (uint32_t)(czp >> 1));
    }
    const float q = QU_lod[4u * d];
    return j == 0 ? q : QU_lod[4u * d + (uint32_t)j] * q;
}
template <int ND> ION_HD void src_phase_x(int tid, int nthreads, const Geom& g, const SourceSet& ss, const float* QU_lod, int kx, int j, float2* plane) {
    typedef Cfg<ND> C;
    for (int idx = tid; idx < (ND + 1) * ND; idx += nthreads) {
        const int czp = idx / ND, cy = idx % ND;
        float re = 0.0f, im = 0.0f;
#pragma unroll
        for (int cx = 0; cx < ND; cx++) {
            const float v = src_value<ND>(g, ss, QU_lod, j, cx, cy, czp);
            const int ti = ((kx * cx) % C::M) * (32 / C::M);
            re = fmaf(v, tw_cos32(ti), re);
            im = fmaf(v, -tw_sin32(ti), im);
        }
        plane[czp * C::ROW + cy] = make_float2(re, im);
    }
}
template <int ND> ION_HD void src_phase_y(int tid, int nthreads, float2* plane) {
    typedef Cfg<ND> C;
    for (int czp = tid; czp <= ND; czp += nthreads) {
        float2* row = plane + czp * C::ROW;
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = i < ND ? row[i] : make_float2(0.0f, 0.0f);
        fft_reg<C::M, false>(a);
#pragma unroll
        for (int i = 0; i < C::M; i++) row[i] = a[i];
    }
}
// layout of s^: [j][kx][kz][ROW] (same padded rows as K^)
template <int ND> ION_HD void src_phase_z(int tid, int nthreads, int kx, int j, const float2* plane, float2* shat, float2* shatc) {
    typedef Cfg<ND> C;
    for (int ky = tid; ky < C::M; ky += nthreads) {
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = i <= ND ? plane[i * C::ROW + ky] : make_float2(0.0f, 0.0f);
        fft_reg<C::M, false>(a);
        float2* out = shat + ((size_t)j * C::H + kx) * C::SLOT + ky;
#pragma unroll
        for (int kz = 0; kz < C::M; kz++) out[(size_t)kz * C::ROW] = a[kz];
        if (shatc && ky < ND) {  // compact copy [j][kx][kz < ND][ky < ND] for the sets with the odd-position symmetry
            float2* oc = shatc + ((size_t)j * C::H + kx) * C::CSLOT + ky;
#pragma unroll
            for (int kz = 0; kz < ND; kz++) oc[(size_t)kz * ND] = a[kz];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// main kernel: grid ntasks, Cfg::T threads, shared W[P][6][M][ROW] float2 | tw[M] float2
// ------------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define ION_LDG2(p) __ldg(p)
#else
#define ION_LDG2(p) (*(p))
#endif

// phase 0 (CPU emulation of what the kernel does with cp.async.bulk): the K^ slots of planes kx0 .. kx0+np-1 land in the E
// slots (0..2) of the work buffer and the s^_1..3 slots (q vx, q vy, q vz) in the B slots (3..5)
template <int ND> inline void main_stage_host(const float2* khat_task, const float2* shat, int kx0, int np, float2* W, float2* S0) {
    typedef Cfg<ND> C;
    for (int p = 0; p < np; p++) {
        for (int i = 0; i < C::SLOT; i++) S0[(size_t)p * C::SLOT + i] = shat[((size_t)kx0 + p) * C::SLOT + i];
        for (int c = 0; c < 3; c++)
            for (int i = 0; i < C::SLOT; i++) {
                W[(size_t)p * C::PLANE + (size_t)c * C::SLOT + i] = khat_task[((size_t)c * C::H + kx0 + p) * C::SLOT + i];
                W[(size_t)p * C::PLANE + (size_t)(3 + c) * C::SLOT + i] = shat[((size_t)(1 + c) * C::H + kx0 + p) * C::SLOT + i];
            }
    }
}
// phase 1: frequency-domain products of planes kx0 .. kx0+np-1, IN PLACE: the point's K^ (E slots) and s^_1..3 (B slots) are
// read from where the bulk copies put them, s^_0 from its own staging slots S0, and the six products overwrite the same point of
// the six slots.
template <int ND> ION_HD void main_phase_product(int tid, const float2* S0, int np, float2* W) {
    typedef Cfg<ND> C;
    constexpr int MM = C::M * C::M;
#pragma unroll 2
    for (int idx = tid; idx < np * MM; idx += C::T) {
        const int p = idx / MM, f = idx % MM;
        const int kz = f / C::M, ky = f % C::M;
        const int o = kz * C::ROW + ky;
        float2* w = W + (size_t)p * C::PLANE + o;
        const float2 s0 = S0[(size_t)p * C::SLOT + o];
        const float2 k0 = w[0], k1 = w[C::SLOT], k2 = w[2 * C::SLOT];
        const float2 s1 = w[3 * C::SLOT], s2 = w[4 * C::SLOT], s3 = w[5 * C::SLOT];
        w[0] = cmul(k0, s0);
        w[C::SLOT] = cmul(k1, s0);
        w[2 * C::SLOT] = cmul(k2, s0);
        w[3 * C::SLOT] = cmul_sub(s2, k2, s3, k1);  // (w x K)_x = w_y K_z - w_z K_y   (sim.cl:935)
        w[4 * C::SLOT] = cmul_sub(s3, k0, s1, k2);
        w[5 * C::SLOT] = cmul_sub(s1, k1, s2, k0);
    }
}
// phase 1 with TWO source sets (a slab with a lower neighbour): K^ of the own level was staged in the E slots, K^' of the
// neighbour's level D-1 in the B slots; the own source spectrum comes from L2 (loads batched three points deep), the neighbour's
// from its compact staged form S2c[P][4][ND*ND].  The spectra of the two convolutions are ADDED here, so that one inverse transform
// serves both (the transform is linear) -- the second set costs products, not FFTs.
template <int ND> ION_HD void main_phase_product2(int tid, const float2* shat, const float2* S2c, int kx0, int np, float2* W) {
    typedef Cfg<ND> C;
    constexpr int MM = C::M * C::M;
    constexpr size_t CS = (size_t)C::H * C::SLOT;  // component stride of s^
    constexpr int B = 3;                           // points per batch: all loads of a batch are in flight before the first product
    for (int idx0 = tid; idx0 < np * MM; idx0 += B * C::T) {
        float2 sv[B][4];
#pragma unroll
        for (int u = 0; u < B; u++) {
            const int idx = idx0 + u * C::T;
            if (idx < np * MM) {
                const int p = idx / MM, f = idx % MM;
                const size_t gi = (size_t)(kx0 + p) * C::SLOT + (f / C::M) * C::ROW + f % C::M;
#pragma unroll
                for (int j = 0; j < 4; j++) sv[u][j] = ION_LDG2(shat + (size_t)j * CS + gi);
            }
        }
#pragma unroll
        for (int u = 0; u < B; u++) {
            const int idx = idx0 + u * C::T;
            if (idx >= np * MM) continue;
            const int p = idx / MM, f = idx % MM;
            const int kz = f / C::M, ky = f % C::M;
            float2* w = W + (size_t)p * C::PLANE + kz * C::ROW + ky;
            // s^' from its compact form: sign flips when kz or ky leaves [0, ND)
            const float sg = ((kz >= ND) != (ky >= ND)) ? -1.0f : 1.0f;
            const float2* c2 = S2c + (size_t)p * 4 * C::CSLOT + (kz & (ND - 1)) * ND + (ky & (ND - 1));
            float2 t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 v = c2[(size_t)j * C::CSLOT];
                t[j] = make_float2(sg * v.x, sg * v.y);
            }
            const float2 s0 = sv[u][0], s1 = sv[u][1], s2 = sv[u][2], s3 = sv[u][3];
            const float2 k0 = w[0], k1 = w[C::SLOT], k2 = w[2 * C::SLOT];
            const float2 l0 = w[3 * C::SLOT], l1 = w[4 * C::SLOT], l2 = w[5 * C::SLOT];
            const float2 E0 = cmul(k0, s0), E1 = cmul(k1, s0), E2 = cmul(k2, s0);
            const float2 B0 = cmul_sub(s2, k2, s3, k1);  // (w x K)_x = w_y K_z - w_z K_y   (sim.cl:935)
            const float2 B1 = cmul_sub(s3, k0, s1, k2);
            const float2 B2 = cmul_sub(s1, k1, s2, k0);
            const float2 e0 = cmul(l0, t[0]), e1 = cmul(l1, t[0]), e2 = cmul(l2, t[0]);
            const float2 b0 = cmul_sub(t[2], l2, t[3], l1), b1 = cmul_sub(t[3], l0, t[1], l2), b2 = cmul_sub(t[1], l1, t[2], l0);
            w[0] = make_float2(E0.x + e0.x, E0.y + e0.y);
            w[C::SLOT] = make_float2(E1.x + e1.x, E1.y + e1.y);
            w[2 * C::SLOT] = make_float2(E2.x + e2.x, E2.y + e2.y);
            w[3 * C::SLOT] = make_float2(B0.x + b0.x, B0.y + b0.y);
            w[4 * C::SLOT] = make_float2(B1.x + b1.x, B1.y + b1.y);
            w[5 * C::SLOT] = make_float2(B2.x + b2.x, B2.y + b2.y);
        }
    }
}
// ------------------------------------------------------------------------------------------------------
// Mirror symmetry of the kernel in the in-block offset along x.  A cell with offset ox' = dsx - ox sees the own-level sources at
// r'(d) = (-r_x(-d_x), r_y(d_y), r_z(d_z)) -- the positions of offset ox reflected -- and the neighbour's level (kind 1, centres on
// block edges) at r'_x(d_x) = -r_x(-d_x - 1).  With K real that is, in frequency space,
//     K^'_c(kx, ky, kz) = s_c conj(K^_c(kx, -ky, -kz))                        s_x = -1, s_y = s_z = +1,
// times e^{+2 pi i kx / M} for the neighbour's level: a task with 2 ox > dsx needs no kernel spectrum of its own, it reads its
// partner's slots at the reflected in-plane index.  (The circular index d_x = -ND, which no output uses, is the only kernel sample the
// two forms disagree on; the fields differ by rounding only.)  Because the products overwrite the operands in place, a thread forms
// the products of a point f = (kz, ky) and of its reflection g = (-kz, -ky) together: both are read before either is written.
// ------------------------------------------------------------------------------------------------------
ION_HD float2 conj_sign(float2 k, float sign) { return make_float2(sign * k.x, -(sign * k.y)); }
template <int ND> ION_HD void main_phase_product_mirror(int tid, const float2* S0, int np, float2* W) {
    typedef Cfg<ND> C;
    constexpr int MM = C::M * C::M;
    for (int idx = tid; idx < np * MM; idx += C::T) {
        const int p = idx / MM, f = idx % MM;
        const int kz = f / C::M, ky = f % C::M;
        const int gz = (C::M - kz) & (C::M - 1), gy = (C::M - ky) & (C::M - 1);
        if (gz * C::M + gy < f) continue;  // the pair belongs to the thread that holds the smaller index
        const int of = kz * C::ROW + ky, og = gz * C::ROW + gy;
        float2* w = W + (size_t)p * C::PLANE;
        const float2* s0p = S0 + (size_t)p * C::SLOT;
        float2 kf[3], kg[3], sf[4], sg[4];
        sf[0] = s0p[of]; sg[0] = s0p[og];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            kf[c] = w[(size_t)c * C::SLOT + of]; kg[c] = w[(size_t)c * C::SLOT + og];
            sf[1 + c] = w[(size_t)(3 + c) * C::SLOT + of]; sg[1 + c] = w[(size_t)(3 + c) * C::SLOT + og];
        }
        // the mirrored task's kernel at f is the partner's at g (and the other way round)
        const float2 mf[3] = {conj_sign(kg[0], -1.0f), conj_sign(kg[1], 1.0f), conj_sign(kg[2], 1.0f)};
        const float2 mg[3] = {conj_sign(kf[0], -1.0f), conj_sign(kf[1], 1.0f), conj_sign(kf[2], 1.0f)};
#pragma unroll
        for (int c = 0; c < 3; c++) w[(size_t)c * C::SLOT + of] = cmul(mf[c], sf[0]);
        w[3 * C::SLOT + of] = cmul_sub(sf[2], mf[2], sf[3], mf[1]);  // (w x K)_x = w_y K_z - w_z K_y   (sim.cl:935)
        w[4 * C::SLOT + of] = cmul_sub(sf[3], mf[0], sf[1], mf[2]);
        w[5 * C::SLOT + of] = cmul_sub(sf[1], mf[1], sf[2], mf[0]);
        if (og != of) {
#pragma unroll
            for (int c = 0; c < 3; c++) w[(size_t)c * C::SLOT + og] = cmul(mg[c], sg[0]);
            w[3 * C::SLOT + og] = cmul_sub(sg[2], mg[2], sg[3], mg[1]);
            w[4 * C::SLOT + og] = cmul_sub(sg[3], mg[0], sg[1], mg[2]);
            w[5 * C::SLOT + og] = cmul_sub(sg[1], mg[1], sg[2], mg[0]);
        }
    }
}
// the same with two source sets: K^ of the own level in the E slots, K^' of the neighbour's level in the B slots, both of the
// PARTNER task; tw = the table of main_tw4 (entry (kx, 0) holds cos and -sin of 2 pi kx / M)
template <int ND> ION_HD void main_phase_product2_mirror(int tid, const float2* shat, const float2* S2c, const float4* tw, int kx0, int np, float2* W) {
    typedef Cfg<ND> C;
    constexpr int MM = C::M * C::M;
    constexpr size_t CS = (size_t)C::H * C::SLOT;
    for (int idx = tid; idx < np * MM; idx += C::T) {
        const int p = idx / MM, f = idx % MM;
        const int kz = f / C::M, ky = f % C::M;
        const int gz = (C::M - kz) & (C::M - 1), gy = (C::M - ky) & (C::M - 1);
        if (gz * C::M + gy < f) continue;
        const int of = kz * C::ROW + ky, og = gz * C::ROW + gy;
        const size_t base = (size_t)(kx0 + p) * C::SLOT;
        float2 sf[4], sg[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sf[j] = ION_LDG2(shat + (size_t)j * CS + base + of);
            sg[j] = ION_LDG2(shat + (size_t)j * CS + base + og);
        }
        float2* w = W + (size_t)p * C::PLANE;
        float2 kf[3], kg[3], lf[3], lg[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            kf[c] = w[(size_t)c * C::SLOT + of]; kg[c] = w[(size_t)c * C::SLOT + og];
            lf[c] = w[(size_t)(3 + c) * C::SLOT + of]; lg[c] = w[(size_t)(3 + c) * C::SLOT + og];
        }
        // s^' from its compact form: sign flips when kz or ky leaves [0, ND)
        const float sgf = ((kz >= ND) != (ky >= ND)) ? -1.0f : 1.0f, sgg = ((gz >= ND) != (gy >= ND)) ? -1.0f : 1.0f;
        const float2* cf = S2c + (size_t)p * 4 * C::CSLOT + (kz & (ND - 1)) * ND + (ky & (ND - 1));
        const float2* cg = S2c + (size_t)p * 4 * C::CSLOT + (gz & (ND - 1)) * ND + (gy & (ND - 1));
        float2 tf[4], tg[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 a = cf[(size_t)j * C::CSLOT], b = cg[(size_t)j * C::CSLOT];
            tf[j] = make_float2(sgf * a.x, sgf * a.y);
            tg[j] = make_float2(sgg * b.x, sgg * b.y);
        }
        const float4 t4 = tw[(size_t)(kx0 + p) * (ND / 2)];
        const float2 ph = make_float2(t4.y, -t4.w);  // e^{+2 pi i kx / M}
        float2 mf[3], mg[3], nf[3], ng[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float sc = c == 0 ? -1.0f : 1.0f;
            mf[c] = conj_sign(kg[c], sc); mg[c] = conj_sign(kf[c], sc);
            nf[c] = cmul(ph, conj_sign(lg[c], sc)); ng[c] = cmul(ph, conj_sign(lf[c], sc));
        }
        {
            const float2 E0 = cmul(mf[0], sf[0]), E1 = cmul(mf[1], sf[0]), E2 = cmul(mf[2], sf[0]);
            const float2 B0 = cmul_sub(sf[2], mf[2], sf[3], mf[1]), B1 = cmul_sub(sf[3], mf[0], sf[1], mf[2]), B2 = cmul_sub(sf[1], mf[1], sf[2], mf[0]);
            const float2 e0 = cmul(nf[0], tf[0]), e1 = cmul(nf[1], tf[0]), e2 = cmul(nf[2], tf[0]);
            const float2 b0 = cmul_sub(tf[2], nf[2], tf[3], nf[1]), b1 = cmul_sub(tf[3], nf[0], tf[1], nf[2]), b2 = cmul_sub(tf[1], nf[1], tf[2], nf[0]);
            w[of] = make_float2(E0.x + e0.x, E0.y + e0.y);
            w[C::SLOT + of] = make_float2(E1.x + e1.x, E1.y + e1.y);
            w[2 * C::SLOT + of] = make_float2(E2.x + e2.x, E2.y + e2.y);
            w[3 * C::SLOT + of] = make_float2(B0.x + b0.x, B0.y + b0.y);
            w[4 * C::SLOT + of] = make_float2(B1.x + b1.x, B1.y + b1.y);
            w[5 * C::SLOT + of] = make_float2(B2.x + b2.x, B2.y + b2.y);
        }
        if (og != of) {
            const float2 E0 = cmul(mg[0], sg[0]), E1 = cmul(mg[1], sg[0]), E2 = cmul(mg[2], sg[0]);
            const float2 B0 = cmul_sub(sg[2], mg[2], sg[3], mg[1]), B1 = cmul_sub(sg[3], mg[0], sg[1], mg[2]), B2 = cmul_sub(sg[1], mg[1], sg[2], mg[0]);
            const float2 e0 = cmul(ng[0], tg[0]), e1 = cmul(ng[1], tg[0]), e2 = cmul(ng[2], tg[0]);
            const float2 b0 = cmul_sub(tg[2], ng[2], tg[3], ng[1]), b1 = cmul_sub(tg[3], ng[0], tg[1], ng[2]), b2 = cmul_sub(tg[1], ng[1], tg[2], ng[0]);
            w[og] = make_float2(E0.x + e0.x, E0.y + e0.y);
            w[C::SLOT + og] = make_float2(E1.x + e1.x, E1.y + e1.y);
            w[2 * C::SLOT + og] = make_float2(E2.x + e2.x, E2.y + e2.y);
            w[3 * C::SLOT + og] = make_float2(B0.x + b0.x, B0.y + b0.y);
            w[4 * C::SLOT + og] = make_float2(B1.x + b1.x, B1.y + b1.y);
            w[5 * C::SLOT + og] = make_float2(B2.x + b2.x, B2.y + b2.y);
        }
    }
}
template <int ND> inline void main_stage2_host(const float2* khat_task, const float2* khat2_task, const float2* shat2c, int kx0, int np, float2* W, float2* S2c) {
    typedef Cfg<ND> C;
    for (int p = 0; p < np; p++)
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < C::CSLOT; i++) S2c[((size_t)p * 4 + j) * C::CSLOT + i] = shat2c[((size_t)j * C::H + kx0 + p) * C::CSLOT + i];
    for (int p = 0; p < np; p++)
        for (int c = 0; c < 3; c++)
            for (int i = 0; i < C::SLOT; i++) {
                W[(size_t)p * C::PLANE + (size_t)c * C::SLOT + i] = khat_task[((size_t)c * C::H + kx0 + p) * C::SLOT + i];
                W[(size_t)p * C::PLANE + (size_t)(3 + c) * C::SLOT + i] = khat2_task[((size_t)c * C::H + kx0 + p) * C::SLOT + i];
            }
}
// phase 2: inverse FFT along kz of every column (p, spectrum, ky); only the ND outputs z that exist are kept (in place)
template <int ND> ION_HD void main_phase_z(int tid, int np, float2* W) {
    typedef Cfg<ND> C;
    for (int idx = tid; idx < np * 6 * C::M; idx += C::T) {
        const int ps = idx / C::M, ky = idx % C::M;  // ps = p * 6 + spectrum
        float2* col = W + (size_t)ps * C::M * C::ROW + ky;
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = col[i * C::ROW];
        fft_reg<C::M, true>(a);
#pragma unroll
        for (int i = 0; i < ND; i++) col[i * C::ROW] = a[i];
    }
}
// phase 3: inverse FFT along ky of every row (p, spectrum, z < ND), pruned to y < ND (in place)
template <int ND> ION_HD void main_phase_y(int tid, int np, float2* W) {
    typedef Cfg<ND> C;
    for (int idx = tid; idx < np * 6 * ND; idx += C::T) {
        const int ps = idx / ND, z = idx % ND;
        float2* row = W + ((size_t)ps * C::M + z) * C::ROW;
        float2 a[C::M];
#pragma unroll
        for (int i = 0; i < C::M; i++) a[i] = row[i];
        fft_reg<C::M, true>(a);
#pragma unroll
        for (int i = 0; i < ND; i++) row[i] = a[i];
    }
}
// phase 4: the x axis as a direct Hermitian DFT.  Thread = (y, z, x group); acc[6][XPT] stays in registers across the iterations.
// tw4[kx][x/2] = (cos t(x), cos t(x+1), -sin t(x), -sin t(x+1)), t(x) = 2 pi kx x / M: two x outputs per packed FMA.
template <int ND> ION_HD void main_phase_accumulate(int tid, int kx0, int np, const float2* W, const float4* tw4, float (&acc)[6][Cfg<ND>::XPT]) {
    typedef Cfg<ND> C;
    constexpr int YZ = ND * ND;
    if (tid >= C::TACC) return;
    const int yz = tid % YZ, xg = tid / YZ;
    const int y = yz % ND, z = yz / ND;
    for (int p = 0; p < np; p++) {
        const int kx = kx0 + p;
        float2 v[6];
#pragma unroll
        for (int s = 0; s < 6; s++) v[s] = W[((size_t)(p * 6 + s) * C::M + z) * C::ROW + y];
        const float4* t4 = tw4 + (size_t)kx * (ND / 2) + xg * (C::XPT / 2);
#pragma unroll
        for (int i = 0; i < C::XPT; i += 2) {
            const float4 cs = t4[i / 2];
#if defined(__CUDA_ARCH__) && ION_FFT_PACKED
            const float2 c2 = make_float2(cs.x, cs.y), s2 = make_float2(cs.z, cs.w);
#pragma unroll
            for (int s = 0; s < 6; s++) {
                float2 a2 = make_float2(acc[s][i], acc[s][i + 1]);
                a2 = __ffma2_rn(make_float2(v[s].x, v[s].x), c2, __ffma2_rn(make_float2(v[s].y, v[s].y), s2, a2));
                acc[s][i] = a2.x;
                acc[s][i + 1] = a2.y;
            }
#else
#pragma unroll
            for (int s = 0; s < 6; s++) {
                acc[s][i] = fmaf(v[s].x, cs.x, fmaf(v[s].y, cs.z, acc[s][i]));
                acc[s][i + 1] = fmaf(v[s].x, cs.y, fmaf(v[s].y, cs.w, acc[s][i + 1]));
            }
#endif
        }
    }
}
// the table of phase 4: entry (kx, j) for j < ND / 2
template <int ND> ION_HD float4 main_tw4(int kx, int j) {
    typedef Cfg<ND> C;
    const int a0 = ((kx * (2 * j)) & (C::M - 1)) * (32 / C::M), a1 = ((kx * (2 * j + 1)) & (C::M - 1)) * (32 / C::M);
    return make_float4(tw_cos32(a0), tw_cos32(a1), -tw_sin32(a0), -tw_sin32(a1));
}
// Hand-over of a task's sums.  The cells of one task are ds apart in x (one cell per 64 bytes at ds = 16), so writing E_dyn / B_dyn
// -- and reading flags, E_stat, B_stat -- from here would touch one 32-byte sector per 4-byte access: 53 000 sector operations per
// task, measured as 60 % of the kernel.  Instead the sums go to a scratch field whose ROWS ARE PERMUTED: inside row (y, z) the value
// of cell x = bx * ds + ox sits at ox * ND + bx, so a thread's ND block positions are contiguous (64 bytes, vector stores).
// main_combine_row below undoes the permutation with fully coalesced traffic.  Layout: scratch[6][nz][ny][nx] floats.
template <int ND> ION_HD void main_phase_store(int tid, const Geom& g, const Task t, float* scratch, bool accumulate, float (&acc)[6][Cfg<ND>::XPT]) {
    typedef Cfg<ND> C;
    constexpr int YZ = ND * ND;
    if (tid >= C::TACC) return;
    const int yz = tid % YZ, xg = tid / YZ;
    const uint32_t y = (uint32_t)(yz % ND) * g.dsy + t.oy;
    const uint32_t z = ((uint32_t)(yz / ND) + (uint32_t)ND * t.wz) * g.dsz + t.oz;
    if (z >= g.nz || y >= g.ny) return;
    float* row = scratch + ((uint64_t)y + (uint64_t)z * g.ny) * g.nx + (uint64_t)t.ox * ND + (uint64_t)xg * C::XPT;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        float* p = row + (uint64_t)c * g.N;
        if (accumulate) {  // a second source set adds to what the first pass left in the scratch field
#pragma unroll
            for (int i = 0; i < C::XPT; i++) acc[c][i] += p[i];
        }
        if (C::XPT % 4 == 0) {
#pragma unroll
            for (int i = 0; i < C::XPT; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(acc[c][i], acc[c][i + 1], acc[c][i + 2], acc[c][i + 3]);
        } else if (C::XPT % 2 == 0) {
#pragma unroll
            for (int i = 0; i < C::XPT; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(acc[c][i], acc[c][i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < C::XPT; i++) p[i] = acc[c][i];
        }
    }
}
// sim.cl:986-992 for one row (y, z): E_dyn = E_stat + KE * e, B_dyn = B_stat + KMU * b for every non-solid, non-halo cell.
// `tile` holds the permuted row of one component, padded (ox * (ND + 1) + bx) so that the un-permuting reads are conflict-free.
template <int ND>
ION_HD void combine_load(int tid, int nthreads, const Geom& g, uint32_t y, uint32_t z, int c, const float* scratch, float* tile) {
    const float* row = scratch + (uint64_t)c * g.N + ((uint64_t)y + (uint64_t)z * g.ny) * g.nx;
    for (uint32_t i = tid; i < g.nx; i += nthreads) tile[(i / ND) * (ND + 1) + i % ND] = row[i];
}
template <int ND>
ION_HD void combine_write(int tid, int nthreads, const Geom& g, uint32_t y, uint32_t z, int c, const float* tile, const uint8_t* flags,
                          const float* E_stat, const float* B_stat, float* E_dyn, float* B_dyn) {
    const uint64_t base = ((uint64_t)y + (uint64_t)z * g.ny) * g.nx;
    const float* stat = (c < 3 ? E_stat : B_stat) + (uint64_t)(c % 3) * g.N + base;
    float* dyn = (c < 3 ? E_dyn : B_dyn) + (uint64_t)(c % 3) * g.N + base;
    const float k = c < 3 ? g.ke : g.kmu;
    for (uint32_t x = tid; x < g.nx; x += nthreads) {
        if ((g.dx > 1u) & (x == 0u || x >= g.nx - 1u)) continue;  // is_halo, sim.cl:899
        if ((flags[base + x] & 0x1Fu) == 0x01u) continue;         // (flags & TYPE_BO) == TYPE_S, sim.cl:900-902
        dyn[x] = stat[x] + k * tile[(x % g.dsx) * (ND + 1) + x / g.dsx];
    }
}

// ------------------------------------------------------------------------------------------------------
// Far slabs (sim.cl:957-983, domains two and more slabs below): at most 8^(D-2) + 8^(D-3) + ... sources, hundreds of cells away.
// Their field varies slowly over a few cells, so it is summed ONCE per FARB^3 block of cells as a second-order Taylor polynomial
// about the block centre (value, gradient, Hessian of q f(r) and w x f(r), f = r/|r|^3) and every cell evaluates the polynomial:
// the cost per cell no longer depends on the number of far sources.  Truncation error relative to the far field itself is
// (|delta| / R)^3, |delta| <= 6.1 cells for 8^3 blocks and 2.6 for 4^3.  The host picks the block size from the smallest
// (cell, source) distance R: 8^3 when R >= 40 |delta| (error <= 1.6e-5 of the far field), 4^3 when R >= 25 |delta| (<= 6.4e-5 of a
// contribution that is itself a fraction of the field); closer than that, the direct kernel sums these sources.
// Table layout: tens[FAR_T][nblocks], block index bx + nbx * (by + nby * bz); entries 0..29 = E part, 30..59 = B part, each
// F0[3] | G[3][3] | H[3][6] with the pair order (xx, xy, xz, yy, yz, zz).
// ------------------------------------------------------------------------------------------------------
constexpr int FAR_T = 60;
struct FarSource {
    float cx, cy, cz, q;
    float wx, wy, wz, pad;
};
ION_HD int far_pair(int l, int m) {  // index of the symmetric pair (l, m)
    const int a = l < m ? l : m, b = l < m ? m : l;
    return a == 0 ? b : (a == 1 ? 2 + b : 5);
}
// accumulates one source into the 60 tensor entries of the block centred at (x0, y0, z0)
ION_HD void far_accumulate(float x0, float y0, float z0, const FarSource& s, float (&t)[FAR_T]) {
    const float r[3] = {x0 - s.cx, y0 - s.cy, z0 - s.cz};
    const float R2 = fmaf(r[0], r[0], fmaf(r[1], r[1], r[2] * r[2]));
    const float inv = 1.0f / sqrtf(R2), inv2 = inv * inv, inv3 = inv * inv2, inv5 = inv3 * inv2, inv7 = inv5 * inv2;
    float f[3], df[3][3], d2[3][6];
    const float w[3] = {s.wx, s.wy, s.wz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        f[k] = r[k] * inv3;
#pragma unroll
        for (int l = 0; l < 3; l++) df[k][l] = (k == l ? inv3 : 0.0f) - 3.0f * r[k] * r[l] * inv5;
#pragma unroll
        for (int l = 0; l < 3; l++)
#pragma unroll
            for (int m = l; m < 3; m++)
                d2[k][far_pair(l, m)] = -3.0f * ((k == l ? r[m] : 0.0f) + (k == m ? r[l] : 0.0f) + (l == m ? r[k] : 0.0f)) * inv5 +
                                        15.0f * r[k] * r[l] * r[m] * inv7;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;  // (w x f)_i = w_j f_k - w_k f_j
        t[i] = fmaf(s.q, f[i], t[i]);
        t[30 + i] += w[j] * f[k] - w[k] * f[j];
#pragma unroll
        for (int l = 0; l < 3; l++) {
            t[3 + 3 * i + l] = fmaf(s.q, df[i][l], t[3 + 3 * i + l]);
            t[33 + 3 * i + l] += w[j] * df[k][l] - w[k] * df[j][l];
        }
#pragma unroll
        for (int p = 0; p < 6; p++) {
            t[12 + 6 * i + p] = fmaf(s.q, d2[i][p], t[12 + 6 * i + p]);
            t[42 + 6 * i + p] += w[j] * d2[k][p] - w[k] * d2[j][p];
        }
    }
}
// value of component i (0..2) of the polynomial at offset (dx, dy, dz) from the block centre; `t` = the 30 entries of the E or B
// part.  The full form, kept as the definition: k_eb_combine evaluates the same polynomial after reducing it to a + dx (b + c dx)
// per (row, block), since dy and dz are constant along a row.
ION_HD float far_eval(const float* t, int i, float dx, float dy, float dz) {
    const float* G = t + 3 + 3 * i;
    const float* H = t + 12 + 6 * i;
    float v = t[i];
    v = fmaf(G[0], dx, fmaf(G[1], dy, fmaf(G[2], dz, v)));
    v = fmaf(0.5f * H[0], dx * dx, fmaf(H[1], dx * dy, fmaf(H[2], dx * dz, v)));
    v = fmaf(0.5f * H[3], dy * dy, fmaf(H[4], dy * dz, fmaf(0.5f * H[5], dz * dz, v)));
    return v;
}
// The same polynomial along a row of cells: dy and dz are constants there, so component i is a + dx (b + c dx).  F0, G[3], H[6] are
// the entries of component i (any stride: the kernel reads them straight from the [FAR_T][blocks] table).
ION_HD void far_reduce_x(float F0, float G0, float G1, float G2, float H0, float H1, float H2, float H3, float H4, float H5, float dy, float dz,
                         float& a, float& b, float& c) {
    float a0 = fmaf(G1, dy, fmaf(G2, dz, F0));
    a = fmaf(0.5f * H3, dy * dy, fmaf(H4, dy * dz, fmaf(0.5f * H5, dz * dz, a0)));
    b = fmaf(H1, dy, fmaf(H2, dz, G0));
    c = 0.5f * H0;
}
ION_HD float far_eval_x(float a, float b, float c, float dx) { return fmaf(dx, fmaf(c, dx, b), a); }

}  // namespace ebfft
}  // namespace ion
