// fields.cu -- electromagnetic field update and LOD pyramid kernels.
//
// Behaviour: /root/reference/src/kernels/sim_kernels.cl ("sim.cl") update_e_b_dynamic :897-993,
// lod_part_2_gather :864-895, clear_qu_lod :995-1003, lod helpers :425-447.
//
// B200 design of update_e_b_dynamic (compute bound: 8^depth source terms per cell):
//   1. k_build_sources turns the LOD pyramid section the reference loops over (own finest level window
//      [NUM_LOD_OWN-8^D, NUM_LOD_OWN) -- quirk Q5 kept -- plus the foreign-domain levels with their shifted
//      centres, quirk Q8 kept) into a flat table {cx,cy,cz,q | vx,vy,vz,-} of 32 B entries, once per step.
//   2. k_update_e_b stages that table through shared memory in chunks (all threads of a CTA walk the same
//      source list, so every LDS is a conflict-free broadcast) and keeps E/B accumulators in registers.
//      r/|r|^3 is evaluated as r*rsqrt(r^2)^3 (one MUFU instead of sqrt + 3 IEEE divisions); results agree with
//      the reference expression to ~1e-6 relative, inside the stated E/B tolerance (tests/test_gpu_parity.py).
//   3. the near-cell loop of sim.cl:919-938 (only non-empty for depth <= 2 or N > 2^16, quirk Q4) reads Q,u
//      directly; neighbouring threads read the same addresses, which L1 serves as broadcasts.
#include <cstdlib>

#include "lattice.cuh"

namespace ion {

struct __align__(16) LodSource {
    float cx, cy, cz, q;
    float vx, vy, vz;
    uint32_t d;  // LOD index d of the reference loop (for the d==ndi self-skip), 0xFFFFFFFF for foreign entries
};

// One foreign domain's contribution to update_e_b_dynamic (sim.cl:957-983), described for the packed kernel: where its
// pyramid level sits in QU_lod (entry0) and in the flat source table (flat0), its level, and the centre shift of
// sim.cl:970-972 (quirks Q8/Q18 included).  `fast` = the level is lod_depth-1 and its block is exactly two own blocks
// wide in x, so the (cell, source) tile is again Toeplitz-like and the packed tile code applies.
struct ForeignDesc {
    uint32_t entry0, flat0, count, level;
    float sx, sy, sz;
    uint32_t fast;
};
constexpr int MAX_FOREIGN = 32;
struct ForeignSet {
    uint32_t n;  // 0: no descriptors, the flat table is walked as a whole
    ForeignDesc d[MAX_FOREIGN];
};
// centres shifted by the uint wrap of quirk Q18 lie ~4.29e9 cells away: their terms are < 1e-16 of the sums they are
// added to.  The fast path skips them (the deterministic path keeps every term).
#define ION_FAR_SHIFT 1.0e9f

// lod_coordinates, sim.cl:440-447
__device__ __forceinline__ void lod_coordinates(const KArgs& a, uint32_t n, uint32_t d, float& cx, float& cy, float& cz) {
    const uint32_t nd = 1u << d;
    const float dsx = (float)(a.nx / nd), dsy = (float)(a.ny / nd), dsz = (float)(a.nz / nd);
    const uint32_t t = n % (nd * nd);
    cx = (float)(t % nd) * dsx + (0.5f * dsx);
    cy = (float)(t / nd) * dsy + (0.5f * dsy);
    cz = (float)(n / (nd * nd)) * dsz + (0.5f * dsz);
}

__device__ __forceinline__ uint32_t to_d3(uint32_t x) { return x * x * x; }  // to_d for 3-D sets, sim.cl:110-116

// number of sources update_e_b_dynamic visits: own window + foreign levels (sim.cl:943,960-983)
__host__ __device__ inline uint32_t source_count(uint32_t lod_depth, uint32_t n_lod_own, uint32_t dx, uint32_t dy, uint32_t dz,
                                                 uint32_t di) {
    const uint32_t fine = (1u << lod_depth) * (1u << lod_depth) * (1u << lod_depth);
    uint32_t cnt = n_lod_own >= fine ? fine : n_lod_own;
    const uint32_t cdx = (di % (dx * dy)) % dx, cdy = (di % (dx * dy)) / dx, cdz = di / (dx * dy);
    for (uint32_t d = 0; d < dx * dy * dz; d++) {
        if (d == di) continue;
        const int fx = (int)((d % (dx * dy)) % dx), fy = (int)((d % (dx * dy)) / dx), fz = (int)(d / (dx * dy));
        int dist = abs((int)cdx - fx);
        if (abs((int)cdy - fy) > dist) dist = abs((int)cdy - fy);
        if (abs((int)cdz - fz) > dist) dist = abs((int)cdz - fz);
        const int depth = (int)lod_depth - dist > 0 ? (int)lod_depth - dist : 0;
        cnt += (1u << depth) * (1u << depth) * (1u << depth);
    }
    return cnt;
}

__global__ void k_build_sources(const __grid_constant__ KArgs a, LodSource* __restrict__ src, uint32_t count) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t fine = to_d3(1u << a.lod_depth);
    const uint32_t lo = a.n_lod_own >= fine ? a.n_lod_own - fine : 0u;  // imax(NUM_LOD_OWN - 8^D, 0), sim.cl:943
    const uint32_t own = a.n_lod_own - lo;
    LodSource s;
    uint32_t entry;
    if (k < own) {
        const uint32_t d = lo + k;
        lod_coordinates(a, d, a.lod_depth, s.cx, s.cy, s.cz);  // sim.cl:945 (index d incl. offset: quirk Q5)
        entry = d;
        s.d = d;
    } else {
        // foreign domains in ascending index, skipping self (sim.cl:958-983)
        uint32_t rem = k - own, offset = a.n_lod_own;
        const uint32_t dxy = a.dx * a.dy;
        const int cdx = (int)((a.di % dxy) % a.dx), cdy = (int)((a.di % dxy) / a.dx), cdz = (int)(a.di / dxy);
        s.cx = s.cy = s.cz = 0.0f;
        entry = 0u;
        for (uint32_t d = 0; d < dxy * a.dz; d++) {
            if (d == a.di) continue;
            const int ddx = cdx - (int)((d % dxy) % a.dx), ddy = cdy - (int)((d % dxy) / a.dx), ddz = cdz - (int)(d / dxy);
            int dist = abs(ddx);
            if (abs(ddy) > dist) dist = abs(ddy);
            if (abs(ddz) > dist) dist = abs(ddz);
            const uint32_t depth = (int)a.lod_depth - dist > 0 ? (uint32_t)((int)a.lod_depth - dist) : 0u;
            const uint32_t n_fd = to_d3(1u << depth);
            if (rem < n_fd) {
                lod_coordinates(a, rem, depth, s.cx, s.cy, s.cz);
                // sim.cl:970-972: halo-inclusive shift (quirk Q8); `domain_diff.x * DEF_NX` is int * uint = uint in
                // OpenCL C, so a negative domain difference wraps to ~4.29e9 before the float conversion (quirk Q18)
                s.cx -= (float)((uint32_t)ddx * a.nx);
                s.cy -= (float)((uint32_t)ddy * a.ny);
                s.cz -= (float)((uint32_t)ddz * a.nz);
                entry = offset + rem;
                break;
            }
            rem -= n_fd;
            offset += n_fd;
        }
        s.d = 0xFFFFFFFFu;
    }
    const float4 quv = reinterpret_cast<const float4*>(a.QU_lod)[entry];
    s.q = quv.x;
    s.vx = quv.y;
    s.vy = quv.z;
    s.vz = quv.w;
    src[k] = s;
}

template <bool EXACT>
__device__ __forceinline__ void pre_field(float rx, float ry, float rz, float& px, float& py, float& pz) {
    if (EXACT) {  // vec_r / cbmagnitude(vec_r), sim.cl:91-94,931-932
        const float l = sqrtf(rx * rx + ry * ry + rz * rz);
        const float l3 = l * l * l;
        px = rx / l3; py = ry / l3; pz = rz / l3;
    } else {
        const float r2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
        const float ri = r2 > 0.0f ? rsqrtf(r2) : 0.0f;
        const float ri3 = ri * ri * ri;
        px = rx * ri3; py = ry * ri3; pz = rz * ri3;
    }
}
// e += q*pre; b += q*cross(v,pre)   (sim.cl:934-935).  Fast mode gets w = q*v from the staged table.
template <bool EXACT>
__device__ __forceinline__ void accumulate_pair(float* e, float* b, const float4 s, float px, float py, float pz, bool skip) {
    if (EXACT) {
        float tx = s.x * px, ty = s.x * py, tz = s.x * pz;
        float cx = s.x * (s.z * pz - s.w * py), cy = s.x * (s.w * px - s.y * pz), cz = s.x * (s.y * py - s.z * px);
        if (skip) { tx = ty = tz = cx = cy = cz = 0.0f; }
        e[0] += tx; e[1] += ty; e[2] += tz;
        b[0] += cx; b[1] += cy; b[2] += cz;
    } else {
        const float q = skip ? 0.0f : s.x, wx = skip ? 0.0f : s.y, wy = skip ? 0.0f : s.z, wz = skip ? 0.0f : s.w;
        e[0] = fmaf(q, px, e[0]); e[1] = fmaf(q, py, e[1]); e[2] = fmaf(q, pz, e[2]);
        b[0] = fmaf(wy, pz, fmaf(-wz, py, b[0]));
        b[1] = fmaf(wz, px, fmaf(-wx, pz, b[1]));
        b[2] = fmaf(wx, py, fmaf(-wy, px, b[2]));
    }
}

constexpr int EB_BLOCK = 256;
constexpr int EB_CHUNK = 1024;  // sources per shared-memory stage (32 KB)

template <bool EXACT>
__global__ void __launch_bounds__(EB_BLOCK) k_update_e_b(const __grid_constant__ KArgs a, const LodSource* __restrict__ src,
                                                          const uint32_t count) {
    __shared__ float4 s_pos[EB_CHUNK];  // cx,cy,cz,q
    __shared__ float4 s_vel[EB_CHUNK];  // vx,vy,vz,bit-cast d
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    bool active = x < a.nx && !is_halo(a, x, y, z);  // sim.cl:899
    const uint32_t n = (active ? x : 0u) + (y + z * a.ny) * a.nx;
    if (active) active = (a.flags[n] & ION_TYPE_BO) != ION_TYPE_S;  // sim.cl:900-902
    const uint64_t N = a.N;
    const float px = (float)x, py = (float)y, pz = (float)z;
    float e[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f};

    // ---- close distance: individual cells of the near block (sim.cl:907-938) ----
    if (active) {
        const uint32_t nd = 1u << a.lod_depth;
        const uint32_t sh = nd < 32u ? (1u << nd) : 0u;  // 1<<nd, sim.cl:908 (depth <= 4 enforced at create)
        const uint32_t dsx = sh && a.nx / sh > 1u ? a.nx / sh : 1u;
        const uint32_t dsy = sh && a.ny / sh > 1u ? a.ny / sh : 1u;
        const uint32_t dsz = sh && a.nz / sh > 1u ? a.nz / sh : 1u;
        if (dsx * dsy * dsz > 1u) {
            const uint32_t xu = min((x / dsx) * dsx + dsx, a.dx > 1u ? a.nx - 1u : a.nx);
            const uint32_t yu = min((y / dsy) * dsy + dsy, a.dy > 1u ? a.ny - 1u : a.ny);
            const uint32_t zu = min((z / dsz) * dsz + dsz, a.dz > 1u ? a.nz - 1u : a.nz);
            for (uint32_t xc = max((x / dsx) * dsx, a.dx > 1u ? 1u : 0u); xc < xu; xc++) {
                for (uint32_t yc = max((y / dsy) * dsy, a.dy > 1u ? 1u : 0u); yc < yu; yc++) {
                    for (uint32_t zc = max((z / dsz) * dsz, a.dz > 1u ? 1u : 0u); zc < zu; zc++) {
                        const uint32_t nc = xc + (yc + zc * a.ny) * a.nx;
                        if (nc == n) continue;
                        const float qc = a.Q[nc];
                        if (qc == 0.0f) continue;
                        float4 s = make_float4(qc, a.u[nc], a.u[N + nc], a.u[2ull * N + nc]);
                        if (!EXACT) { s.y *= s.x; s.z *= s.x; s.w *= s.x; }
                        float fx, fy, fz;
                        pre_field<EXACT>(px - (float)xc, py - (float)yc, pz - (float)zc, fx, fy, fz);
                        accumulate_pair<EXACT>(e, b, s, fx, fy, fz, false);
                    }
                }
            }
        }
    }

    // ---- medium + large distance: the flat LOD source table (sim.cl:940-983) ----
    const uint32_t ndi = active ? lod_index(a, x, y, z, a.lod_depth) : 0u;  // sim.cl:941
    for (uint32_t base = 0; base < count; base += EB_CHUNK) {
        const uint32_t m = min((uint32_t)EB_CHUNK, count - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) {
            const float4* p = reinterpret_cast<const float4*>(src + base + k);
            s_pos[k] = __ldg(p);
            float4 v = __ldg(p + 1);
            if (!EXACT) { const float q = s_pos[k].w; v.x *= q; v.y *= q; v.z *= q; }
            s_vel[k] = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (uint32_t k = 0; k < m; k++) {
                const float4 c = s_pos[k];
                const float4 v = s_vel[k];
                float fx, fy, fz;
                pre_field<EXACT>(px - c.x, py - c.y, pz - c.z, fx, fy, fz);
                // self-skip (sim.cl:944): the own LOD block contributes nothing (its centre may coincide with the cell)
                accumulate_pair<EXACT>(e, b, make_float4(c.w, v.x, v.y, v.z), fx, fy, fz, __float_as_uint(v.w) == ndi);
            }
        }
    }
    if (!active) return;
    // sim.cl:986-992
    a.E_dyn[n] = a.E_stat[n] + a.ke * e[0];
    a.E_dyn[N + n] = a.E_stat[N + n] + a.ke * e[1];
    a.E_dyn[2ull * N + n] = a.E_stat[2ull * N + n] + a.ke * e[2];
    a.B_dyn[n] = a.B_stat[n] + a.kmu * b[0];
    a.B_dyn[N + n] = a.B_stat[N + n] + a.kmu * b[1];
    a.B_dyn[2ull * N + n] = a.B_stat[2ull * N + n] + a.kmu * b[2];
}

// ------------------------------------------------------------------------------------------------------
// Tiled update_e_b_dynamic: the LOD source grid is regular, so r = cell - centre depends only on the block
// DIFFERENCE along x and on the cell's offset inside its block.  A thread owns the ND cells of one x-row that share
// the in-block offset ox (x = k*dsx + ox, k = 0..ND-1).  For one row of ND sources (fixed cy, cz) the ND x ND
// (cell, source) pairs need only 2*ND-1 distinct r/|r|^3 vectors (a Toeplitz tile): those are evaluated once per
// diagonal and reused by up to ND pairs, which cuts the work per pair from ~24 instructions (sub, 3 fma, rsqrt,
// 6 mul, 3 add, 6 fma, 2 LDS) to the 9 FMAs of the accumulation plus 1/8 of a vector evaluation.
//   EXACT = false: rsqrt-based r/|r|^3, fused accumulation, sources staged as (q, q*v).
//   EXACT = true : the reference's arithmetic to the bit -- sqrt, cube, three IEEE divisions, unfused
//                  `e += q*pre; b += q*cross(v,pre)` -- and, per cell, the reference's summation order (near cells,
//                  own LODs by ascending index, foreign LODs); with the ordered LOD deposit this makes E_dyn/B_dyn
//                  bit-identical to the reference kernels run sequentially.
// Requirements: depth 3 or 4 (ND = 8 / 16) and nx % ND == 0; everything else (other depths, ragged x) uses
// k_update_e_b above.  Quirks Q4/Q5/Q7/Q8/Q18 are reproduced: sources are addressed by their table index d, the
// skipped own block is the one whose INDEX equals lod_index(cell), foreign levels keep their shifted centres.
// ------------------------------------------------------------------------------------------------------
constexpr int EBT_BLOCK = 128;
static inline size_t fine_bytes(int nd) { return (size_t)nd * nd * nd * sizeof(float4); }

template <int ND, bool EXACT>
__global__ void __launch_bounds__(EBT_BLOCK) k_update_e_b_tiled(const __grid_constant__ KArgs a, const LodSource* __restrict__ foreign,
                                                                  const uint32_t n_foreign) {
    extern __shared__ float4 s_tab[];  // source table by index d - lo: (q, vx, vy, vz) or (q, q*vx, q*vy, q*vz)
    const uint32_t fine = (uint32_t)ND * ND * ND;
    const uint32_t lo = a.n_lod_own >= fine ? a.n_lod_own - fine : 0u;  // sim.cl:943
    const uint32_t cnt = a.n_lod_own - lo;
    for (uint32_t k = threadIdx.x; k < cnt; k += EBT_BLOCK) {
        float4 v = __ldg(reinterpret_cast<const float4*>(a.QU_lod) + lo + k);
        if (!EXACT) { v.y *= v.x; v.z *= v.x; v.w *= v.x; }
        s_tab[k] = v;
    }
    __syncthreads();
    // Rows of the source grid whose ND entries are all zero (no charge in those LOD blocks) add exactly nothing: flag them
    // once per block and skip them in the row loop (warp-uniform).  Single-domain runs always have such rows: the window
    // [NUM_LOD_OWN - 8^D, NUM_LOD_OWN) of quirk Q5 ends in 8^0 + ... + 8^(D-1) coarse-level slots that nothing ever fills.
    // The deterministic path visits every row (adding a zero can flip the sign of a -0 accumulator).
    __shared__ uint8_t s_nz[ND * ND + 2];
    {
        const uint32_t r_lo = lo / ND, n_rows = (a.n_lod_own + ND - 1) / ND - r_lo;
        for (uint32_t r = threadIdx.x; r < n_rows; r += EBT_BLOCK) {
            bool nz = EXACT;
            const int d0 = (int)((r_lo + r) * ND) - (int)lo;
            for (int cx = 0; cx < ND && !nz; cx++) {
                const int slot = d0 + cx;
                if (slot < 0 || slot >= (int)cnt) continue;
                const float4 v = s_tab[slot];
                nz = v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f;
            }
            s_nz[r] = nz ? 1 : 0;
        }
    }
    __syncthreads();
    const uint32_t dsx = a.nx / ND, dsy = a.ny / ND, dsz = a.nz / ND;
    const uint32_t m = blockIdx.x * EBT_BLOCK + threadIdx.x;  // (ox, y) of this thread, z = blockIdx.y
    if (m >= dsx * a.ny) return;
    const uint32_t ox = m % dsx, y = m / dsx, z = blockIdx.y;
    const uint64_t N = a.N;
    const uint32_t row0 = y * a.nx + z * a.nx * a.ny;
    const float fy = (float)y, fz = (float)z;
    const float dsxf = (float)dsx, dsyf = (float)dsy, dszf = (float)dsz;
    const uint32_t self_row = y / dsy + (z / dsz) * ND;  // (ndi - bx) / ND: row index of the skipped own block

    float e[ND][3], b[ND][3];
#pragma unroll
    for (int k = 0; k < ND; k++) { e[k][0] = e[k][1] = e[k][2] = 0.0f; b[k][0] = b[k][1] = b[k][2] = 0.0f; }

    // ---- close distance (sim.cl:907-938): only non-empty when nx >= 2^(2^depth+1), i.e. depth 3 and nx >= 512 ----
    {
        const uint32_t sh = 1u << ND;  // ND = 8 -> 256; ND = 16 -> 65536
        const uint32_t nsx = a.nx / sh > 1u ? a.nx / sh : 1u, nsy = a.ny / sh > 1u ? a.ny / sh : 1u, nsz = a.nz / sh > 1u ? a.nz / sh : 1u;
        if (nsx * nsy * nsz > 1u) {
            for (int k = 0; k < ND; k++) {
                const uint32_t x = (uint32_t)k * dsx + ox, n = row0 + x;
                const uint32_t xu = min((x / nsx) * nsx + nsx, a.dx > 1u ? a.nx - 1u : a.nx);
                const uint32_t yu = min((y / nsy) * nsy + nsy, a.dy > 1u ? a.ny - 1u : a.ny);
                const uint32_t zu = min((z / nsz) * nsz + nsz, a.dz > 1u ? a.nz - 1u : a.nz);
                float ek[3] = {0.f, 0.f, 0.f}, bk[3] = {0.f, 0.f, 0.f};
                for (uint32_t xc = max((x / nsx) * nsx, a.dx > 1u ? 1u : 0u); xc < xu; xc++)
                    for (uint32_t yc = max((y / nsy) * nsy, a.dy > 1u ? 1u : 0u); yc < yu; yc++)
                        for (uint32_t zc = max((z / nsz) * nsz, a.dz > 1u ? 1u : 0u); zc < zu; zc++) {
                            const uint32_t nc = xc + (yc + zc * a.ny) * a.nx;
                            if (nc == n) continue;
                            const float qc = a.Q[nc];
                            if (qc == 0.0f) continue;
                            float4 s = make_float4(qc, a.u[nc], a.u[N + nc], a.u[2ull * N + nc]);
                            if (!EXACT) { s.y *= s.x; s.z *= s.x; s.w *= s.x; }
                            float px, py, pz;
                            pre_field<EXACT>((float)x - (float)xc, fy - (float)yc, fz - (float)zc, px, py, pz);
                            accumulate_pair<EXACT>(ek, bk, s, px, py, pz, false);
                        }
#pragma unroll
                for (int kk = 0; kk < ND; kk++)
                    if (kk == k) { e[kk][0] = ek[0]; e[kk][1] = ek[1]; e[kk][2] = ek[2]; b[kk][0] = bk[0]; b[kk][1] = bk[1]; b[kk][2] = bk[2]; }
            }
        }
    }

    // ---- own LODs, ascending index d = cx + ND*row (sim.cl:940-955) ----
    const float rx0 = (float)ox - 0.5f * dsxf;  // r_x for block difference 0: (k*dsx+ox) - (cx*dsx + dsx/2), k == cx
    const uint32_t row_lo = lo / ND, row_hi = (a.n_lod_own + ND - 1) / ND;
    for (uint32_t row = row_lo; row < row_hi; row++) {
        if (!s_nz[row - row_lo]) continue;  // no charge in this row of LOD blocks
        const uint32_t cy = row % ND, cz = row / ND;
        const float ry = fy - ((float)cy * dsyf + 0.5f * dsyf);
        const float rz = fz - ((float)cz * dszf + 0.5f * dszf);
        const bool is_self = row == self_row;
        const int d0 = (int)(row * ND) - (int)lo;  // table slot of cx = 0 (negative / past the end on the two ragged rows)
        if (d0 < 0 || d0 + ND > (int)cnt) {
            // ragged first / last row: entries outside [lo, n_lod_own) are not visited by the reference loop, so they
            // cannot simply get a zero weight (0 * r/|r|^3 is NaN when the cell sits on that block's centre)
            for (int k = 0; k < ND; k++) {
                float ek[3] = {0.f, 0.f, 0.f}, bk[3] = {0.f, 0.f, 0.f};
                bool first = true;
#pragma unroll
                for (int kk = 0; kk < ND; kk++)
                    if (kk == k) { ek[0] = e[kk][0]; ek[1] = e[kk][1]; ek[2] = e[kk][2]; bk[0] = b[kk][0]; bk[1] = b[kk][1]; bk[2] = b[kk][2]; }
                for (int cx = 0; cx < ND; cx++) {
                    const int slot = d0 + cx;
                    if (slot < 0 || slot >= (int)cnt || (is_self && cx == k)) continue;
                    float px, py, pz;
                    pre_field<EXACT>((float)(k - cx) * dsxf + rx0, ry, rz, px, py, pz);
                    accumulate_pair<EXACT>(ek, bk, s_tab[slot], px, py, pz, false);
                }
                (void)first;
#pragma unroll
                for (int kk = 0; kk < ND; kk++)
                    if (kk == k) { e[kk][0] = ek[0]; e[kk][1] = ek[1]; e[kk][2] = ek[2]; b[kk][0] = bk[0]; b[kk][1] = bk[1]; b[kk][2] = bk[2]; }
            }
            continue;
        }
        // diagonals from +(ND-1) down to -(ND-1): for a fixed cell k the sources cx = k - D are then visited in
        // ascending order, as in the reference loop
#pragma unroll
        for (int D = ND - 1; D >= -(ND - 1); D--) {
            // r_x = (k - cx)*dsx + ox - dsx/2, evaluated like the reference: (float)x - ((float)cx*dsx + 0.5*dsx);
            // every operand is a small integer or half-integer, so the value is exact and depends on D only
            const float rx = (float)D * dsxf + rx0;
            float px, py, pz;
            pre_field<EXACT>(rx, ry, rz, px, py, pz);
#pragma unroll
            for (int k = 0; k < ND; k++) {
                const int cx = k - D;
                if (cx < 0 || cx >= ND) continue;
                accumulate_pair<EXACT>(e[k], b[k], s_tab[d0 + cx], px, py, pz, (D == 0) && is_self);
            }
        }
    }

    // ---- foreign-domain LODs (sim.cl:957-983), flat list prepared by k_build_sources ----
    for (uint32_t f = 0; f < n_foreign; f++) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(foreign + f));
        float4 s = __ldg(reinterpret_cast<const float4*>(foreign + f) + 1);
        if (!EXACT && (fabsf(c.x) > ION_FAR_SHIFT || fabsf(c.y) > ION_FAR_SHIFT || fabsf(c.z) > ION_FAR_SHIFT)) continue;  // quirk Q18 terms
        s = make_float4(c.w, s.x, s.y, s.z);
        if (!EXACT) { s.y *= s.x; s.z *= s.x; s.w *= s.x; }
        const float ry = fy - c.y, rz = fz - c.z;
#pragma unroll
        for (int k = 0; k < ND; k++) {
            float px, py, pz;
            pre_field<EXACT>((float)((uint32_t)k * dsx + ox) - c.x, ry, rz, px, py, pz);
            accumulate_pair<EXACT>(e[k], b[k], s, px, py, pz, false);
        }
    }

    // ---- sim.cl:986-992 ----
#pragma unroll
    for (int k = 0; k < ND; k++) {
        const uint32_t x = (uint32_t)k * dsx + ox;
        const uint32_t n = row0 + x;
        if (is_halo(a, x, y, z)) continue;
        if ((a.flags[n] & ION_TYPE_BO) == ION_TYPE_S) continue;
        a.E_dyn[n] = a.E_stat[n] + a.ke * e[k][0];
        a.E_dyn[N + n] = a.E_stat[N + n] + a.ke * e[k][1];
        a.E_dyn[2ull * N + n] = a.E_stat[2ull * N + n] + a.ke * e[k][2];
        a.B_dyn[n] = a.B_stat[n] + a.kmu * b[k][0];
        a.B_dyn[N + n] = a.B_stat[N + n] + a.kmu * b[k][1];
        a.B_dyn[2ull * N + n] = a.B_stat[2ull * N + n] + a.kmu * b[k][2];
    }
}

// ------------------------------------------------------------------------------------------------------
// Circular packed-FP32 variant of the tiled kernel (fast path, EXACT = false only).
//
// Why: ncu on k_update_e_b_tiled<16,false> (profiles/r1_ncu_update_e_b.md) shows FMA pipe 54 %, issue 63 %, and the top
// stall "no_instruction": its fully unrolled 16x16 Toeplitz tile is ~56 KB of straight-line code per source row, well
// past the 32 KB L1.5 instruction cache.  The arithmetic floor is 9 FMA per (cell, source) pair; scalar FFMA already
// runs at full rate on sm_100 (tests/tools/fp32_peak.cu: 3.6e13 FMA/s scalar, 3.3e13 packed), so packing does not
// raise the FMA ceiling -- it halves the instruction count, which (a) makes the unrolled tile fit the instruction
// cache and (b) frees issue slots for the LDS / MUFU / address work next to a saturated FMA pipe.
//
// Tile: a thread owns NC cells k = kbase + kl (kl = 0..NC-1, same in-block offset ox) of one x-row; a source row has
// ND sources c.  Instead of walking the 2*ND-1 Toeplitz diagonals (ragged: 1..ND pairs each), the tile is walked
// CIRCULARLY: step D' = 0..ND-1 pairs cell kl with source c = (kl - D') mod ND, whose true diagonal is
// D = D' (kl >= D', "hi") or D' - ND (kl < D', "lo").  Every step then has exactly NC pairs and every (cell, source)
// pair occurs once.  FFMA2 lanes = cells (2j, 2j+1); their sources (c, c+1 mod ND) are one aligned entry of a
// shared-memory table that stores, for every source t of a row, the pair (S_t, S_(t+1) mod ND) as two float4
// {q_t, q_t', wx_t, wx_t'}, {wy_t, wy_t', wz_t, wz_t'} -- two broadcast LDS.128 feed nine FFMA2 (18 pair terms).
// r/|r|^3 enters as a scalar broadcast operand (hi or lo), or as a true (lo, hi) pair on the one lane pair per odd D'
// that straddles the wrap.  Everything is unrolled: ND*NC/2*11 + ~2*ND*14 instructions (16 KB for ND=16, NC=8).
// Per cell the sources are NOT visited in ascending order any more (cell kl starts at source kl and wraps), so E/B
// differ from k_update_e_b_tiled<ND,false> by summation-order rounding (~1e-7 relative); the deterministic path
// (EXACT) is untouched.  The own block (sim.cl:944) is removed by zeroing r/|r|^3 of its diagonal.
// ------------------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void pair_get(const float2 (&e2)[NC / 2][3], const float2 (&b2)[NC / 2][3], int k, float* ek, float* bk) {
#pragma unroll
    for (int j = 0; j < NC / 2; j++) {
        if (2 * j == k) { ek[0] = e2[j][0].x; ek[1] = e2[j][1].x; ek[2] = e2[j][2].x; bk[0] = b2[j][0].x; bk[1] = b2[j][1].x; bk[2] = b2[j][2].x; }
        if (2 * j + 1 == k) { ek[0] = e2[j][0].y; ek[1] = e2[j][1].y; ek[2] = e2[j][2].y; bk[0] = b2[j][0].y; bk[1] = b2[j][1].y; bk[2] = b2[j][2].y; }
    }
}
template <int NC>
__device__ __forceinline__ void pair_put(float2 (&e2)[NC / 2][3], float2 (&b2)[NC / 2][3], int k, const float* ek, const float* bk) {
#pragma unroll
    for (int j = 0; j < NC / 2; j++) {
        if (2 * j == k) { e2[j][0].x = ek[0]; e2[j][1].x = ek[1]; e2[j][2].x = ek[2]; b2[j][0].x = bk[0]; b2[j][1].x = bk[1]; b2[j][2].x = bk[2]; }
        if (2 * j + 1 == k) { e2[j][0].y = ek[0]; e2[j][1].y = ek[1]; e2[j][2].y = ek[2]; b2[j][0].y = bk[0]; b2[j][1].y = bk[1]; b2[j][2].y = bk[2]; }
    }
}
// ION_EB_MIX: how many of the nine packed FMAs per lane pair run as scalar FFMA pairs instead (see fma_pair)
#ifndef ION_EB_MIX
#define ION_EB_MIX 1
#endif
#ifndef ION_EB_MIX_FIRST
#define ION_EB_MIX_FIRST 1
#endif
#ifndef ION_EB_NOGUARD
#define ION_EB_NOGUARD 0
#endif
// Schedule experiments on the hot loop (A/B builds, update_e_b_dynamic at 256^3 depth 4; 0 = adopted: 22.07 ms):
// 1 = lane pairs descending 23.48 ms, 2 = cross product interleaved with E 22.48, 3 = register cap 232 22.41,
// 5 = r/|r|^3 with scalar FP32 22.57, 6 = step loop unrolled by 8 instead of 16 27.85.
#ifndef ION_EB_EXP
#define ION_EB_EXP 0
#endif
// nine FFMA2: e += q*p, b += w x p for two cells at once (same rounding sequence as accumulate_pair<false>)
__device__ __forceinline__ void fma_pair(float2* e, float2* b, const float4 A, const float4 B, const float2 PX, const float2 PY, const float2 PZ) {
    const float2 q = make_float2(A.x, A.y), wx = make_float2(A.z, A.w), wy = make_float2(B.x, B.y), wz = make_float2(B.z, B.w);
    const float2 NX = make_float2(-PX.x, -PX.y), NY = make_float2(-PY.x, -PY.y), NZ = make_float2(-PZ.x, -PZ.y);
    // ION_EB_MIX = S: the last S of the nine packed FMAs (order: e0 e1 e2 | b first stage 0 1 2 | b second stage 0 1 2) are issued
    // as two scalar FFMA each -- a scalar FFMA occupies the FMA pipe for 1 cycle per 32 FMAs, FFMA2 for 2.27 per 64, and the
    // kernel has spare issue slots (measured at 256^3: S = 0 23.55 ms, S = 1 22.32, S = 2 22.71, S = 3 23.05, S = 4 24.16, S = 6 25.61)
    auto f2 = [](float2 a, float2 b, float2 c, bool scalar) {
        return scalar ? make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)) : __ffma2_rn(a, b, c);
    };
    constexpr int S = ION_EB_MIX;
#if ION_EB_MIX_FIRST  // scalarise from the front (e0 first; measured S = 1: 22.07 ms) instead of from the back (22.30 ms)
#if ION_EB_EXP == 2
    const float2 v0 = f2(wz, NY, b[0], S >= 4), v1 = f2(wx, NZ, b[1], S >= 5), v2 = f2(wy, NX, b[2], S >= 6);
    e[0] = f2(q, PX, e[0], S >= 1);
    b[0] = f2(wy, PZ, v0, S >= 7);
    e[1] = f2(q, PY, e[1], S >= 2);
    b[1] = f2(wz, PX, v1, S >= 8);
    e[2] = f2(q, PZ, e[2], S >= 3);
    b[2] = f2(wx, PY, v2, S >= 9);
    return;
#endif
    e[0] = f2(q, PX, e[0], S >= 1);
    e[1] = f2(q, PY, e[1], S >= 2);
    e[2] = f2(q, PZ, e[2], S >= 3);
    const float2 u0 = f2(wz, NY, b[0], S >= 4), u1 = f2(wx, NZ, b[1], S >= 5), u2 = f2(wy, NX, b[2], S >= 6);
    b[0] = f2(wy, PZ, u0, S >= 7);
    b[1] = f2(wz, PX, u1, S >= 8);
    b[2] = f2(wx, PY, u2, S >= 9);
    return;
#endif
    e[0] = f2(q, PX, e[0], S >= 9);
    e[1] = f2(q, PY, e[1], S >= 8);
    e[2] = f2(q, PZ, e[2], S >= 7);
    const float2 t0 = f2(wz, NY, b[0], S >= 6), t1 = f2(wx, NZ, b[1], S >= 5), t2 = f2(wy, NX, b[2], S >= 4);
    b[0] = f2(wy, PZ, t0, S >= 3);
    b[1] = f2(wz, PX, t1, S >= 2);
    b[2] = f2(wx, PY, t2, S >= 1);
}

// VOL = true: every (step, lane pair) re-reads its source entry with a volatile LDS.128 (2 LDS per 9 FFMA2, few registers);
// VOL = false: the compiler may keep the whole row table (2*ND float4 = 8*ND registers) in registers across the steps.
template <bool VOL> __device__ __forceinline__ float4 lds128(const float4* p) {
    if (VOL) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
    return *p;
}

#if ION_EB_EXP == 3
#define ION_EB_BOUNDS(B) __maxnreg__(232)
#else
#define ION_EB_BOUNDS(B) __launch_bounds__(B, 1)
#endif
template <int ND, int NC, int BLOCK, bool VOL>
__global__ void ION_EB_BOUNDS(BLOCK) k_update_e_b_pair(const __grid_constant__ KArgs a, const LodSource* __restrict__ foreign,
                                                                const uint32_t n_foreign, const __grid_constant__ ForeignSet fs, const uint32_t mode) {
    // mode bit 0: foreign-domain sources only (the own pyramid was summed by the polyphase FFT kernels of eb_fft.cu);
    // mode bit 1: add to E_dyn / B_dyn instead of starting from E_stat / B_stat
    static_assert(ND % NC == 0 && NC % 2 == 0, "cells per thread");
    constexpr int PARTS = ND / NC;
    extern __shared__ float4 s_pair[];  // slot t -> {q_t, q_t', wx_t, wx_t'}, {wy_t, wy_t', wz_t, wz_t'}; t' = next source of the row, cyclic
    const uint32_t fine = (uint32_t)ND * ND * ND;
    const uint32_t lo = a.n_lod_own >= fine ? a.n_lod_own - fine : 0u;  // sim.cl:943
    const uint32_t cnt = (mode & 1u) ? 0u : a.n_lod_own - lo;
    for (uint32_t k = threadIdx.x; k < cnt; k += BLOCK) {
        const uint32_t d = lo + k;
        const uint32_t dn = (d % ND == ND - 1u) ? d - (ND - 1u) : d + 1u;  // cyclic successor inside the row
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.QU_lod) + d);
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dn >= lo && dn < a.n_lod_own) v1 = __ldg(reinterpret_cast<const float4*>(a.QU_lod) + dn);
        s_pair[2u * k] = make_float4(v0.x, v1.x, v0.y * v0.x, v1.y * v1.x);
        s_pair[2u * k + 1u] = make_float4(v0.z * v0.x, v1.z * v1.x, v0.w * v0.x, v1.w * v1.x);
    }
    __syncthreads();
    __shared__ uint8_t s_nz[ND * ND + 2];  // rows without any charge are skipped, see k_update_e_b_tiled
    {
        const uint32_t r_lo = lo / ND, n_rows = (mode & 1u) ? 0u : (a.n_lod_own + ND - 1) / ND - r_lo;
        for (uint32_t r = threadIdx.x; r < n_rows; r += BLOCK) {
            bool nz = false;
            const int d0 = (int)((r_lo + r) * ND) - (int)lo;
            for (int cx = 0; cx < ND && !nz; cx++) {
                const int slot = d0 + cx;
                if (slot < 0 || slot >= (int)cnt) continue;
                const float4 A = s_pair[2 * slot], B = s_pair[2 * slot + 1];
                nz = A.x != 0.0f || A.z != 0.0f || B.x != 0.0f || B.z != 0.0f;
            }
            s_nz[r] = nz ? 1 : 0;
        }
    }
    __syncthreads();
    const uint32_t dsx = a.nx / ND, dsy = a.ny / ND, dsz = a.nz / ND;
    const uint32_t m = blockIdx.x * BLOCK + threadIdx.x;  // (ox, part, y) of this thread, z = blockIdx.y
    if (m >= dsx * PARTS * a.ny) return;
    const uint32_t ox = m % dsx, part = (m / dsx) % PARTS, y = m / (dsx * PARTS), z = blockIdx.y;
    const uint32_t kbase = part * NC;
    const uint64_t N = a.N;
    const uint32_t row0 = y * a.nx + z * a.nx * a.ny;
    const float fy = (float)y, fz = (float)z;
    const float dsxf = (float)dsx, dsyf = (float)dsy, dszf = (float)dsz;
    const uint32_t self_row = y / dsy + (z / dsz) * ND;

    float2 e2[NC / 2][3], b2[NC / 2][3];
#pragma unroll
    for (int j = 0; j < NC / 2; j++)
#pragma unroll
        for (int c = 0; c < 3; c++) { e2[j][c] = make_float2(0.f, 0.f); b2[j][c] = make_float2(0.f, 0.f); }

    // ---- close distance (sim.cl:907-938), as in k_update_e_b_tiled ----
    {
        const uint32_t sh = 1u << ND;
        const uint32_t nsx = a.nx / sh > 1u ? a.nx / sh : 1u, nsy = a.ny / sh > 1u ? a.ny / sh : 1u, nsz = a.nz / sh > 1u ? a.nz / sh : 1u;
        if (nsx * nsy * nsz > 1u) {
            for (int kl = 0; kl < NC; kl++) {
                const uint32_t x = (kbase + (uint32_t)kl) * dsx + ox, n = row0 + x;
                const uint32_t xu = min((x / nsx) * nsx + nsx, a.dx > 1u ? a.nx - 1u : a.nx);
                const uint32_t yu = min((y / nsy) * nsy + nsy, a.dy > 1u ? a.ny - 1u : a.ny);
                const uint32_t zu = min((z / nsz) * nsz + nsz, a.dz > 1u ? a.nz - 1u : a.nz);
                float ek[3] = {0.f, 0.f, 0.f}, bk[3] = {0.f, 0.f, 0.f};
                for (uint32_t xc = max((x / nsx) * nsx, a.dx > 1u ? 1u : 0u); xc < xu; xc++)
                    for (uint32_t yc = max((y / nsy) * nsy, a.dy > 1u ? 1u : 0u); yc < yu; yc++)
                        for (uint32_t zc = max((z / nsz) * nsz, a.dz > 1u ? 1u : 0u); zc < zu; zc++) {
                            const uint32_t nc = xc + (yc + zc * a.ny) * a.nx;
                            if (nc == n) continue;
                            const float qc = a.Q[nc];
                            if (qc == 0.0f) continue;
                            const float4 s = make_float4(qc, a.u[nc] * qc, a.u[N + nc] * qc, a.u[2ull * N + nc] * qc);
                            float px, py, pz;
                            pre_field<false>((float)x - (float)xc, fy - (float)yc, fz - (float)zc, px, py, pz);
                            accumulate_pair<false>(ek, bk, s, px, py, pz, false);
                        }
                pair_put<NC>(e2, b2, kl, ek, bk);
            }
        }
    }

    // ---- own LODs (sim.cl:940-955) ----
    const float rx0 = (float)(kbase * dsx + ox) - 0.5f * dsxf;  // r_x of cell kl = 0 against source c = 0
    const uint32_t row_lo = lo / ND, row_hi = (mode & 1u) ? row_lo : (a.n_lod_own + ND - 1) / ND;
    for (uint32_t row = row_lo; row < row_hi; row++) {
        if (!s_nz[row - row_lo]) continue;  // no charge in this row of LOD blocks
        const uint32_t cy = row % ND, cz = row / ND;
        const float ry = fy - ((float)cy * dsyf + 0.5f * dsyf);
        const float rz = fz - ((float)cz * dszf + 0.5f * dszf);
        const bool is_self = row == self_row;
        const int d0 = (int)(row * ND) - (int)lo;
        if (d0 < 0 || d0 + ND > (int)cnt) {  // ragged first / last row: only slots inside [0, cnt) exist
            for (int kl = 0; kl < NC; kl++) {
                float ek[3], bk[3];
                pair_get<NC>(e2, b2, kl, ek, bk);
                const int k = (int)kbase + kl;
                for (int cx = 0; cx < ND; cx++) {
                    const int slot = d0 + cx;
                    if (slot < 0 || slot >= (int)cnt || (is_self && cx == k)) continue;
                    const float4 A = s_pair[2 * slot], B = s_pair[2 * slot + 1];
                    float px, py, pz;
                    pre_field<false>((float)(kl - cx) * dsxf + rx0, ry, rz, px, py, pz);
                    accumulate_pair<false>(ek, bk, make_float4(A.x, A.z, B.x, B.z), px, py, pz, false);
                }
                pair_put<NC>(e2, b2, kl, ek, bk);
            }
            continue;
        }
        const float4* __restrict__ srow = s_pair + 2 * d0;
        const int self_dl = is_self ? -(int)kbase : 0x7fffffff;  // local diagonal of the own block (true diagonal 0)
        const float ryz2 = fmaf(ry, ry, rz * rz);
#if ION_EB_EXP == 6
#pragma unroll 8
#else
#pragma unroll
#endif
        for (int Dp = 0; Dp < ND; Dp++) {
            // hi (.y): cells kl >= D' on local diagonal D'; lo (.x): cells kl < D' on local diagonal D' - ND.  Both r/|r|^3
            // vectors are evaluated together with packed FP32 (same roundings as pre_field<false>).
            const float2 rx = make_float2((float)(Dp - ND) * dsxf + rx0, (float)Dp * dsxf + rx0);
#if ION_EB_EXP == 5
            const float2 r2 = make_float2(fmaf(rx.x, rx.x, ryz2), fmaf(rx.y, rx.y, ryz2));
#else
            const float2 r2 = __ffma2_rn(rx, rx, make_float2(ryz2, ryz2));
#endif
#if ION_EB_NOGUARD
            // r = 0 can only occur on the self-skipped pair of the self row (the one source whose coordinates are the cell's own
            // block, see the comment above k_update_e_b_tiled), where ri3 is overwritten with 0 below.  Measured SLOWER (24.7 vs
            // 22.3 ms: the schedule ptxas finds without the predicates is worse), so the guards stay.
            const float2 ri = make_float2(rsqrtf(r2.x), rsqrtf(r2.y));
#else
            const float2 ri = make_float2(r2.x > 0.0f ? rsqrtf(r2.x) : 0.0f, r2.y > 0.0f ? rsqrtf(r2.y) : 0.0f);
#endif
#if ION_EB_EXP == 5
            float2 ri3 = make_float2(ri.x * ri.x * ri.x, ri.y * ri.y * ri.y);
#else
            float2 ri3 = __fmul2_rn(__fmul2_rn(ri, ri), ri);
#endif
            if (Dp - ND == self_dl) ri3.x = 0.0f;  // the own block contributes nothing (sim.cl:944)
            if (Dp == self_dl) ri3.y = 0.0f;
#if ION_EB_EXP == 5
            const float2 px = make_float2(rx.x * ri3.x, rx.y * ri3.y), py = make_float2(ry * ri3.x, ry * ri3.y), pz = make_float2(rz * ri3.x, rz * ri3.y);
#else
            const float2 px = __fmul2_rn(rx, ri3), py = __fmul2_rn(make_float2(ry, ry), ri3), pz = __fmul2_rn(make_float2(rz, rz), ri3);
#endif
#pragma unroll
            for (int jj = 0; jj < NC / 2; jj++) {
                const int j = ION_EB_EXP == 1 ? NC / 2 - 1 - jj : jj;
                const int t = ((2 * j - Dp) % ND + ND) % ND;  // source of lane .x; lane .y uses its cyclic successor
                const float4 A = lds128<VOL>(srow + 2 * t), B = lds128<VOL>(srow + 2 * t + 1);
                const bool xh = 2 * j >= Dp, yh = 2 * j + 1 >= Dp;
                fma_pair(e2[j], b2[j], A, B, make_float2(xh ? px.y : px.x, yh ? px.y : px.x), make_float2(xh ? py.y : py.x, yh ? py.y : py.x),
                         make_float2(xh ? pz.y : pz.x, yh ? pz.y : pz.x));
            }
        }
    }

    // ---- foreign-domain LODs (sim.cl:957-983) ----
    // generic: one source against the NC cells, r/|r|^3 evaluated per (cell, source) with packed FP32
    auto foreign_generic = [&](uint32_t f0, uint32_t f1) {
        for (uint32_t f = f0; f < f1; f++) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(foreign + f));
            if (fabsf(c.x) > ION_FAR_SHIFT || fabsf(c.y) > ION_FAR_SHIFT || fabsf(c.z) > ION_FAR_SHIFT) continue;  // quirk Q18 terms
            float4 s = __ldg(reinterpret_cast<const float4*>(foreign + f) + 1);
            s = make_float4(c.w, s.x * c.w, s.y * c.w, s.z * c.w);
            const float ry = fy - c.y, rz = fz - c.z;
            const float ryz2 = fmaf(ry, ry, rz * rz);
#pragma unroll
            for (int j = 0; j < NC / 2; j++) {
                const float2 rx = make_float2((float)((kbase + (uint32_t)(2 * j)) * dsx + ox) - c.x, (float)((kbase + (uint32_t)(2 * j + 1)) * dsx + ox) - c.x);
                const float2 r2 = __ffma2_rn(rx, rx, make_float2(ryz2, ryz2));
                const float2 ri = make_float2(r2.x > 0.0f ? rsqrtf(r2.x) : 0.0f, r2.y > 0.0f ? rsqrtf(r2.y) : 0.0f);
                const float2 ri3 = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                fma_pair(e2[j], b2[j], make_float4(s.x, s.x, s.y, s.y), make_float4(s.z, s.z, s.w, s.w), __fmul2_rn(rx, ri3),
                         __fmul2_rn(make_float2(ry, ry), ri3), __fmul2_rn(make_float2(rz, rz), ri3));
            }
        }
    };
    if (fs.n == 0u) {
        foreign_generic(0u, n_foreign);
    } else {
        for (uint32_t fi = 0; fi < fs.n; fi++) {
            const ForeignDesc& fd = fs.d[fi];
            if (fabsf(fd.sx) > ION_FAR_SHIFT || fabsf(fd.sy) > ION_FAR_SHIFT || fabsf(fd.sz) > ION_FAR_SHIFT) continue;  // quirk Q18 domains
            if (fd.fast == 2u) continue;  // already in E_dyn / B_dyn
            if (!fd.fast) {
                foreign_generic(fd.flat0, fd.flat0 + fd.count);
                continue;
            }
            // Level lod_depth-1 of a neighbour: ND/2 sources per row, each two own blocks wide.  Cells (2j, 2j+1) against
            // source cx have r_x = (2m-1)*dsx + ox + sx and 2m*dsx + ox + sx with m = j - cx: 2*(NC/2 + ND/2) - 1 packed
            // r/|r|^3 pairs serve the NC*ND/2 pairs of a row; the source enters as the broadcast operand.
            constexpr int NF = ND / 2;
            const float bsy = (float)(a.ny / NF), bsz = (float)(a.nz / NF);
            const float rxb = (float)(kbase * dsx + ox) + fd.sx;  // r_x of cell kl = 0 against a source whose centre is at 0
            const float4* __restrict__ lod = reinterpret_cast<const float4*>(a.QU_lod) + fd.entry0;
            for (uint32_t row = 0; row < (uint32_t)(NF * NF); row++) {
                const uint32_t cy = row % NF, cz = row / NF;
                const float ry = fy - (((float)cy * bsy + 0.5f * bsy) - fd.sy);
                const float rz = fz - (((float)cz * bsz + 0.5f * bsz) - fd.sz);
                const float ryz2 = fmaf(ry, ry, rz * rz);
                float4 src[NF];
                bool nz = false;
#pragma unroll
                for (int cx = 0; cx < NF; cx++) {
                    const float4 v = __ldg(lod + row * NF + cx);
                    src[cx] = make_float4(v.x, v.y * v.x, v.z * v.x, v.w * v.x);
                    nz = nz || src[cx].x != 0.0f || src[cx].y != 0.0f || src[cx].z != 0.0f || src[cx].w != 0.0f;
                }
                if (!nz) continue;  // same for every thread: the row holds no charge
#pragma unroll
                for (int m = -(NF - 1); m < NC / 2; m++) {
                    const float2 rx = make_float2((float)(2 * m - 1) * dsxf + rxb, (float)(2 * m) * dsxf + rxb);
                    const float2 r2 = __ffma2_rn(rx, rx, make_float2(ryz2, ryz2));
                    const float2 ri = make_float2(r2.x > 0.0f ? rsqrtf(r2.x) : 0.0f, r2.y > 0.0f ? rsqrtf(r2.y) : 0.0f);
                    const float2 ri3 = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                    const float2 px = __fmul2_rn(rx, ri3), py = __fmul2_rn(make_float2(ry, ry), ri3), pz = __fmul2_rn(make_float2(rz, rz), ri3);
#pragma unroll
                    for (int j = 0; j < NC / 2; j++) {
                        const int cx = j - m;
                        if (cx < 0 || cx >= NF) continue;
                        const float4 sv = src[cx];
                        fma_pair(e2[j], b2[j], make_float4(sv.x, sv.x, sv.y, sv.y), make_float4(sv.z, sv.z, sv.w, sv.w), px, py, pz);
                    }
                }
            }
        }
    }

    // ---- sim.cl:986-992 ----
#pragma unroll
    for (int kl = 0; kl < NC; kl++) {
        const uint32_t x = (kbase + (uint32_t)kl) * dsx + ox;
        const uint32_t n = row0 + x;
        if (is_halo(a, x, y, z)) continue;
        if ((a.flags[n] & ION_TYPE_BO) == ION_TYPE_S) continue;
        const int j = kl / 2;
        const float ex = (kl & 1) ? e2[j][0].y : e2[j][0].x, ey = (kl & 1) ? e2[j][1].y : e2[j][1].x, ez = (kl & 1) ? e2[j][2].y : e2[j][2].x;
        const float bx = (kl & 1) ? b2[j][0].y : b2[j][0].x, by = (kl & 1) ? b2[j][1].y : b2[j][1].x, bz = (kl & 1) ? b2[j][2].y : b2[j][2].x;
        const float* Eb = (mode & 2u) ? a.E_dyn : a.E_stat;
        const float* Bb = (mode & 2u) ? a.B_dyn : a.B_stat;
        a.E_dyn[n] = Eb[n] + a.ke * ex;
        a.E_dyn[N + n] = Eb[N + n] + a.ke * ey;
        a.E_dyn[2ull * N + n] = Eb[2ull * N + n] + a.ke * ez;
        a.B_dyn[n] = Bb[n] + a.kmu * bx;
        a.B_dyn[N + n] = Bb[N + n] + a.kmu * by;
        a.B_dyn[2ull * N + n] = Bb[2ull * N + n] + a.kmu * bz;
    }
}

// Deterministic LOD deposit (ION_EXT_DETERMINISTIC): the reference adds every cell's (Q, u/cells) to its finest-level LOD
// entry with float atomics (sim.cl:666-677), i.e. in execution order.  Here one thread per (entry, component) walks the
// cells of its block in ascending cell index -- the order of a sequential run of the reference kernel -- so the sums
// are reproducible and bit-identical to that run.  Needs block-aligned lattices (no lod_index overflow, quirk Q7).
__global__ void k_lod_deposit_ordered(const __grid_constant__ KArgs a) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nd = 1u << a.lod_depth;
    const uint32_t entry = tid >> 2, comp = tid & 3u;
    if (entry >= nd * nd * nd) return;
    const uint32_t dsx = a.nx / nd, dsy = a.ny / nd, dsz = a.nz / nd;
    const uint32_t bx = entry % nd, by = (entry / nd) % nd, bz = entry / (nd * nd);
    uint32_t off = 0u;
    if (a.dx > 1u || a.dy > 1u || a.dz > 1u)
        for (uint32_t d = 0u; d < a.lod_depth; d++) off += 1u << (d * 3u);
    const float* __restrict__ v = comp == 0u ? a.Q : a.lod_u + (uint64_t)(comp - 1u) * a.N;
    float acc = a.QU_lod[(uint64_t)(entry + off) * 4u + comp];
    for (uint32_t z = bz * dsz; z < (bz + 1u) * dsz; z++)
        for (uint32_t y = by * dsy; y < (by + 1u) * dsy; y++)
            for (uint32_t x = bx * dsx; x < (bx + 1u) * dsx; x++) {
                if (is_halo(a, x, y, z)) continue;
                const uint32_t n = x + (y + z * a.ny) * a.nx;
                if ((a.flags[n] & ION_TYPE_BO) == ION_TYPE_S) continue;
                acc += v[n];
            }
    a.QU_lod[(uint64_t)(entry + off) * 4u + comp] = acc;
}
cudaError_t launch_lod_deposit_ordered(const KArgs& a, cudaStream_t s) {
    const uint32_t nd = 1u << a.lod_depth;
    const uint32_t threads = nd * nd * nd * 4u;
    k_lod_deposit_ordered<<<(threads + 63u) / 64u, 64, 0, s>>>(a);
    return cudaGetLastError();
}

// Folds the private replicas of the finest LOD level (stream_collide.cuh, lod_deposit_warp) into QU_lod and clears them for
// the next step.  One warp per LOD entry: lane r reads (and zeroes) the entry's float4 in replica r, a shuffle tree adds the
// lanes, lane 0 adds the sum to QU_lod.  (A thread looping over the replicas took 31 us; this is one load per thread.)
__global__ void __launch_bounds__(256) k_lod_fold(const __grid_constant__ KArgs a, uint32_t own_offset) {
    const uint32_t entry = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (entry >= a.lod_rep_entries) return;  // whole warps leave together (blockDim is a multiple of 32)
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t r = lane; r <= a.lod_rep_mask; r += 32u) {
        float4* p = reinterpret_cast<float4*>(a.lod_rep) + (size_t)r * a.lod_rep_entries + entry;
        const float4 v = *p;
        *p = make_float4(0.f, 0.f, 0.f, 0.f);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s.x += __shfl_down_sync(0xffffffffu, s.x, off);
        s.y += __shfl_down_sync(0xffffffffu, s.y, off);
        s.z += __shfl_down_sync(0xffffffffu, s.z, off);
        s.w += __shfl_down_sync(0xffffffffu, s.w, off);
    }
    const uint32_t e = entry + own_offset;
    if (lane == 0u && e < a.n_lod) {
        float4* q = reinterpret_cast<float4*>(a.QU_lod) + e;
        float4 v = *q;
        v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
        *q = v;
    }
}
cudaError_t launch_lod_fold(const KArgs& a, cudaStream_t s) {
    uint32_t off = 0u;
    if (a.dx > 1u || a.dy > 1u || a.dz > 1u)
        for (uint32_t d = 0u; d < a.lod_depth; d++) off += 1u << (d * 3u);  // sim.cl:667-670 (to_d of the 3-D sets)
    const uint64_t threads = (uint64_t)a.lod_rep_entries * 32u;
    k_lod_fold<<<(unsigned)((threads + 255u) / 256u), 256, 0, s>>>(a, off);
    return cudaGetLastError();
}

// clear_qu_lod, sim.cl:995-1003: global size n_lod (domain.rs:277), guard n > NUM_LOD_OWN (quirk Q12)
__global__ void k_clear_qu_lod(float* __restrict__ QU_lod, uint32_t n_lod, uint32_t n_lod_own) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_lod || n > n_lod_own) return;
    reinterpret_cast<float4*>(QU_lod)[n] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// lod_part_2_gather, sim.cl:864-895: 8 children of the next finer level -> charge sum, velocity mean
__global__ void k_lod_gather(float* __restrict__ lods, uint32_t depth) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (1u << (depth * 3u))) return;
    const uint32_t nd = 1u << depth;
    const uint32_t t = n % (nd * nd);
    const uint32_t bx = (t % nd) * 2u, by = (t / nd) * 2u, bz = (n / (nd * nd)) * 2u;
    const uint32_t nnd = 1u << (depth + 1u);
    uint32_t off = 0u;
    for (uint32_t d = 0; d < depth; d++) off += 1u << (d * 3u);
    const uint32_t fine = off + (1u << (depth * 3u));
    // child visiting order of sim.cl:875-882 (the float sums depend on it)
    const uint32_t ox[8] = {0, 1, 1, 1, 1, 0, 0, 0}, oy[8] = {0, 0, 1, 1, 0, 1, 1, 0}, oz[8] = {0, 0, 0, 1, 1, 0, 1, 1};
    float qs = 0.0f, uxs = 0.0f, uys = 0.0f, uzs = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t j = fine + (bx + ox[i]) + (by + oy[i]) * nnd + (bz + oz[i]) * nnd * nnd;
        const float4 v = reinterpret_cast<const float4*>(lods)[j];
        qs += v.x; uxs += v.y; uys += v.z; uzs += v.w;
    }
    reinterpret_cast<float4*>(lods)[off + n] = make_float4(qs, uxs * 0.125f, uys * 0.125f, uzs * 0.125f);
}

// ---- launchers used by api.cu ----
// descriptors of the foreign domains in the order k_build_sources lays them out (ascending index, self skipped)
static void describe_foreign(const KArgs& a, uint32_t nd, ForeignSet& fs, int skip_domain = -1) {
    fs.n = 0;
    const uint32_t dxy = a.dx * a.dy, dn = dxy * a.dz;
    if (dn <= 1u || dn - 1u > (uint32_t)MAX_FOREIGN) return;  // single domain, or too many: the kernel walks the flat table
    const int cdx = (int)((a.di % dxy) % a.dx), cdy = (int)((a.di % dxy) / a.dx), cdz = (int)(a.di / dxy);
    uint32_t entry = a.n_lod_own, flat = 0;
    for (uint32_t d = 0; d < dn; d++) {
        if (d == a.di) continue;
        const int ddx = cdx - (int)((d % dxy) % a.dx), ddy = cdy - (int)((d % dxy) / a.dx), ddz = cdz - (int)(d / dxy);
        int dist = abs(ddx);
        if (abs(ddy) > dist) dist = abs(ddy);
        if (abs(ddz) > dist) dist = abs(ddz);
        const uint32_t level = (int)a.lod_depth - dist > 0 ? (uint32_t)((int)a.lod_depth - dist) : 0u;
        const uint32_t ndf = 1u << level, cnt = ndf * ndf * ndf;
        ForeignDesc& f = fs.d[fs.n++];
        f.entry0 = entry;
        f.flat0 = flat;
        f.count = cnt;
        f.level = level;
        f.sx = (float)((uint32_t)ddx * a.nx);  // sim.cl:970-972 incl. the uint wrap of negative differences (quirk Q18)
        f.sy = (float)((uint32_t)ddy * a.ny);
        f.sz = (float)((uint32_t)ddz * a.nz);
        f.fast = (level + 1u == a.lod_depth && 2u * ndf == nd && a.nx / ndf == 2u * (a.nx / nd) && a.ny >= ndf && a.nz >= ndf) ? 1u : 0u;
        if ((int)d == skip_domain) f.fast = 2u;  // summed by the polyphase FFT pass (eb_fft.cu, second source set)
        entry += cnt;
        flat += cnt;
    }
}

template <int ND, int NC, int BLOCK, bool VOL>
static cudaError_t launch_pair(const KArgs& a, const LodSource* src, uint32_t own, uint32_t count, cudaStream_t s, uint32_t mode = 0u, int skip_domain = -1) {
    const size_t smem = (mode & 1u) ? 0 : (size_t)own * 2 * sizeof(float4);
    if (smem > 48u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(k_update_e_b_pair<ND, NC, BLOCK, VOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * fine_bytes(ND)));
        if (e != cudaSuccess) return e;
    }
    ForeignSet fs;
    describe_foreign(a, ND, fs, skip_domain);
    const uint32_t threads_per_plane = (a.nx / ND) * (ND / NC) * a.ny;
    const dim3 grid((threads_per_plane + BLOCK - 1) / BLOCK, a.nz);
    k_update_e_b_pair<ND, NC, BLOCK, VOL><<<grid, BLOCK, smem, s>>>(a, src + own, count - own, fs, mode);
    return cudaGetLastError();
}

template <int ND, bool EXACT>
static cudaError_t launch_tiled(const KArgs& a, const LodSource* src, uint32_t own, uint32_t count, cudaStream_t s) {
    const size_t smem = (size_t)own * sizeof(float4);
    if (smem > 48u * 1024u) {  // per device, so set on every launch (a host-side table lookup)
        cudaError_t e = cudaFuncSetAttribute(k_update_e_b_tiled<ND, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_bytes(ND));
        if (e != cudaSuccess) return e;
    }
    const uint32_t dsx = a.nx / ND;
    const dim3 grid((dsx * a.ny + EBT_BLOCK - 1) / EBT_BLOCK, a.nz);
    k_update_e_b_tiled<ND, EXACT><<<grid, EBT_BLOCK, smem, s>>>(a, src + own, count - own);
    return cudaGetLastError();
}

// exact: reproduce the reference's arithmetic and summation order bit for bit (slower); see k_update_e_b_tiled
cudaError_t launch_update_e_b(const KArgs& a, void* scratch_sources, cudaStream_t s, uint64_t* launches, bool exact) {
    const uint32_t count = source_count(a.lod_depth, a.n_lod_own, a.dx, a.dy, a.dz, a.di);
    LodSource* src = reinterpret_cast<LodSource*>(scratch_sources);
    k_build_sources<<<(count + 255u) / 256u, 256, 0, s>>>(a, src, count);
    *launches += 2;
    const uint32_t nd = 1u << a.lod_depth;
    const uint32_t fine = nd * nd * nd;
    const uint32_t own = a.n_lod_own >= fine ? fine : a.n_lod_own;
    if ((a.lod_depth == 3u || a.lod_depth == 4u) && a.nx % nd == 0u && a.n_lod_own >= fine && a.ny >= nd && a.nz >= nd) {
        // Fast path.  Depth 4 (16^3 sources): circular packed-FFMA2 kernel, 16 cells per thread (26.2 ms at 256^3 vs 36.3 ms
        // for the scalar tiled kernel, whose unrolled tile overflows the instruction cache).  Depth 3 (8^3 sources): the
        // scalar tiled kernel -- its 8x8 tile is 13 KB of code and runs at 16 warps/SM (4.8 ms vs 6.1 ms packed).
        // ION_EB_VARIANT overrides for A/B timing: 0 scalar tiled, 1 = 8 cells/thread with the row table in registers,
        // 2 = 16 cells/thread + volatile LDS, 3 = 8 cells/thread + volatile LDS at 512 threads, 4 = 16 cells/thread, row table
        // in registers.
        if (exact) return a.lod_depth == 4u ? launch_tiled<16, true>(a, src, own, count, s) : launch_tiled<8, true>(a, src, own, count, s);
        static const int variant = getenv("ION_EB_VARIANT") ? atoi(getenv("ION_EB_VARIANT")) : -1;
        if (variant == 0 || (variant < 0 && a.lod_depth == 3u))
            return a.lod_depth == 4u ? launch_tiled<16, false>(a, src, own, count, s) : launch_tiled<8, false>(a, src, own, count, s);
        if (a.lod_depth == 4u) {
            if (variant == 1) return launch_pair<16, 8, 256, false>(a, src, own, count, s);
            if (variant == 3) return launch_pair<16, 8, 512, true>(a, src, own, count, s);
            if (variant == 4) return launch_pair<16, 16, 256, false>(a, src, own, count, s);
            if (variant == 5) return launch_pair<16, 16, 384, true>(a, src, own, count, s);
            if (variant == 6) return launch_pair<16, 8, 384, true>(a, src, own, count, s);
            return launch_pair<16, 16, 256, true>(a, src, own, count, s);
        }
        return variant == 2 || variant == 3 ? launch_pair<8, 8, 256, true>(a, src, own, count, s) : launch_pair<8, 8, 256, false>(a, src, own, count, s);
    }
    unsigned b = ((a.nx + 31u) / 32u) * 32u;
    if (b > (unsigned)EB_BLOCK) b = EB_BLOCK;
    const dim3 grid((a.nx + b - 1u) / b, a.ny, a.nz);
    if (exact) k_update_e_b<true><<<grid, b, 0, s>>>(a, src, count);
    else k_update_e_b<false><<<grid, b, 0, s>>>(a, src, count);
    return cudaGetLastError();
}
// Foreign-domain pyramids only (sim.cl:957-983), added to the E_dyn / B_dyn the polyphase FFT pass (eb_fft.cu) has written.
// `skip_domain`: a domain whose pyramid the FFT pass has summed already (-1 = none).
cudaError_t launch_update_e_b_foreign(const KArgs& a, void* scratch_sources, cudaStream_t s, uint64_t* launches, int skip_domain) {
    const uint32_t count = source_count(a.lod_depth, a.n_lod_own, a.dx, a.dy, a.dz, a.di);
    const uint32_t nd = 1u << a.lod_depth;
    const uint32_t own = nd * nd * nd;
    if (count <= own) return cudaSuccess;  // single domain
    {  // anything left to sum?  Domains whose centre shift wraps through the uint product (quirk Q18) are skipped by the fast path,
       // and `skip_domain` is already in E_dyn / B_dyn: the first slab of a z decomposition has nothing to do here
        ForeignSet fs;
        describe_foreign(a, nd, fs, skip_domain);
        bool any = fs.n == 0u;  // too many domains for descriptors: the kernel walks the flat table
        for (uint32_t i = 0; i < fs.n && !any; i++) {
            const ForeignDesc& f = fs.d[i];
            any = f.fast != 2u && !(fabsf(f.sx) > ION_FAR_SHIFT || fabsf(f.sy) > ION_FAR_SHIFT || fabsf(f.sz) > ION_FAR_SHIFT);
        }
        if (!any) return cudaSuccess;
    }
    LodSource* src = reinterpret_cast<LodSource*>(scratch_sources);
    k_build_sources<<<(count + 255u) / 256u, 256, 0, s>>>(a, src, count);
    *launches += 2;
    return a.lod_depth == 4u ? launch_pair<16, 16, 256, true>(a, src, own, count, s, 3u, skip_domain)
                             : launch_pair<8, 8, 256, false>(a, src, own, count, s, 3u, skip_domain);
}
size_t lod_source_bytes(uint32_t lod_depth, uint32_t n_lod_own, uint32_t dx, uint32_t dy, uint32_t dz, uint32_t di) {
    return (size_t)source_count(lod_depth, n_lod_own, dx, dy, dz, di) * sizeof(LodSource);
}
cudaError_t launch_clear_qu_lod(const KArgs& a, cudaStream_t s) {
    k_clear_qu_lod<<<(a.n_lod + 255u) / 256u, 256, 0, s>>>(a.QU_lod, a.n_lod, a.n_lod_own);
    return cudaGetLastError();
}
cudaError_t launch_lod_gather(const KArgs& a, cudaStream_t s, uint64_t* launches) {
    for (int i = (int)a.lod_depth - 1; i >= 0; i--) {  // domain.rs:456-459
        const uint32_t cnt = 1u << (i * 3);
        k_lod_gather<<<(cnt + 127u) / 128u, 128, 0, s>>>(a.QU_lod, (uint32_t)i);
        (*launches)++;
    }
    return cudaGetLastError();
}

}  // namespace ion
