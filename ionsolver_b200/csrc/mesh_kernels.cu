// mesh_kernels.cu -- STL voxeliser and static-field precompute (the kernels that run before the hot loop for
// scenes with magnets / charged electrodes).
//
// Behaviour: /root/reference/src/kernels/sim_kernels.cl ("sim.cl") voxelize_mesh :1150-1231,
// psi_from_mesh :1234-1251, nabla :1253-1263, static_b_from_mesh :1265-1276, static_e_from_mesh :1278-1300.
//
// B200 design of the O(N^2) precompute: the reference loops over ALL N cells per output and tests the flag inside
// the loop.  Here the source cells (magnet / charged) are first compacted, in ascending cell order, into a
// packed table {x,y,z | Mx,My,Mz} (k_compact_sources, warp-ballot prefix + one atomic per warp would scramble
// the order, so a deterministic two-pass count/scan/scatter is used); the field kernels then stream that table
// through shared memory.  Because the order of the float sums (ascending source index) and every arithmetic
// operation (IEEE sqrt/div, no contraction) equal the reference's, psi/B_stat/E_stat come out bit-identical.
#include <cufft.h>
#include <dlfcn.h>

#include <climits>
#include <cstring>

#include "lattice.cuh"

namespace ion {

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ------------------------------------------------------------------------------------------------------
// voxelize_mesh (behaviour: sim.cl:1150-1231; host side mesh.rs:281-343).  One thread per column of the face perpendicular
// to the ray direction, organised for the GPU rather than after the reference's loop:
//   * everything of the ray-triangle test that does not depend on the column -- the edges u = p1-p0, v = p2-p0, h = dir x v and
//     f = 1/(u.h) -- is computed ONCE per triangle by the block and kept in shared memory (64-byte records, broadcast reads);
//     a column only evaluates w = origin - p0, s = f (w.h), q = w x u, t = f q_dir, d = f (v.q).  The expressions and their
//     operation order are the reference's (Moeller-Trumbore), so s, t, d -- and with them the flags -- are bit-identical.
//   * the hit distances are not kept as a sorted list: a hit toggles bit floor(d) of a per-column bitmap in shared memory.
//     The reference's fill loop only ever asks "how many of the (first 64) hits lie strictly below this cell", and only for the
//     PARITY of that count; walking up the column, the running XOR of the bitmap is that parity.  Its error-correction rules
//     reduce to closed forms: start index k0 = [hits ahead and behind differ in parity] skips the nearest hit (toggles
//     = max(k0, cnt) - k0), and nothing is inside at or beyond the farthest stored hit (hmesh).  No sort, no per-thread array.
// ------------------------------------------------------------------------------------------------------
struct __align__(16) TriRecord {
    float p0x, p0y, p0z, f;
    float ux, uy, uz, hx;
    float vx, vy, vz, hy;
    float hz, pad0, pad1, pad2;
};
constexpr int VOX_BLOCK = 128;
constexpr int VOX_CHUNK = 256;  // triangle records per shared-memory stage

__global__ void __launch_bounds__(VOX_BLOCK) k_voxelize(const __grid_constant__ KArgs a, const uint32_t direction, const uint8_t flag,
                                                         const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                                                         const uint32_t triangle_number, const float x0, const float y0, const float z0, const float x1,
                                                         const float y1, const float z1, const float mpc_x, const float mpc_y, const float mpc_z,
                                                         const int mhd, const uint32_t words) {
    extern __shared__ __align__(16) unsigned char vox_smem[];
    TriRecord* rec = reinterpret_cast<TriRecord*>(vox_smem);
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(rec + VOX_CHUNK);  // [word][thread]
    const uint32_t col = blockIdx.x * VOX_BLOCK + threadIdx.x;
    const uint32_t A = direction == 0u ? a.ny * a.nz : direction == 1u ? a.nx * a.nz : a.nx * a.ny;
    const int nx = (int)a.nx, ny = (int)a.ny, nz = (int)a.nz;
    // the column's first cell: where the ray enters the mesh's bounding box (sim.cl:1163-1168)
    uint32_t cx = 0u, cy = 0u, cz = 0u;
    if (direction == 0u) { cx = (uint32_t)clampi((int)x0 - a.ox, 0, nx - 1); cy = col % a.ny; cz = col / a.ny; }
    else if (direction == 1u) { cx = col / a.nz; cy = (uint32_t)clampi((int)y0 - a.oy, 0, ny - 1); cz = col % a.nz; }
    else { cx = col % a.nx; cy = col / a.nx; cz = (uint32_t)clampi((int)z0 - a.oz, 0, nz - 1); }
    // ray origin in global lattice coordinates: position(xyz) + offset, sim.cl:131-133,1169-1170
    const float ox = ((float)cx + 0.5f - 0.5f * (float)a.nx) + (0.5f * (float)(nx + 2 * a.ox) - 0.5f);
    const float oy = ((float)cy + 0.5f - 0.5f * (float)a.ny) + (0.5f * (float)(ny + 2 * a.oy) - 0.5f);
    const float oz = ((float)cz + 0.5f - 0.5f * (float)a.nz) + (0.5f * (float)(nz + 2 * a.oz) - 0.5f);
    const float dirx = (float)(direction == 0u), diry = (float)(direction == 1u), dirz = (float)(direction == 2u);
    bool live = col < A;
    if (live)  // columns that miss the bounding box do nothing (sim.cl:1174-1176)
        live = direction == 0u ? !(oy < y0 || oz < z0 || oy >= y1 || oz >= z1)
               : direction == 1u ? !(ox < x0 || oz < z0 || ox >= x1 || oz >= z1)
                                 : !(ox < x0 || oy < y0 || ox >= x1 || oy >= y1);
    for (uint32_t w = 0; w < words; w++) bitmap[w * VOX_BLOCK + threadIdx.x] = 0u;
    uint32_t ahead = 0u, behind = 0u, dmin = 0xFFFFFFFFu, dmax = 0u;
    for (uint32_t base = 0u; base < triangle_number; base += VOX_CHUNK) {
        const uint32_t m = min((uint32_t)VOX_CHUNK, triangle_number - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += VOX_BLOCK) {  // per-triangle part of the test, once per block
            const uint32_t i = 3u * (base + k);
            const float ax = p0[i], ay = p0[i + 1u], az = p0[i + 2u];
            TriRecord r;
            r.p0x = ax; r.p0y = ay; r.p0z = az;
            r.ux = p1[i] - ax; r.uy = p1[i + 1u] - ay; r.uz = p1[i + 2u] - az;
            r.vx = p2[i] - ax; r.vy = p2[i + 1u] - ay; r.vz = p2[i + 2u] - az;
            r.hx = diry * r.vz - dirz * r.vy;  // h = dir x v
            r.hy = dirz * r.vx - dirx * r.vz;
            r.hz = dirx * r.vy - diry * r.vx;
            r.f = 1.0f / (r.ux * r.hx + r.uy * r.hy + r.uz * r.hz);
            r.pad0 = r.pad1 = r.pad2 = 0.0f;
            rec[k] = r;
        }
        __syncthreads();
        if (!live) continue;
        for (uint32_t k = 0u; k < m; k++) {
            const TriRecord& r = rec[k];
            const float wx = ox - r.p0x, wy = oy - r.p0y, wz = oz - r.p0z;
            const float sv = r.f * (wx * r.hx + wy * r.hy + wz * r.hz);
            const float qx = wy * r.uz - wz * r.uy, qy = wz * r.ux - wx * r.uz, qz = wx * r.uy - wy * r.ux;  // q = w x u
            const float tv = r.f * (dirx * qx + diry * qy + dirz * qz);
            const float dv = r.f * (r.vx * qx + r.vy * qy + r.vz * qz);
            if (sv >= 0.0f && sv < 1.0f && tv >= 0.0f && sv + tv < 1.0f) {  // the ray's line crosses the triangle
                if (dv > 0.0f) {
                    if (ahead < 64u && dv < 65536.0f) {  // the reference keeps the first 64 hits, as ushort
                        const uint32_t dist = (uint32_t)dv;
                        if (dist < 32u * words) bitmap[(dist >> 5) * VOX_BLOCK + threadIdx.x] ^= 1u << (dist & 31u);
                        dmin = min(dmin, dist);
                        dmax = max(dmax, dist);
                    }
                    ahead++;
                } else {
                    behind++;  // second ray, backwards: is the starting point really outside? (error correction)
                }
            }
        }
    }
    if (!live || ahead == 0u) return;
    const bool inside0 = (ahead & 1u) && (behind & 1u);
    const bool skip_nearest = (ahead & 1u) != (behind & 1u);  // start at hit 1 instead of hit 0
    const uint32_t h0 = direction == 0u ? cx : direction == 1u ? cy : cz;
    const uint32_t hmax = direction == 0u ? (uint32_t)clampi((int)x1 - a.ox, 0, nx)
                          : direction == 1u ? (uint32_t)clampi((int)y1 - a.oy, 0, ny)
                                            : (uint32_t)clampi((int)z1 - a.oz, 0, nz);
    const uint32_t hend = min(hmax, h0 + dmax);  // nothing is inside at or beyond the farthest stored hit
    const uint64_t stride = direction == 0u ? 1ull : direction == 1u ? (uint64_t)a.nx : (uint64_t)a.nx * a.ny;
    uint64_t n = (uint64_t)cx + ((uint64_t)cy + (uint64_t)cz * a.ny) * a.nx;
    bool below_odd = false;  // parity of the hits strictly below the current cell
    uint32_t word = 0u;
    for (uint32_t h = h0; h < hend; h++, n += stride) {
        const uint32_t m = h - h0;
        if (m > 0u) {
            const uint32_t b = m - 1u;
            if ((b & 31u) == 0u) word = bitmap[(b >> 5) * VOX_BLOCK + threadIdx.x];
            below_odd ^= (word >> (b & 31u)) & 1u;
        }
        const bool toggled = skip_nearest ? (m > dmin ? !below_odd : false) : below_odd;
        if (inside0 != toggled) {
            a.flags[n] = (uint8_t)((a.flags[n] & (uint8_t)~ION_TYPE_BO) | flag);
            if (mhd) {  // scratch aliasing of quirk Q15: M / charge parked in B_dyn until initialize
                if (flag & ION_TYPE_M) {
                    a.B_dyn[n] = mpc_x;
                    a.B_dyn[a.N + n] = mpc_y;
                    a.B_dyn[2ull * a.N + n] = mpc_z;
                } else if ((flag & ION_TYPE_F) || (flag & ION_TYPE_C)) {
                    a.B_dyn[n] = mpc_x;
                }
            }
        }
    }
}

// ---- deterministic, order-preserving compaction of source cells ----
struct __align__(16) FieldSource {
    float x, y, z, pad;  // cell coordinates (already +1 for the padded psi grid when used by psi_from_mesh)
    float mx, my, mz, pad2;
};
constexpr int CP_BLOCK = 256;

__global__ void k_count_sources(const uint8_t* __restrict__ flags, uint64_t N, uint8_t mask, uint32_t* __restrict__ block_counts) {
    const uint64_t n = (uint64_t)blockIdx.x * CP_BLOCK + threadIdx.x;
    const int hit = n < N && (flags[n] & mask) != 0;
    const int total = __syncthreads_count(hit);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)total;
}
// single-CTA exclusive scan over block counts (N/256 entries; 65k for 256^3 -- negligible)
__global__ void k_scan_counts(uint32_t* __restrict__ block_counts, uint32_t nblocks, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (nblocks + 1023u) / 1024u;
    const uint32_t b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += block_counts[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; i++) { const uint32_t v = s_part[i]; s_part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t b = b0; b < b1; b++) { const uint32_t v = block_counts[b]; block_counts[b] = run; run += v; }
}
__global__ void k_scatter_sources(const __grid_constant__ KArgs a, uint8_t mask, const uint32_t* __restrict__ block_offsets,
                                  FieldSource* __restrict__ out, const float* __restrict__ M, int vector_valued, float pad) {
    __shared__ uint32_t s_warp[CP_BLOCK / 32];
    const uint64_t n = (uint64_t)blockIdx.x * CP_BLOCK + threadIdx.x;
    const bool hit = n < a.N && (a.flags[n] & mask) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t base = block_offsets[blockIdx.x];
    for (unsigned wi = 0; wi < warp; wi++) base += s_warp[wi];
    if (hit) {
        const uint32_t idx = base + __popc(bal & ((1u << lane) - 1u));
        const uint32_t nxy = a.nx * a.ny;
        const uint32_t t = (uint32_t)(n % nxy);
        FieldSource s;
        s.x = (float)(t % a.nx + (uint32_t)pad);
        s.y = (float)(t / a.nx + (uint32_t)pad);
        s.z = (float)((uint32_t)(n / nxy) + (uint32_t)pad);
        s.pad = 0.f;
        s.mx = M[n];
        s.my = vector_valued ? M[a.N + n] : 0.f;
        s.mz = vector_valued ? M[2ull * a.N + n] : 0.f;
        s.pad2 = 0.f;
        out[idx] = s;
    }
}

constexpr int SF_BLOCK = 128;
constexpr int SF_CHUNK = 512;

// psi_from_mesh, sim.cl:1234-1251: psi on the (n+2)^3 padded grid, written into E_dyn (scratch, quirk Q15)
__global__ void __launch_bounds__(SF_BLOCK) k_psi(const __grid_constant__ KArgs a, float* __restrict__ psi,
                                                   const FieldSource* __restrict__ src, const uint32_t* __restrict__ count_p) {
    __shared__ float4 s_a[SF_CHUNK], s_b[SF_CHUNK];
    const uint32_t count = *count_p;
    const uint64_t total = (uint64_t)(a.nx + 2u) * (a.ny + 2u) * (a.nz + 2u);
    const uint64_t n = (uint64_t)blockIdx.x * SF_BLOCK + threadIdx.x;
    const uint32_t lx = a.nx + 2u, lxy = (a.nx + 2u) * (a.ny + 2u);
    const uint32_t t = (uint32_t)(n % lxy);
    const float cx = (float)(t % lx), cy = (float)(t / lx), cz = (float)(uint32_t)(n / lxy);
    float psic = 0.0f;
    for (uint32_t base = 0; base < count; base += SF_CHUNK) {
        const uint32_t m = min((uint32_t)SF_CHUNK, count - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += SF_BLOCK) {
            const float4* p = reinterpret_cast<const float4*>(src + base + k);
            s_a[k] = p[0];
            s_b[k] = p[1];
        }
        __syncthreads();
        for (uint32_t k = 0; k < m; k++) {
            const float4 c = s_a[k], mg = s_b[k];
            const float dx = cx - c.x, dy = cy - c.y, dz = cz - c.z;
            const float l = sqrtf(dx * dx + dy * dy + dz * dz);
            if (!(l == 0.0f)) psic += (dx * mg.x + dy * mg.y + dz * mg.z) / cb(l);
        }
    }
    if (n < total) psi[n] = (float)((double)psic / (4.0 * 3.14159265358979323846));  // `4.0f * M_PI` is a double product, sim.cl:1250
}

// Fast mode of psi_from_mesh (opt-in, ion_domain_set_precompute_mode): the same compacted-source sum, organised for FP32 issue
// instead of for bit-identity -- four consecutive outputs of a row per thread (dy, dz, dy^2 + dz^2 and the y/z part of M.r are shared
// by the four), r^-3 from one rsqrt instead of sqrt + cube + IEEE division, fused multiply-adds, packed FP32 for the two output
// pairs.  ~9 instructions per (output, source) pair instead of ~35.  Differs from the exact mode by rounding only (relative L2
// ~1e-6 in psi; tests/test_gpu_parity.py::test_fast_precompute_mode).
constexpr int PSI_OUT = 4;
__global__ void __launch_bounds__(SF_BLOCK) k_psi_fast(const __grid_constant__ KArgs a, float* __restrict__ psi, const FieldSource* __restrict__ src,
                                                        const uint32_t* __restrict__ count_p) {
    __shared__ float4 s_a[SF_CHUNK], s_b[SF_CHUNK];
    const uint32_t count = *count_p;
    const uint32_t lx = a.nx + 2u, ly = a.ny + 2u, lz = a.nz + 2u;
    const uint32_t x0 = (blockIdx.x * SF_BLOCK + threadIdx.x) * PSI_OUT, y = blockIdx.y, z = blockIdx.z;
    const float cy = (float)y, cz = (float)z;
    const float2 cxa = make_float2((float)x0, (float)(x0 + 1u)), cxb = make_float2((float)(x0 + 2u), (float)(x0 + 3u));
    float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
    for (uint32_t base = 0; base < count; base += SF_CHUNK) {
        const uint32_t m = min((uint32_t)SF_CHUNK, count - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += SF_BLOCK) {
            const float4* p = reinterpret_cast<const float4*>(src + base + k);
            s_a[k] = p[0];
            s_b[k] = p[1];
        }
        __syncthreads();
#pragma unroll 4
        for (uint32_t k = 0; k < m; k++) {
            const float4 c = s_a[k], mg = s_b[k];
            const float dy = cy - c.y, dz = cz - c.z;
            const float dyz2 = fmaf(dy, dy, dz * dz), myz = fmaf(dy, mg.y, dz * mg.z);
            const float2 ncx = make_float2(-c.x, -c.x), d2 = make_float2(dyz2, dyz2), m2 = make_float2(myz, myz), mx2 = make_float2(mg.x, mg.x);
            const float2 dxa = __fadd2_rn(cxa, ncx), dxb = __fadd2_rn(cxb, ncx);
            const float2 ra = __ffma2_rn(dxa, dxa, d2), rb = __ffma2_rn(dxb, dxb, d2);
            // r = 0 only for the output that sits on the source cell itself: the reference skips it (sim.cl:1245)
            const float2 ia = make_float2(ra.x > 0.f ? rsqrtf(ra.x) : 0.f, ra.y > 0.f ? rsqrtf(ra.y) : 0.f);
            const float2 ib = make_float2(rb.x > 0.f ? rsqrtf(rb.x) : 0.f, rb.y > 0.f ? rsqrtf(rb.y) : 0.f);
            const float2 i3a = __fmul2_rn(__fmul2_rn(ia, ia), ia), i3b = __fmul2_rn(__fmul2_rn(ib, ib), ib);
            pa = __ffma2_rn(__ffma2_rn(dxa, mx2, m2), i3a, pa);
            pb = __ffma2_rn(__ffma2_rn(dxb, mx2, m2), i3b, pb);
        }
    }
    if (y >= ly || z >= lz) return;
    const uint64_t row = ((uint64_t)y + (uint64_t)z * ly) * lx;
    const float inv4pi = 0.07957747154594767f;
    const float out[4] = {pa.x * inv4pi, pa.y * inv4pi, pb.x * inv4pi, pb.y * inv4pi};
#pragma unroll
    for (int i = 0; i < PSI_OUT; i++)
        if (x0 + (uint32_t)i < lx) psi[row + x0 + (uint32_t)i] = out[i];
}

// static_b_from_mesh, sim.cl:1265-1276
__global__ void k_static_b(const __grid_constant__ KArgs a, const float* __restrict__ psi) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= a.nx || is_halo(a, x, y, z)) return;
    const uint64_t n = x + (y + (uint64_t)z * a.ny) * a.nx;
    if ((a.flags[n] & ION_TYPE_S) == ION_TYPE_S) return;  // any flag with the solid bit (quirk Q1)
    const uint32_t l0 = a.nx + 2u, l1 = a.ny + 2u;
    const uint64_t m = (x + 1u) + ((y + 1u) + (uint64_t)(z + 1u) * l1) * l0;
    const uint64_t yo = l0, zo = (uint64_t)l0 * l1;
    const float nkmu0 = -a.kmu0;
    a.B_stat[n] += nkmu0 * ((psi[m + 1] - psi[m - 1]) / 2.0f);
    a.B_stat[a.N + n] += nkmu0 * ((psi[m + yo] - psi[m - yo]) / 2.0f);
    a.B_stat[2ull * a.N + n] += nkmu0 * ((psi[m + zo] - psi[m - zo]) / 2.0f);
}

// static_e_from_mesh, sim.cl:1278-1300 (charge per cell is read from the x-plane of B_dyn)
__global__ void __launch_bounds__(SF_BLOCK) k_static_e(const __grid_constant__ KArgs a, float* __restrict__ E,
                                                        const FieldSource* __restrict__ src, const uint32_t* __restrict__ count_p) {
    __shared__ float4 s_a[SF_CHUNK], s_b[SF_CHUNK];
    const uint32_t count = *count_p;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    bool active = x < a.nx && !is_halo(a, x, y, z);
    const uint64_t n = (active ? x : 0u) + (y + (uint64_t)z * a.ny) * a.nx;
    if (active) active = (a.flags[n] & ION_TYPE_S) != ION_TYPE_S;
    const float cx = (float)x, cy = (float)y, cz = (float)z;
    float ex = 0.f, ey = 0.f, ez = 0.f;
    for (uint32_t base = 0; base < count; base += SF_CHUNK) {
        const uint32_t m = min((uint32_t)SF_CHUNK, count - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) {
            const float4* p = reinterpret_cast<const float4*>(src + base + k);
            s_a[k] = p[0];
            s_b[k] = p[1];
        }
        __syncthreads();
        if (active) {
            for (uint32_t k = 0; k < m; k++) {
                const float4 c = s_a[k];
                const float charge = s_b[k].x;
                const float dx = cx - c.x, dy = cy - c.y, dz = cz - c.z;
                const float l = sqrtf(dx * dx + dy * dy + dz * dz);
                if (!(l == 0.0f)) {
                    const float l3 = cb(l);
                    ex += dx * charge / l3;
                    ey += dy * charge / l3;
                    ez += dz * charge / l3;
                }
            }
        }
    }
    if (!active) return;
    E[n] += ex * a.ke;
    E[a.N + n] += ey * a.ke;
    E[2ull * a.N + n] += ez * a.ke;
}

// ---- launchers ----
cudaError_t launch_voxelize(const KArgs& a, uint32_t direction, uint8_t flag, const float* p0, const float* p1, const float* p2,
                            uint32_t tri, const float* bb6, float mx, float my, float mz, int mhd, cudaStream_t s) {
    const uint32_t A = direction == 0u ? a.ny * a.nz : direction == 1u ? a.nx * a.nz : a.nx * a.ny;
    // bitmap length: the cells a column can fill, [h0, hmax) -- the bounding box along the ray, clipped to the domain
    const int n_dir = (int)(direction == 0u ? a.nx : direction == 1u ? a.ny : a.nz), o_dir = direction == 0u ? a.ox : direction == 1u ? a.oy : a.oz;
    const int lo = clampi((int)bb6[direction] - o_dir, 0, n_dir - 1), hi = clampi((int)bb6[3 + direction] - o_dir, 0, n_dir);
    const uint32_t words = (uint32_t)((hi > lo ? hi - lo : 0) + 31) / 32u + 1u;
    const size_t smem = sizeof(TriRecord) * VOX_CHUNK + (size_t)words * VOX_BLOCK * sizeof(uint32_t);
    if (smem > 200u * 1024u) return cudaErrorInvalidValue;  // a mesh more than ~90 000 cells deep along the ray
    cudaError_t e = cudaFuncSetAttribute(k_voxelize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_voxelize<<<(A + VOX_BLOCK - 1u) / VOX_BLOCK, VOX_BLOCK, smem, s>>>(a, direction, flag, p0, p1, p2, tri, bb6[0], bb6[1], bb6[2], bb6[3], bb6[4],
                                                                          bb6[5], mx, my, mz, mhd, words);
    return cudaGetLastError();
}

size_t compaction_scratch_bytes(uint64_t N) { return ((N + CP_BLOCK - 1) / CP_BLOCK + 2) * sizeof(uint32_t); }

// compacts cells whose flag has any bit of `mask` into `out` (capacity in entries); count lands in counts[nblocks]
static cudaError_t compact(const KArgs& a, uint8_t mask, uint32_t* counts, FieldSource* out, const float* M, int vector_valued,
                           float pad, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    k_count_sources<<<nblocks, CP_BLOCK, 0, s>>>(a.flags, a.N, mask, counts);
    k_scan_counts<<<1, 1024, 0, s>>>(counts, nblocks, counts + nblocks);
    k_scatter_sources<<<nblocks, CP_BLOCK, 0, s>>>(a, mask, counts, out, M, vector_valued, pad);
    return cudaGetLastError();
}

// returns the number of source cells via a blocking read (needed to size the table): counts only
cudaError_t count_sources(const KArgs& a, uint8_t mask, uint32_t* counts, uint32_t* host_total, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    k_count_sources<<<nblocks, CP_BLOCK, 0, s>>>(a.flags, a.N, mask, counts);
    k_scan_counts<<<1, 1024, 0, s>>>(counts, nblocks, counts + nblocks);
    cudaError_t e = cudaMemcpyAsync(host_total, counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}
size_t field_source_bytes(uint32_t count) { return (size_t)(count ? count : 1u) * sizeof(FieldSource); }

cudaError_t launch_precompute_b(const KArgs& a, uint32_t* counts, void* table, cudaStream_t s, int fast) {
    FieldSource* src = reinterpret_cast<FieldSource*>(table);
    cudaError_t e = compact(a, ION_TYPE_M, counts, src, a.B_dyn, 1, 1.0f, s);  // sim.cl:1240-1244 (coordinates(i)+1)
    if (e != cudaSuccess) return e;
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    const uint64_t total = (uint64_t)(a.nx + 2u) * (a.ny + 2u) * (a.nz + 2u);
    if (fast) {
        const uint32_t per_block = SF_BLOCK * PSI_OUT;
        k_psi_fast<<<dim3((a.nx + 2u + per_block - 1u) / per_block, a.ny + 2u, a.nz + 2u), SF_BLOCK, 0, s>>>(a, a.E_dyn, src, counts + nblocks);
    } else {
        k_psi<<<(unsigned)((total + SF_BLOCK - 1) / SF_BLOCK), SF_BLOCK, 0, s>>>(a, a.E_dyn, src, counts + nblocks);
    }
    unsigned b = ((a.nx + 31u) / 32u) * 32u;
    if (b > 128u) b = 128u;
    k_static_b<<<dim3((a.nx + b - 1u) / b, a.ny, a.nz), b, 0, s>>>(a, a.E_dyn);
    return cudaGetLastError();
}
cudaError_t launch_precompute_e(const KArgs& a, float* E, uint32_t* counts, void* table, cudaStream_t s) {
    FieldSource* src = reinterpret_cast<FieldSource*>(table);
    cudaError_t e = compact(a, ION_TYPE_F | ION_TYPE_C, counts, src, a.B_dyn, 0, 0.0f, s);  // sim.cl:1287-1291
    if (e != cudaSuccess) return e;
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    unsigned b = ((a.nx + 31u) / 32u) * 32u;
    if (b > (unsigned)SF_BLOCK) b = SF_BLOCK;
    k_static_e<<<dim3((a.nx + b - 1u) / b, a.ny, a.nz), b, 0, s>>>(a, E, src, counts + nblocks);
    return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------------------------------------
// Mode 2 of the static-field precompute (opt-in, ion_domain_set_precompute_mode): the sums of psi_from_mesh (sim.cl:1234-1251) and
// static_e_from_mesh (sim.cl:1278-1300) are convolutions of the source cells with G(d) = d / |d|^3 on the cell lattice,
//     psi(x) = 1/(4 pi) sum_s M_s . G(x - c_s),        E(x) = ke sum_s q_s G(x - c_s),
// so they cost O(P^3 log P) through a zero-padded 3-D FFT instead of O(cells x sources): cfg3's 5.4 M magnet cells x 34 M outputs are
// 1.8e14 pair terms for the direct kernels.  P_a = (outputs + extent of the sources' bounding box - 1) per axis, rounded up to a
// 7-smooth length: with K[j mod P] = G(j - min) for j in [-(S-1), L-1] the circular convolution of the box-relative sources with K
// has no wrap-around on the L outputs.  The transforms are plain library FFTs (cuFFT R2C / C2R, bound at run time so that the library
// loads without it) off the time-step path; source scatter, kernel sampling, spectral products and the write-back are kernels here.
// Everything between the FP32 inputs and the FP32 result is double precision (D2Z / Z2D transforms, kernel samples, products): FP32
// transforms leave white noise of ~1e-6 of the LARGEST |psi| everywhere, which the central differences of static_b_from_mesh turn
// into 1e-3 of B_stat; in FP64 the mode returns the correctly rounded sum, i.e. it differs from the reference-order FP32 sum by that
// sum's own rounding only.  Opt-in because the result is not bit-identical to the reference's.
// ------------------------------------------------------------------------------------------------------------------------------
struct CufftApi {
    bool ok;
    cufftResult (*Plan3d)(cufftHandle*, int, int, int, cufftType);
    cufftResult (*SetStream)(cufftHandle, cudaStream_t);
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal*, cufftDoubleComplex*);
    cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex*, cufftDoubleReal*);
    cufftResult (*Destroy)(cufftHandle);
};
static const CufftApi& cufft_api() {
    static const CufftApi api = [] {
        CufftApi a;
        memset(&a, 0, sizeof(a));
        const char* names[] = {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so"};
        void* h = nullptr;
        for (const char* n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
        if (!h) return a;
        *(void**)(&a.Plan3d) = dlsym(h, "cufftPlan3d");
        *(void**)(&a.SetStream) = dlsym(h, "cufftSetStream");
        *(void**)(&a.ExecD2Z) = dlsym(h, "cufftExecD2Z");
        *(void**)(&a.ExecZ2D) = dlsym(h, "cufftExecZ2D");
        *(void**)(&a.Destroy) = dlsym(h, "cufftDestroy");
        a.ok = a.Plan3d && a.SetStream && a.ExecD2Z && a.ExecZ2D && a.Destroy;
        return a;
    }();
    return api;
}

struct ConvGeom {
    int mn[3];       // smallest source coordinate per axis
    int S[3];        // extent of the sources' bounding box
    int L[3];        // outputs per axis: 0 .. L-1, same frame as the source coordinates
    int P[3];        // transform lengths
};

__global__ void k_conv_bbox(const FieldSource* __restrict__ src, const uint32_t* __restrict__ count_p, int* __restrict__ box) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count_p) return;
    const FieldSource s = src[k];
    atomicMin(box + 0, (int)s.x); atomicMin(box + 1, (int)s.y); atomicMin(box + 2, (int)s.z);
    atomicMax(box + 3, (int)s.x); atomicMax(box + 4, (int)s.y); atomicMax(box + 5, (int)s.z);
}
// component `comp` (0..2: mx, my, mz; the charge of static_e sits in mx) of every source into the zeroed real array
__global__ void k_conv_scatter(const FieldSource* __restrict__ src, const uint32_t* __restrict__ count_p, const int comp, const __grid_constant__ ConvGeom g,
                               double* __restrict__ R) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count_p) return;
    const FieldSource s = src[k];
    const size_t i = ((size_t)((int)s.z - g.mn[2]) * g.P[1] + (size_t)((int)s.y - g.mn[1])) * g.P[0] + (size_t)((int)s.x - g.mn[0]);
    R[i] = (double)(comp == 0 ? s.mx : comp == 1 ? s.my : s.mz);
}
// K[j mod P] = G_comp(j - min) for j in [-(S-1), L-1], zero elsewhere and at d = 0 (the sums skip l == 0, sim.cl:1245, :1293)
__global__ void __launch_bounds__(128) k_conv_kernel(const int comp, const __grid_constant__ ConvGeom g, double* __restrict__ R) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y, iz = blockIdx.z;
    if (ix >= g.P[0]) return;
    const int jx = ix < g.L[0] ? ix : ix - g.P[0], jy = iy < g.L[1] ? iy : iy - g.P[1], jz = iz < g.L[2] ? iz : iz - g.P[2];
    double v = 0.0;
    if (jx > -g.S[0] && jy > -g.S[1] && jz > -g.S[2]) {
        const double dx = (double)(jx - g.mn[0]), dy = (double)(jy - g.mn[1]), dz = (double)(jz - g.mn[2]);
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > 0.0) v = (comp == 0 ? dx : comp == 1 ? dy : dz) / (r2 * sqrt(r2));
    }
    R[((size_t)iz * g.P[1] + iy) * g.P[0] + ix] = v;
}
// out = (accumulate ? out : 0) + A * B
__global__ void k_conv_mac(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ out, const size_t n, const int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = A[i], b = B[i];
        double2 r = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
        if (accumulate) { const double2 o = out[i]; r.x += o.x; r.y += o.y; }
        out[i] = r;
    }
}
__global__ void k_conv_psi_out(const double* __restrict__ R, const __grid_constant__ ConvGeom g, const double scale, float* __restrict__ psi) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= g.L[0]) return;
    psi[((size_t)z * g.L[1] + y) * g.L[0] + x] = (float)(R[((size_t)z * g.P[1] + y) * g.P[0] + x] * scale);
}
// E[comp] += ke * sum for every fluid cell that is not a halo cell (sim.cl:1281-1284, :1297-1299)
__global__ void k_conv_e_out(const __grid_constant__ KArgs a, const double* __restrict__ R, const __grid_constant__ ConvGeom g, const int comp, const double scale,
                             float* __restrict__ E) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= a.nx || is_halo(a, x, y, z)) return;
    const uint64_t n = x + (y + (uint64_t)z * a.ny) * a.nx;
    if ((a.flags[n] & ION_TYPE_S) == ION_TYPE_S) return;
    E[(uint64_t)comp * a.N + n] += (float)(R[((size_t)z * g.P[1] + y) * g.P[0] + x] * scale) * a.ke;  // the rounded sum, then * ke like sim.cl:1297
}

static int smooth7(int n) {  // smallest length >= n whose prime factors are 2, 3, 5, 7
    for (;; n++) {
        int m = n;
        for (int p : {2, 3, 5, 7})
            while (m % p == 0) m /= p;
        if (m == 1) return n;
    }
}

// which = 0: psi on the padded (n+2)^3 grid into a.E_dyn followed by static_b; which != 0: static E into `E`.
// Returns cudaErrorNotSupported with *why set when cuFFT or the memory for the transforms is not available.
cudaError_t launch_precompute_fft(const KArgs& a, int which, float* E, uint32_t* counts, void* table, uint32_t total, cudaStream_t s, uint64_t* launches,
                                  const char** why) {
    *why = nullptr;
    const CufftApi& fft = cufft_api();
    if (!fft.ok) { *why = "cuFFT could not be loaded (libcufft.so.11)"; return cudaErrorNotSupported; }
    FieldSource* src = reinterpret_cast<FieldSource*>(table);
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    const uint32_t* count_p = counts + nblocks;
    cudaError_t e = which == 0 ? compact(a, ION_TYPE_M, counts, src, a.B_dyn, 1, 1.0f, s) : compact(a, ION_TYPE_F | ION_TYPE_C, counts, src, a.B_dyn, 0, 0.0f, s);
    if (e != cudaSuccess) return e;
    *launches += 3;
    ConvGeom g;
    const int pad = which == 0 ? 2 : 0;
    g.L[0] = (int)a.nx + pad; g.L[1] = (int)a.ny + pad; g.L[2] = (int)a.nz + pad;
    int* box = nullptr;
    int hbox[6] = {INT_MAX, INT_MAX, INT_MAX, -1, -1, -1};
    e = cudaMallocAsync((void**)&box, sizeof(hbox), s);
    if (e != cudaSuccess) return e;
    cudaMemcpyAsync(box, hbox, sizeof(hbox), cudaMemcpyHostToDevice, s);
    k_conv_bbox<<<(total + 255u) / 256u, 256, 0, s>>>(src, count_p, box);
    cudaMemcpyAsync(hbox, box, sizeof(hbox), cudaMemcpyDeviceToHost, s);
    cudaFreeAsync(box, s);
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    *launches += 1;
    size_t real_n = 1, cplx_n = 1;
    for (int i = 0; i < 3; i++) {
        g.mn[i] = hbox[i];
        g.S[i] = hbox[3 + i] - hbox[i] + 1;
        g.P[i] = smooth7(g.L[i] + g.S[i] - 1);
        real_n *= (size_t)g.P[i];
        cplx_n *= (size_t)(i == 0 ? g.P[0] / 2 + 1 : g.P[i]);
    }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = real_n * sizeof(double) + 3 * cplx_n * sizeof(double2);
    if (need * 2 > free_b) { *why = "not enough free device memory for the transforms of the FFT precompute mode"; return cudaErrorNotSupported; }
    double* R = nullptr;
    double2 *A = nullptr, *B = nullptr, *C = nullptr;
    cufftHandle fwd = 0, inv = 0;
    bool have_fwd = false, have_inv = false;
    auto cleanup = [&]() {
        if (have_fwd) fft.Destroy(fwd);
        if (have_inv) fft.Destroy(inv);
        cudaFree(R); cudaFree(A); cudaFree(B); cudaFree(C);
    };
    e = cudaMalloc((void**)&R, real_n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&A, cplx_n * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&B, cplx_n * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&C, cplx_n * sizeof(double2));
    if (e != cudaSuccess) { cleanup(); return e; }
    // cuFFT: the LAST dimension varies fastest
    have_fwd = fft.Plan3d(&fwd, g.P[2], g.P[1], g.P[0], CUFFT_D2Z) == CUFFT_SUCCESS;
    have_inv = have_fwd && fft.Plan3d(&inv, g.P[2], g.P[1], g.P[0], CUFFT_Z2D) == CUFFT_SUCCESS;
    if (!have_fwd || !have_inv || fft.SetStream(fwd, s) != CUFFT_SUCCESS || fft.SetStream(inv, s) != CUFFT_SUCCESS) {
        cleanup();
        *why = "cufftPlan3d failed (transform lengths or work-area memory)";
        return cudaErrorNotSupported;
    }
    const dim3 gk((unsigned)(g.P[0] + 127) / 128, (unsigned)g.P[1], (unsigned)g.P[2]);
    const unsigned mac_blocks = (unsigned)((cplx_n + 255) / 256 < 148 * 16 ? (cplx_n + 255) / 256 : 148 * 16);
    const double norm = 1.0 / ((double)g.P[0] * (double)g.P[1] * (double)g.P[2]);
    bool ok = true;
    auto forward_source = [&](int comp) {
        cudaMemsetAsync(R, 0, real_n * sizeof(double), s);
        k_conv_scatter<<<(total + 255u) / 256u, 256, 0, s>>>(src, count_p, comp, g, R);
        ok = ok && fft.ExecD2Z(fwd, R, reinterpret_cast<cufftDoubleComplex*>(A)) == CUFFT_SUCCESS;
    };
    auto forward_kernel = [&](int comp) {
        k_conv_kernel<<<gk, 128, 0, s>>>(comp, g, R);
        ok = ok && fft.ExecD2Z(fwd, R, reinterpret_cast<cufftDoubleComplex*>(B)) == CUFFT_SUCCESS;
    };
    if (which == 0) {
        for (int c = 0; c < 3; c++) {
            forward_source(c);
            forward_kernel(c);
            k_conv_mac<<<mac_blocks, 256, 0, s>>>(A, B, C, cplx_n, c > 0 ? 1 : 0);
        }
        ok = ok && fft.ExecZ2D(inv, reinterpret_cast<cufftDoubleComplex*>(C), R) == CUFFT_SUCCESS;
        k_conv_psi_out<<<dim3((unsigned)(g.L[0] + 127) / 128, (unsigned)g.L[1], (unsigned)g.L[2]), 128, 0, s>>>(R, g, norm / (4.0 * 3.14159265358979323846), a.E_dyn);
        unsigned b = ((a.nx + 31u) / 32u) * 32u;
        if (b > 128u) b = 128u;
        k_static_b<<<dim3((a.nx + b - 1u) / b, a.ny, a.nz), b, 0, s>>>(a, a.E_dyn);
        *launches += 11;  // own kernels; the transforms are library launches
    } else {
        forward_source(0);
        for (int c = 0; c < 3; c++) {
            forward_kernel(c);
            k_conv_mac<<<mac_blocks, 256, 0, s>>>(A, B, C, cplx_n, 0);
            ok = ok && fft.ExecZ2D(inv, reinterpret_cast<cufftDoubleComplex*>(C), R) == CUFFT_SUCCESS;
            k_conv_e_out<<<dim3((a.nx + 127u) / 128u, a.ny, a.nz), 128, 0, s>>>(a, R, g, c, norm, E);
        }
        *launches += 10;
    }
    e = cudaGetLastError();
    const cudaError_t e2 = cudaStreamSynchronize(s);  // the work arrays are freed below
    cleanup();
    if (!ok) { *why = "a cuFFT transform failed"; return cudaErrorNotSupported; }
    return e != cudaSuccess ? e : e2;
}

}  // namespace ion
