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
#include "lattice.cuh"

namespace ion {

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// voxelize_mesh, sim.cl:1150-1231: one thread per column of the face perpendicular to `direction`
__global__ void k_voxelize(const __grid_constant__ KArgs a, const uint32_t direction, const uint8_t flag, const float* __restrict__ p0,
                           const float* __restrict__ p1, const float* __restrict__ p2, const uint32_t triangle_number, const float x0,
                           const float y0, const float z0, const float x1, const float y1, const float z1, const float mpc_x,
                           const float mpc_y, const float mpc_z, const int mhd) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t A = direction == 0u ? a.ny * a.nz : direction == 1u ? a.nx * a.nz : a.nx * a.ny;
    if (col >= A) return;
    const int nx = (int)a.nx, ny = (int)a.ny, nz = (int)a.nz;
    uint32_t cx, cy, cz;  // sim.cl:1163-1168
    if (direction == 0u) { cx = (uint32_t)clampi((int)x0 - a.ox, 0, nx - 1); cy = col % a.ny; cz = col / a.ny; }
    else if (direction == 1u) { cx = col / a.nz; cy = (uint32_t)clampi((int)y0 - a.oy, 0, ny - 1); cz = col % a.nz; }
    else { cx = col % a.nx; cy = col / a.nx; cz = (uint32_t)clampi((int)z0 - a.oz, 0, nz - 1); }
    // position(xyz)+offset, sim.cl:131-133,1169-1170
    const float3 offset = f3(0.5f * (float)(nx + 2 * a.ox) - 0.5f, 0.5f * (float)(ny + 2 * a.oy) - 0.5f, 0.5f * (float)(nz + 2 * a.oz) - 0.5f);
    const float3 pos = f3((float)cx + 0.5f - 0.5f * (float)a.nx, (float)cy + 0.5f - 0.5f * (float)a.ny, (float)cz + 0.5f - 0.5f * (float)a.nz);
    const float3 r_origin = f3(pos.x + offset.x, pos.y + offset.y, pos.z + offset.z);
    const float3 r_direction = f3((float)(direction == 0u), (float)(direction == 1u), (float)(direction == 2u));
    uint32_t intersections = 0u, intersections_check = 0u;
    uint16_t distances[64];
    const bool outside = direction == 0u ? (r_origin.y < y0 || r_origin.z < z0 || r_origin.y >= y1 || r_origin.z >= z1)
                         : direction == 1u ? (r_origin.x < x0 || r_origin.z < z0 || r_origin.x >= x1 || r_origin.z >= z1)
                                           : (r_origin.x < x0 || r_origin.y < y0 || r_origin.x >= x1 || r_origin.y >= y1);
    if (outside) return;
    for (uint32_t i = 0u; i < triangle_number; i++) {  // Moeller-Trumbore, sim.cl:1177-1192
        const uint32_t tx = 3u * i, ty = tx + 1u, tz = ty + 1u;
        const float3 p0i = f3(p0[tx], p0[ty], p0[tz]);
        const float3 p1i = f3(p1[tx], p1[ty], p1[tz]);
        const float3 p2i = f3(p2[tx], p2[ty], p2[tz]);
        const float3 u = sub3(p1i, p0i), v = sub3(p2i, p0i), w = sub3(r_origin, p0i), h = cross3(r_direction, v), q = cross3(w, u);
        const float f = 1.0f / dot3(u, h), s = f * dot3(w, h), t = f * dot3(r_direction, q), d = f * dot3(v, q);
        if (s >= 0.0f && s < 1.0f && t >= 0.0f && s + t < 1.0f) {
            if (d > 0.0f) {
                if (intersections < 64u && d < 65536.0f) distances[intersections] = (uint16_t)d;
                intersections++;
            } else {
                intersections_check++;
            }
        }
    }
    for (int i = 1; i < (int)intersections && i < 64; i++) {  // insertion sort, sim.cl:1194-1202
        const uint16_t t = distances[i];
        int j = i - 1;
        while (j >= 0 && distances[j] > t) {
            distances[j + 1] = distances[j];
            j--;
        }
        distances[j + 1] = t;
    }
    bool inside = (intersections % 2u) && (intersections_check % 2u);
    uint32_t intersection = intersections % 2u != intersections_check % 2u;
    const uint32_t h0 = direction == 0u ? cx : direction == 1u ? cy : cz;
    const uint32_t hmax = direction == 0u ? (uint32_t)clampi((int)x1 - a.ox, 0, nx)
                          : direction == 1u ? (uint32_t)clampi((int)y1 - a.oy, 0, ny)
                                            : (uint32_t)clampi((int)z1 - a.oz, 0, nz);
    const uint32_t hmesh = intersections ? h0 + (uint32_t)distances[min(intersections - 1u, 63u)] : 0u;
    for (uint32_t h = h0; h < hmax; h++) {  // sim.cl:1209-1230
        while (intersection < intersections && h > h0 + (uint32_t)distances[min(intersection, 63u)]) {
            inside = !inside;
            intersection++;
        }
        inside = inside && (intersection < intersections && h < hmesh);
        const uint64_t n = (direction == 0u ? h : cx) + ((direction == 1u ? h : cy) + (uint64_t)(direction == 2u ? h : cz) * a.ny) * a.nx;
        if (inside) {
            const uint8_t flagsn = (uint8_t)((a.flags[n] & (uint8_t)~ION_TYPE_BO) | flag);
            if (mhd) {  // scratch aliasing of quirk Q15: M / charge parked in B_dyn
                if (flag & ION_TYPE_M) {
                    a.B_dyn[n] = mpc_x;
                    a.B_dyn[a.N + n] = mpc_y;
                    a.B_dyn[2ull * a.N + n] = mpc_z;
                } else if ((flag & ION_TYPE_F) || (flag & ION_TYPE_C)) {
                    a.B_dyn[n] = mpc_x;
                }
            }
            a.flags[n] = flagsn;
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
    k_voxelize<<<(A + 63u) / 64u, 64, 0, s>>>(a, direction, flag, p0, p1, p2, tri, bb6[0], bb6[1], bb6[2], bb6[3], bb6[4], bb6[5], mx,
                                              my, mz, mhd);
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

cudaError_t launch_precompute_b(const KArgs& a, uint32_t* counts, void* table, cudaStream_t s) {
    FieldSource* src = reinterpret_cast<FieldSource*>(table);
    cudaError_t e = compact(a, ION_TYPE_M, counts, src, a.B_dyn, 1, 1.0f, s);  // sim.cl:1240-1244 (coordinates(i)+1)
    if (e != cudaSuccess) return e;
    const uint32_t nblocks = (uint32_t)((a.N + CP_BLOCK - 1) / CP_BLOCK);
    const uint64_t total = (uint64_t)(a.nx + 2u) * (a.ny + 2u) * (a.nz + 2u);
    k_psi<<<(unsigned)((total + SF_BLOCK - 1) / SF_BLOCK), SF_BLOCK, 0, s>>>(a, a.E_dyn, src, counts + nblocks);
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

}  // namespace ion
