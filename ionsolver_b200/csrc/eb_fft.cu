// eb_fft.cu -- update_e_b_dynamic (sim_kernels.cl:897-993) as a polyphase FFT convolution: kernels, plan, launcher.
// The algorithm, its layouts and the phase functions are in eb_fft_core.cuh; this file only wraps them into kernels.
//
//   k_eb_khat  (once per geometry)  K^_o for every task (in-block offset, z window), 3 components       -> HBM, static
//   k_eb_src   (every step)         s^_j of the LOD source table (q, q v), H x 4 blocks                  -> L2 resident
//   k_eb_fft   (every step)         one block per task: products, 2-D inverse FFT in shared memory, Hermitian DFT along x,
//                                   E_dyn / B_dyn = static + k * sum   (sim.cl:986-992)
#include <cstdlib>
#include <cstring>
#include <vector>

#include "eb_fft_core.cuh"
#include "eb_fft_layout.hpp"
#include "lattice.cuh"

namespace ion {
using namespace ebfft;

// one far slab's pyramid level: where it sits in QU_lod and how its centres are placed (sim.cl:960-972)
struct FarDesc {
    uint32_t entry0, nf;     // first QU_lod entry, sources per axis (2^level)
    float bsx, bsy, bsz;     // block size of that level: N / 2^level (integer division, lod_coordinates sim.cl:440-447)
    float shift_z;           // ddz * nz, halo layers included (quirk Q8)
};
struct FarDescs {
    int n;
    FarDesc d[8];
};
struct EbFftPlan {
    int nd;
    uint32_t ntasks;
    Geom g;
    // source sets summed by FFT: [0] the own finest level, [1] (z slabs with a lower neighbour) that neighbour's level D-1 pyramid
    int nsets;
    SourceSet sets[2];
    int foreign_domain;  // domain index whose pyramid is set 1, -1 = none (the direct kernel must then sum it)
    Task* tasks;
    // Kernel spectra.  Static mode (batch == ntasks): all of them, computed once.  Streamed mode (they do not fit the memory budget,
    // e.g. cfg5's 110 GB): a buffer for `batch` tasks per set, refilled by k_eb_khat in front of every k_eb_fft launch of a step
    // -- the field update then costs one forward and one inverse transform per task instead of an inverse one, and no memory.
    // Mirrored mode (whenever the full set does not fit, or ION_EB_FFT_MIRROR=1): a task with 2 ox > dsx reads the spectra of its
    // partner with x offset dsx - ox at the reflected in-plane index (eb_fft_core.cuh, main_phase_product_mirror), so only the
    // ~(dsx / 2 + 1) / dsx "canonical" tasks own a slot.  `batches`: the task array is laid out batch by batch, canonical tasks first
    // (their position inside the batch = their slot), then the mirrored ones that read those slots.
    uint32_t batch;  // slots of the buffer, per set
    bool streamed;
    struct Batch { uint32_t c0, nc, m0, nm; };
    std::vector<Batch> batches;
    float2* khat;   // nsets * batch * khat_per_task
    float2* shat;   // nsets * shat_count
    float2* shatc;  // compact spectrum of set 1 (odd-position symmetry), shatc_count
    float* scratch;  // the sums of a step before they are combined with the static fields: 6 floats per cell, rows permuted
    size_t khat_bytes;
    // far slabs (two and more below): summed as a second-order Taylor polynomial per FARB^3 block of cells (eb_fft_core.cuh)
    int far_n;               // number of far sources (0 = none); far_handled says whether this plan sums them
    bool far_handled;
    FarDesc far_desc[8];
    int far_ndesc;
    FarSource* far_src;
    float* far_tens;         // [FAR_T][far_blocks]
    uint32_t far_shift;      // log2 of the Taylor block edge: 3 (8^3) or 2 (4^3), chosen from the smallest source distance
    uint32_t far_nbx, far_nby, far_nbz;
};

template <int ND>
__global__ void __launch_bounds__(Cfg<ND>::KT) k_eb_khat(const __grid_constant__ Geom g, const __grid_constant__ SourceSet ss, const Task* __restrict__ tasks,
                                                  float2* __restrict__ khat) {
    extern __shared__ __align__(128) unsigned char eb_smem[];
    float2* S = reinterpret_cast<float2*>(eb_smem);
    const int task = blockIdx.x, comp = blockIdx.y;
    const Task t = tasks[task];
    khat_phase_x<ND>(threadIdx.x, blockDim.x, g, ss, t, comp, S);
    __syncthreads();
    khat_phase_y<ND>(threadIdx.x, blockDim.x, S);
    __syncthreads();
    khat_phase_z<ND>(threadIdx.x, blockDim.x, task, comp, S, khat);
}

template <int ND>
__global__ void __launch_bounds__(128) k_eb_src(const __grid_constant__ Geom g, const __grid_constant__ SourceSet ss, const float* __restrict__ QU_lod,
                                                 float2* __restrict__ shat, float2* __restrict__ shatc) {
    __shared__ float2 plane[Cfg<ND>::M * Cfg<ND>::ROW];
    const int kx = blockIdx.x, j = blockIdx.y;
    src_phase_x<ND>(threadIdx.x, blockDim.x, g, ss, QU_lod, kx, j, plane);
    __syncthreads();
    src_phase_y<ND>(threadIdx.x, blockDim.x, plane);
    __syncthreads();
    src_phase_z<ND>(threadIdx.x, blockDim.x, kx, j, plane, shat, shatc);
}

// ---- mbarrier + 1-D bulk copy (TMA engine, no registers, no issue slots per byte) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One block per task.  The K^ slots of the next group of kx planes stream into the idle half of the work buffer (bulk copies
// armed on an mbarrier by one thread) while the block transforms the current group: the HBM latency of the only large operand
// is hidden behind the FFT phases, and the products are formed in place where the copy landed.
// NSETS = 2 (a slab with a lower neighbour): K^ of the neighbour's level D-1 pyramid is staged into the B slots instead of
// s^_1..3, both source spectra are read from L2 in the product phase, and the two convolutions share every inverse transform.
template <int ND, int NSETS, int MIRROR>
__global__ void __launch_bounds__(Cfg<ND>::T, 1)
    k_eb_fft(const __grid_constant__ Geom g, const Task* __restrict__ tasks, const float2* __restrict__ khat, const float2* __restrict__ shat,
             const float2* __restrict__ khat2, const float2* __restrict__ shat2c, float* __restrict__ scratch, const int accumulate) {
    typedef Cfg<ND> C;
    extern __shared__ __align__(128) unsigned char eb_smem[];
    __shared__ __align__(8) uint64_t bars[2];
    float2* W = reinterpret_cast<float2*>(eb_smem);  // [2][P][6][M][ROW]
    float2* S0 = W + (size_t)2 * C::P * C::PLANE;     // s^_0 of the current planes [P][M][ROW] (single buffer, see below);
                                                      // NSETS = 2: the neighbour's compact spectrum [P][4][ND*ND] instead
    float4* tw = reinterpret_cast<float4*>(S0 + (size_t)C::P * C::SLOT);  // [H][ND/2], see main_phase_accumulate
    const int tid = threadIdx.x;
    if (tid < C::H * (ND / 2)) tw[tid] = main_tw4<ND>(tid / (ND / 2), tid % (ND / 2));
    const Task t = tasks[blockIdx.x];
    const float2* kt = khat + (size_t)t.kslot * C::khat_per_task;  // MIRROR: the partner's slot
    const float2* kt2 = NSETS == 2 ? khat2 + (size_t)t.kslot * C::khat_per_task : nullptr;
    if (tid == 0) {
        mbar_init(&bars[0], 1u);
        mbar_init(&bars[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: operands of iteration `it` -> buffer it & 1: K^ into the E slots, s^_1..3 into the B slots, s^_0 into S0
    auto stage = [&](int it) {
        const int kx0 = it * C::P, np = C::H - kx0 < C::P ? C::H - kx0 : C::P;
        float2* Wb = W + (size_t)(it & 1) * C::P * C::PLANE;
        uint64_t* bar = &bars[it & 1];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the targets were last touched by ordinary loads / stores
        mbar_expect_tx(bar, (uint32_t)(np * ((NSETS == 2 ? 6 : 7) * C::SLOT + (NSETS == 2 ? 4 * C::CSLOT : 0)) * sizeof(float2)));
        for (int p = 0; p < np; p++) {
            if (NSETS == 1) bulk_g2s(S0 + (size_t)p * C::SLOT, shat + ((size_t)kx0 + p) * C::SLOT, (uint32_t)(C::SLOT * sizeof(float2)), bar);
            if (NSETS == 2) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    bulk_g2s(S0 + ((size_t)p * 4 + j) * C::CSLOT, shat2c + ((size_t)j * C::H + kx0 + p) * C::CSLOT, (uint32_t)(C::CSLOT * sizeof(float2)), bar);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                bulk_g2s(Wb + (size_t)p * C::PLANE + (size_t)c * C::SLOT, kt + ((size_t)c * C::H + kx0 + p) * C::SLOT,
                         (uint32_t)(C::SLOT * sizeof(float2)), bar);
                const float2* second = NSETS == 2 ? kt2 + ((size_t)c * C::H + kx0 + p) * C::SLOT : shat + ((size_t)(1 + c) * C::H + kx0 + p) * C::SLOT;
                bulk_g2s(Wb + (size_t)p * C::PLANE + (size_t)(3 + c) * C::SLOT, second, (uint32_t)(C::SLOT * sizeof(float2)), bar);
            }
        }
    };
    if (tid == 0) stage(0);
    float acc[6][C::XPT];
#pragma unroll
    for (int s = 0; s < 6; s++)
#pragma unroll
        for (int i = 0; i < C::XPT; i++) acc[s][i] = 0.0f;
    for (int it = 0; it < C::NIT; it++) {
        const int kx0 = it * C::P, np = C::H - kx0 < C::P ? C::H - kx0 : C::P;
        float2* Wb = W + (size_t)(it & 1) * C::P * C::PLANE;
        mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
        if (NSETS == 2) {
            if (MIRROR) main_phase_product2_mirror<ND>(tid, shat, S0, tw, kx0, np, Wb);
            else main_phase_product2<ND>(tid, shat, S0, kx0, np, Wb);
        } else {
            if (MIRROR) main_phase_product_mirror<ND>(tid, S0, np, Wb);
            else main_phase_product<ND>(tid, S0, np, Wb);
        }
        __syncthreads();
        // Every thread is past the x accumulation of iteration it - 1 (which read the other buffer) and past this iteration's
        // products (which read S0): both are free, the operands of iteration it + 1 can land while this one is transformed.
        if (tid == 0 && it + 1 < C::NIT) stage(it + 1);
        main_phase_z<ND>(tid, np, Wb);
        __syncthreads();
        main_phase_y<ND>(tid, np, Wb);
        __syncthreads();
        main_phase_accumulate<ND>(tid, kx0, np, Wb, tw, acc);  // no barrier: the next iteration works on the other buffer
    }
    main_phase_store<ND>(tid, g, t, scratch, accumulate != 0, acc);
}

// far sources of this step: (centre, q, q v) of every entry of the far slabs' pyramid levels, in the reference's order
__global__ void k_eb_far_sources(const __grid_constant__ FarDescs fd, const float* __restrict__ QU_lod, FarSource* __restrict__ out) {
    uint32_t base = 0;
    for (int i = 0; i < fd.n; i++) {
        const FarDesc& d = fd.d[i];
        const uint32_t cnt = d.nf * d.nf * d.nf;
        for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) {
            const uint32_t t = k % (d.nf * d.nf);
            const float4 v = reinterpret_cast<const float4*>(QU_lod)[d.entry0 + k];
            FarSource s;
            s.cx = (float)(t % d.nf) * d.bsx + 0.5f * d.bsx;
            s.cy = (float)(t / d.nf) * d.bsy + 0.5f * d.bsy;
            s.cz = ((float)(k / (d.nf * d.nf)) * d.bsz + 0.5f * d.bsz) - d.shift_z;
            s.q = v.x;
            s.wx = v.y * v.x; s.wy = v.z * v.x; s.wz = v.w * v.x;
            s.pad = 0.0f;
            out[base + k] = s;
        }
        base += cnt;
    }
}
// Taylor tensors of the far field, one thread per FARB^3 block of cells
__global__ void __launch_bounds__(128) k_eb_far_tensors(const FarSource* __restrict__ src, const int nsrc, float* __restrict__ tens, const uint32_t nbx,
                                                         const uint32_t nby, const uint32_t nbz, const uint32_t FARB) {
    __shared__ FarSource s_src[128];
    for (int k = threadIdx.x; k < nsrc; k += blockDim.x) s_src[k] = src[k];
    __syncthreads();
    const uint32_t nblocks = nbx * nby * nbz;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint32_t bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
    const float c = 0.5f * (float)(FARB - 1u);
    float t[FAR_T];
#pragma unroll
    for (int i = 0; i < FAR_T; i++) t[i] = 0.0f;
    for (int k = 0; k < nsrc; k++) far_accumulate((float)(bx * FARB) + c, (float)(by * FARB) + c, (float)(bz * FARB) + c, s_src[k], t);
#pragma unroll
    for (int i = 0; i < FAR_T; i++) tens[(size_t)i * nblocks + b] = t[i];
}

// E_dyn = E_stat + KE * e, B_dyn = B_stat + KMU * b (sim.cl:986-992): one block per row (y, z), everything coalesced.  A thread
// owns CH cells of the row (x = threadIdx.x + k * blockDim.x); all 13 loads per cell (6 scratch components, flag, 6 static
// components) are in flight before the first is used; the scratch values pass through shared memory, which undoes the row
// permutation (combine_load / combine_write of eb_fft_core.cuh are the per-component form of the same thing, used by the CPU
// emulation).  CH = 0: rows longer than 8 * blockDim, static fields loaded after the barrier instead of being kept in registers.
template <int ND, int CH>
__global__ void __launch_bounds__(256) k_eb_combine(const __grid_constant__ Geom g, const float* __restrict__ scratch, const uint8_t* __restrict__ flags,
                                                     const float* __restrict__ E_stat, const float* __restrict__ B_stat, float* __restrict__ E_dyn,
                                                     float* __restrict__ B_dyn, const float* __restrict__ far_tens, const uint32_t far_nbx,
                                                     const uint32_t far_blocks, const uint32_t far_shift) {
    extern __shared__ __align__(128) unsigned char eb_smem[];
    float* tile = reinterpret_cast<float*>(eb_smem);  // [6][tile_len]
    const uint32_t y = blockIdx.x, z = blockIdx.y;
    if (((g.dy > 1u) & (y == 0u || y >= g.ny - 1u)) || ((g.dz > 1u) & (z == 0u || z >= g.nz - 1u))) return;  // halo rows, sim.cl:899
    const uint32_t tile_len = (g.nx / ND) * (ND + 1) + ND + 1;
    const uint64_t base = ((uint64_t)y + (uint64_t)z * g.ny) * g.nx;
    // far slabs: the Taylor polynomials of this row's FARB^3 blocks, reduced to polynomials in x alone -- along a row dy and dz are
    // constants, so each of the six outputs is a + dx (b + c dx) with (a, b, c) per block: [6][3][far_nbx] floats after the six
    // permuted rows (18 values per block instead of the 60 tensor entries: 37 KB instead of 123 KB for a 2048-cell row)
    float* ftile = tile + (size_t)6 * tile_len;
    const uint32_t FARB = 1u << far_shift;
    if (far_tens) {
        const float dy = (float)(y % FARB) - 0.5f * (float)(FARB - 1u), dz = (float)(z % FARB) - 0.5f * (float)(FARB - 1u);
        const uint32_t brow = (y / FARB + (g.ny + FARB - 1u) / FARB * (z / FARB)) * far_nbx;
        for (uint32_t i = threadIdx.x; i < 6u * far_nbx; i += blockDim.x) {
            const uint32_t o = i / far_nbx, bx = i % far_nbx, c = o % 3u;
            const float* t = far_tens + (size_t)(o / 3u) * 30u * far_blocks + brow + bx;  // E part: entries 0..29, B part: 30..59
            const float F0 = t[(size_t)c * far_blocks];
            const float* G = t + (size_t)(3u + 3u * c) * far_blocks;
            const float* H = t + (size_t)(12u + 6u * c) * far_blocks;
            const float G0 = G[0], G1 = G[far_blocks], G2 = G[2 * (size_t)far_blocks];
            const float H0 = H[0], H1 = H[far_blocks], H2 = H[2 * (size_t)far_blocks], H3 = H[3 * (size_t)far_blocks], H4 = H[4 * (size_t)far_blocks],
                        H5 = H[5 * (size_t)far_blocks];
            float pa, pb, pc;
            far_reduce_x(F0, G0, G1, G2, H0, H1, H2, H3, H4, H5, dy, dz, pa, pb, pc);
            ftile[(size_t)(3u * o) * far_nbx + bx] = pa;
            ftile[(size_t)(3u * o + 1u) * far_nbx + bx] = pb;
            ftile[(size_t)(3u * o + 2u) * far_nbx + bx] = pc;
        }
    }
    constexpr int R = CH > 0 ? CH : 1;
    float st[R][6];
    uint8_t fl[R];
    if (CH > 0) {
        float sv[R][6];
#pragma unroll
        for (int k = 0; k < R; k++) {
            const uint32_t x = threadIdx.x + (uint32_t)k * blockDim.x;
            fl[k] = 0x01;
            if (x < g.nx) {
#pragma unroll
                for (int c = 0; c < 6; c++) sv[k][c] = scratch[(uint64_t)c * g.N + base + x];
                fl[k] = flags[base + x];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    st[k][c] = E_stat[(uint64_t)c * g.N + base + x];
                    st[k][3 + c] = B_stat[(uint64_t)c * g.N + base + x];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; k++) {
            const uint32_t x = threadIdx.x + (uint32_t)k * blockDim.x;
            if (x < g.nx) {
#pragma unroll
                for (int c = 0; c < 6; c++) tile[(size_t)c * tile_len + (x / ND) * (ND + 1) + x % ND] = sv[k][c];
            }
        }
    } else {
        for (int c = 0; c < 6; c++) combine_load<ND>(threadIdx.x, blockDim.x, g, y, z, c, scratch, tile + (size_t)c * tile_len);
    }
    __syncthreads();
    if (CH > 0) {
#pragma unroll
        for (int k = 0; k < R; k++) {
            const uint32_t x = threadIdx.x + (uint32_t)k * blockDim.x;
            if (x >= g.nx || ((g.dx > 1u) & (x == 0u || x >= g.nx - 1u))) continue;  // is_halo, sim.cl:899
            if ((fl[k] & 0x1Fu) == 0x01u) continue;                                   // (flags & TYPE_BO) == TYPE_S, sim.cl:900-902
            const uint32_t p = (x % g.dsx) * (ND + 1) + x / g.dsx;
            float sum[6];
#pragma unroll
            for (int c = 0; c < 6; c++) sum[c] = tile[(size_t)c * tile_len + p];
            if (far_tens) {  // + the far slabs' field, evaluated from the block's Taylor polynomial
                const float fdx = (float)(x % FARB) - 0.5f * (float)(FARB - 1u);
                const float* q = ftile + x / FARB;
#pragma unroll
                for (int o = 0; o < 6; o++)
                    sum[o] += far_eval_x(q[(size_t)(3 * o) * far_nbx], q[(size_t)(3 * o + 1) * far_nbx], q[(size_t)(3 * o + 2) * far_nbx], fdx);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                E_dyn[(uint64_t)c * g.N + base + x] = st[k][c] + g.ke * sum[c];
                B_dyn[(uint64_t)c * g.N + base + x] = st[k][3 + c] + g.kmu * sum[3 + c];
            }
        }
    } else {
        for (int c = 0; c < 6; c++)
            combine_write<ND>(threadIdx.x, blockDim.x, g, y, z, c, tile + (size_t)c * tile_len, flags, E_stat, B_stat, E_dyn, B_dyn);
    }
}

// The polyphase path needs: LOD depth 3 or 4, x and y extents that are whole LOD blocks (z may carry the two halo layers of a
// slab: they become an extra z window), the full fine level inside the pyramid.  Everything else stays on fields.cu's kernels.
bool eb_fft_supported(const KArgs& a) {
    if (a.lod_depth != 3u && a.lod_depth != 4u) return false;
    const uint32_t nd = 1u << a.lod_depth;
    if (a.nx % nd || a.ny % nd || a.nz < nd) return false;
    if (a.n_lod_own < nd * nd * nd) return false;
    if (a.dx > 1u || a.dy > 1u) return false;  // halo layers in x / y would make the blocks ragged
    // the near-cell loop of sim.cl:907-938 must be empty (quirk Q4: block = N / 2^(2^depth); true for every lattice below 512
    // cells per axis at depth 3 and always at depth 4)
    const uint32_t sh = 1u << nd;
    if ((a.nx / sh > 1u) || (a.ny / sh > 1u) || (a.nz / sh > 1u)) return false;
    // a slab with nz % nd layers beyond block nd-1 gets ONE extra z window (eb_fft_geometry): dsz = nz / nd >= 1 gives
    // nz < nd * (dsz + 1) <= 2 * nd * dsz, so two windows always cover it
    return true;
}

static void eb_fft_geometry(const KArgs& a, Geom& g, std::vector<Task>& tasks) {
    const uint32_t nd = 1u << a.lod_depth;
    g.nx = a.nx; g.ny = a.ny; g.nz = a.nz; g.N = a.N;
    g.dsx = a.nx / nd; g.dsy = a.ny / nd; g.dsz = a.nz / nd;
    g.n_lod_own = a.n_lod_own;
    g.lo = a.n_lod_own - nd * nd * nd;
    g.cz0 = g.lo / (nd * nd);
    g.dx = a.dx; g.dy = a.dy; g.dz = a.dz;
    g.ke = a.ke; g.kmu = a.kmu;
    const uint32_t nwz = (a.nz + nd * g.dsz - 1u) / (nd * g.dsz);
    for (uint32_t wz = 0; wz < nwz; wz++)
        for (uint32_t oz = 0; oz < g.dsz; oz++) {
            // does this (window, oz) hold any cell that update_e_b_dynamic writes?  (z < nz, not a halo layer)
            bool any = false;
            for (uint32_t bz = 0; bz < nd && !any; bz++) {
                const uint32_t z = (bz + nd * wz) * g.dsz + oz;
                if (z >= a.nz) break;
                any = !(a.dz > 1u && (z == 0u || z >= a.nz - 1u));
            }
            if (!any) continue;
            // ox fastest: blocks that run at the same time write neighbouring words of the same sectors, which L2 merges
            for (uint32_t oy = 0; oy < g.dsy; oy++)
                for (uint32_t ox = 0; ox < g.dsx; ox++) tasks.push_back(Task{(uint16_t)ox, (uint16_t)oy, (uint16_t)oz, (uint16_t)wz});
        }
}

size_t eb_fft_bytes(const KArgs& a) {
    if (!eb_fft_supported(a)) return 0;
    Geom g;
    std::vector<Task> tasks;
    eb_fft_geometry(a, g, tasks);
    const size_t per = a.lod_depth == 4u ? Cfg<16>::khat_per_task : Cfg<8>::khat_per_task;
    return tasks.size() * per * sizeof(float2);
}

void eb_fft_destroy(EbFftPlan* p) {
    if (!p) return;
    if (p->tasks) cudaFree(p->tasks);
    if (p->khat) cudaFree(p->khat);
    if (p->shat) cudaFree(p->shat);
    if (p->shatc) cudaFree(p->shatc);
    if (p->scratch) cudaFree(p->scratch);
    if (p->far_src) cudaFree(p->far_src);
    if (p->far_tens) cudaFree(p->far_tens);
    delete p;
}

template <int ND> static cudaError_t build_khat(EbFftPlan* p, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_eb_khat<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<ND>::khat_smem);
    if (e != cudaSuccess) return e;
    if (p->streamed) return cudaSuccess;  // computed per batch inside every step
    const EbFftPlan::Batch& b = p->batches[0];  // static: one batch holds every slot
    for (int set = 0; set < p->nsets; set++)
        k_eb_khat<ND><<<dim3(b.nc, 3), Cfg<ND>::KT, Cfg<ND>::khat_smem, s>>>(p->g, p->sets[set], p->tasks + b.c0, p->khat + (size_t)set * p->batch * Cfg<ND>::khat_per_task);
    return cudaGetLastError();
}

// Second source set: the level D-1 pyramid of the z slab directly below (domain di - 1; slabs above sit ~4.29e9 cells away through
// the uint wrap of quirk Q18 and are skipped by the fast paths).  It convolves on the own block lattice when its blocks are exactly
// two own blocks wide on every axis; `entry0` follows the layout of mod.rs:448-465 (foreign levels appended in ascending domain
// index, self skipped).
static void foreign_set(const KArgs& a, const Geom& g, EbFftPlan* p) {
    p->foreign_domain = -1;
    if (a.dx != 1u || a.dy != 1u || a.dz < 2u || a.di == 0u || a.lod_depth < 1u) return;
    if (a.dz - 1u > 32u) return;  // the direct kernel cannot address single foreign domains any more (MAX_FOREIGN in fields.cu)
    const uint32_t nd = 1u << a.lod_depth, nf = nd / 2u;
    if (a.nx / nf != 2u * g.dsx || a.ny / nf != 2u * g.dsy || a.nz / nf != 2u * g.dsz) return;
    uint32_t entry = a.n_lod_own;
    for (uint32_t d = 0; d + 1u < a.di; d++) {  // levels of the domains below the neighbour: max(depth - distance, 0)
        const int level = (int)a.lod_depth - (int)(a.di - d) > 0 ? (int)a.lod_depth - (int)(a.di - d) : 0;
        entry += 1u << (3 * level);
    }
    SourceSet& ss = p->sets[1];
    ss.kind = 1u;
    ss.entry0 = entry;
    ss.zadd = (int32_t)(a.nz / g.dsz);  // slab height incl. halos = zadd * dsz + offz (quirk Q8)
    ss.offx = 0.0f; ss.offy = 0.0f; ss.offz = (float)(a.nz % g.dsz);
    p->nsets = 2;
    p->foreign_domain = (int)a.di - 1;
}

// Far slabs: domains two and more below this one (z slabs; the ones above are skipped like everywhere in the fast paths, quirk Q18).
// Fills the descriptors and decides whether the Taylor path may sum them: every (cell, source) distance must be at least
// 40 x the largest offset inside a FARB^3 block, which bounds the truncation error by 1.6e-5 of the far field.
static void far_set(const KArgs& a, EbFftPlan* p) {
    p->far_n = 0;
    p->far_ndesc = 0;
    p->far_handled = false;
    p->far_shift = 3u;
    if (p->foreign_domain < 0 || a.di < 2u) return;  // no far slab, or the neighbour is not on the FFT path either
    bool ok = a.nx <= 2048u && a.di - 1u <= 8u;
    uint32_t entry = a.n_lod_own;
    float r_min = 1.0e30f;
    for (uint32_t d = 0; d + 1u < a.di; d++) {
        const uint32_t ddz = a.di - d;
        const int level = (int)a.lod_depth - (int)ddz > 0 ? (int)a.lod_depth - (int)ddz : 0;
        const uint32_t nf = 1u << level;
        if (ok) {
            FarDesc& f = p->far_desc[p->far_ndesc++];
            f.entry0 = entry;
            f.nf = nf;
            f.bsx = (float)(a.nx / nf); f.bsy = (float)(a.ny / nf); f.bsz = (float)(a.nz / nf);
            f.shift_z = (float)(ddz * a.nz);
            const float top = ((float)(nf - 1u) * f.bsz + 0.5f * f.bsz) - f.shift_z;  // highest centre of this slab's level
            r_min = fminf(r_min, 1.0f - top);                                          // the lowest cell a slab updates is z = 1
            p->far_n += (int)(nf * nf * nf);
        }
        entry += nf * nf * nf;
    }
    if (p->far_n > 128 || p->far_n == 0) ok = false;
    const float d8 = 3.5f * 1.7320508f, d4 = 1.5f * 1.7320508f;  // largest offset from the centre of an 8^3 / 4^3 block
    if (ok && r_min >= 40.0f * d8) p->far_shift = 3u;
    else if (ok && r_min >= 25.0f * d4) p->far_shift = 2u;
    else ok = false;
    p->far_handled = ok;
    if (!p->far_handled) { p->far_n = 0; p->far_ndesc = 0; }
}

// Builds the static part (task list, K^) on the domain's device.  *out = nullptr (and cudaSuccess) when the geometry is not
// supported or the spectra do not fit `budget_bytes`.
cudaError_t eb_fft_create(const KArgs& a, size_t budget_bytes, cudaStream_t s, EbFftPlan** out, uint64_t* launches) {
    *out = nullptr;
    if (!eb_fft_supported(a)) return cudaSuccess;
    EbFftPlan* p = new EbFftPlan();
    p->nd = 1 << a.lod_depth;
    p->tasks = nullptr; p->khat = nullptr; p->shat = nullptr; p->shatc = nullptr; p->scratch = nullptr;
    p->far_src = nullptr; p->far_tens = nullptr; p->far_n = 0; p->far_handled = false; p->far_ndesc = 0;
    std::vector<Task> tasks;
    eb_fft_geometry(a, p->g, tasks);
    p->ntasks = (uint32_t)tasks.size();
    p->nsets = 1;
    p->sets[0] = SourceSet{0u, 0u, -(int32_t)p->g.cz0, -0.5f * (float)p->g.dsx, -0.5f * (float)p->g.dsy, -0.5f * (float)p->g.dsz};
    foreign_set(a, p->g, p);
    const size_t per = p->nd == 16 ? Cfg<16>::khat_per_task : Cfg<8>::khat_per_task;
    const size_t sh = p->nd == 16 ? Cfg<16>::shat_count : Cfg<8>::shat_count;
    const size_t scratch_bytes = (size_t)6 * a.N * sizeof(float);
    // Which tasks own a slot of kernel spectra.  Preference: (1) every task, all slots resident (fastest products); (2) canonical tasks
    // only, all slots resident (mirrored tasks read their partner's); (3) canonical tasks only, a buffer of `batch` slots refilled
    // batch by batch inside every step (streamed).
    const uint32_t forced = getenv("ION_EB_FFT_BATCH") ? (uint32_t)atoi(getenv("ION_EB_FFT_BATCH")) : 0u;  // test hook: streamed mode with this many slots
    const int mirror_env = getenv("ION_EB_FFT_MIRROR") ? atoi(getenv("ION_EB_FFT_MIRROR")) : -1;       // test hook: 1 = mirrored even if (1) fits, 0 = never
    const size_t slot_bytes = (size_t)p->nsets * per * sizeof(float2);
    std::vector<Task> canon, mirr;
    split_mirror_tasks(tasks, p->g.dsx, canon, mirr);
    const bool fits_full = (size_t)p->ntasks * slot_bytes + scratch_bytes <= budget_bytes;
    const bool fits_canon = canon.size() * slot_bytes + scratch_bytes <= budget_bytes;
    const bool use_mirror = mirror_env != 0 && !mirr.empty() && (mirror_env == 1 || forced > 0u || !fits_full);
    const uint32_t nslots_all = use_mirror ? (uint32_t)canon.size() : p->ntasks;
    p->streamed = (forced > 0u && forced < nslots_all) || !(use_mirror ? fits_canon : fits_full);
    p->batch = nslots_all;
    if (p->streamed) {  // at most ~1.7 GB per set, at least a few waves of blocks
        const uint32_t min_batch = nslots_all < 1024u ? nslots_all : 1024u;
        if (!forced && scratch_bytes + (size_t)min_batch * slot_bytes > budget_bytes) { delete p; return cudaSuccess; }
        size_t room = forced ? forced : (budget_bytes - scratch_bytes) / slot_bytes;
        if (room > 2048u) room = 2048u;
        if (room > nslots_all) room = nslots_all;
        p->batch = (uint32_t)room;
        if (use_mirror) p->batch = mirror_batch_slots(p->batch, p->g.dsx);
    }
    p->khat_bytes = (size_t)p->batch * slot_bytes;
    {  // lay the task array out batch by batch (eb_fft_layout.hpp)
        std::vector<Task> laid;
        std::vector<TaskBatch> batches;
        layout_tasks(tasks, canon, mirr, use_mirror, p->batch, laid, batches);
        for (const TaskBatch& b : batches) p->batches.push_back(EbFftPlan::Batch{b.c0, b.nc, b.m0, b.nm});
        tasks.swap(laid);
    }
    far_set(a, p);
    {  // room for the pinned source spectra in L2 (a device-wide limit; a few MB of 126)
        size_t cur = 0;
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) != cudaSuccess || cur < (size_t)8 << 20) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)8 << 20) != cudaSuccess) cudaGetLastError();
        }
    }
    {
        const uint32_t fb = 1u << p->far_shift;
        p->far_nbx = (a.nx + fb - 1u) / fb; p->far_nby = (a.ny + fb - 1u) / fb; p->far_nbz = (a.nz + fb - 1u) / fb;
    }
    cudaError_t e = cudaMalloc((void**)&p->tasks, tasks.size() * sizeof(Task));
    if (e == cudaSuccess && p->far_handled) e = cudaMalloc((void**)&p->far_src, 128 * sizeof(FarSource));
    if (e == cudaSuccess && p->far_handled) e = cudaMalloc((void**)&p->far_tens, (size_t)FAR_T * p->far_nbx * p->far_nby * p->far_nbz * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->khat, p->khat_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->shat, (size_t)p->nsets * sh * sizeof(float2));
    if (e == cudaSuccess && p->nsets == 2)
        e = cudaMalloc((void**)&p->shatc, (p->nd == 16 ? Cfg<16>::shatc_count : Cfg<8>::shatc_count) * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->scratch, scratch_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->tasks, tasks.data(), tasks.size() * sizeof(Task), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // `tasks` is a local vector
    if (e == cudaSuccess) e = p->nd == 16 ? build_khat<16>(p, s) : build_khat<8>(p, s);
    if (e != cudaSuccess) {
        eb_fft_destroy(p);
        if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return cudaSuccess; }  // no room: the direct kernels stay in charge
        return e;
    }
    *launches += (uint64_t)p->nsets;
    *out = p;
    return cudaSuccess;
}

// The source spectra (1-2 MB) are read by every block of k_eb_fft while 2-4 GB of kernel spectra stream through L2 once: pin them
// with a persisting access-policy window for the duration of the kernel so that the stream does not evict them.
static void pin_source_spectra(const EbFftPlan* p, cudaStream_t s, size_t bytes, bool on) {
    static const bool enabled = !(getenv("ION_EB_L2PIN") && atoi(getenv("ION_EB_L2PIN")) == 0);
    if (!enabled) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = (void*)p->shat;
    v.accessPolicyWindow.num_bytes = on ? bytes : 0;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = on ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();  // best effort
}

template <int ND, int NSETS, int MIRROR> static void launch_fft(const EbFftPlan* p, uint32_t first, uint32_t count, cudaStream_t s) {
    if (count == 0u) return;
    const float2* k2 = NSETS == 2 ? p->khat + (size_t)p->batch * Cfg<ND>::khat_per_task : nullptr;
    k_eb_fft<ND, NSETS, MIRROR><<<count, Cfg<ND>::T, Cfg<ND>::main_smem, s>>>(p->g, p->tasks + first, p->khat, p->shat, k2, NSETS == 2 ? p->shatc : nullptr, p->scratch, 0);
}
template <int ND> static cudaError_t launch_nd(const EbFftPlan* p, const KArgs& a, cudaStream_t s) {
    // function attributes are per device: set on every launch (a host-side table lookup)
    cudaError_t e = cudaSuccess;
    const int smem = (int)Cfg<ND>::main_smem;
    if (p->nsets == 2) {
        e = cudaFuncSetAttribute(k_eb_fft<ND, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eb_fft<ND, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    } else {
        e = cudaFuncSetAttribute(k_eb_fft<ND, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eb_fft<ND, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (e != cudaSuccess) return e;
    const size_t kstride = (size_t)p->batch * Cfg<ND>::khat_per_task, sstride = Cfg<ND>::shat_count;
    for (int set = 0; set < p->nsets; set++)
        k_eb_src<ND><<<dim3(Cfg<ND>::H, 4), 128, 0, s>>>(p->g, p->sets[set], a.QU_lod, p->shat + (size_t)set * sstride, set == 1 ? p->shatc : nullptr);
    if (p->streamed) {
        e = cudaFuncSetAttribute(k_eb_khat<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<ND>::khat_smem);
        if (e != cudaSuccess) return e;
    }
    pin_source_spectra(p, s, (size_t)p->nsets * sstride * sizeof(float2), true);
    for (const EbFftPlan::Batch& b : p->batches) {
        if (p->streamed)
            for (int set = 0; set < p->nsets; set++)
                k_eb_khat<ND><<<dim3(b.nc, 3), Cfg<ND>::KT, Cfg<ND>::khat_smem, s>>>(p->g, p->sets[set], p->tasks + b.c0, p->khat + (size_t)set * kstride);
        if (p->nsets == 2) {
            launch_fft<ND, 2, 0>(p, b.c0, b.nc, s);
            launch_fft<ND, 2, 1>(p, b.m0, b.nm, s);
        } else {
            launch_fft<ND, 1, 0>(p, b.c0, b.nc, s);
            launch_fft<ND, 1, 1>(p, b.m0, b.nm, s);
        }
    }
    pin_source_spectra(p, s, 0, false);
    const uint32_t far_blocks = p->far_nbx * p->far_nby * p->far_nbz;
    if (p->far_handled) {
        FarDescs fd;
        fd.n = p->far_ndesc;
        for (int i = 0; i < fd.n; i++) fd.d[i] = p->far_desc[i];
        k_eb_far_sources<<<1, 128, 0, s>>>(fd, a.QU_lod, p->far_src);
        k_eb_far_tensors<<<(far_blocks + 127u) / 128u, 128, 0, s>>>(p->far_src, p->far_n, p->far_tens, p->far_nbx, p->far_nby, p->far_nbz, 1u << p->far_shift);
    }
    const float* far_tens = p->far_handled ? p->far_tens : nullptr;
    const uint32_t tile_len = (a.nx / ND) * (ND + 1) + ND + 1;
    const size_t csmem = (size_t)6 * tile_len * sizeof(float) + (p->far_handled ? (size_t)18 * p->far_nbx * sizeof(float) : 0);
    const uint32_t threads = a.nx < 256u ? ((a.nx + 31u) / 32u) * 32u : 256u;
    const uint32_t chunks = (a.nx + threads - 1u) / threads;
    const dim3 grid(a.ny, a.nz);
#define ION_COMBINE(CH)                                                                                                          \
    do {                                                                                                                           \
        if (csmem > 48u * 1024u) {                                                                                                 \
            e = cudaFuncSetAttribute(k_eb_combine<ND, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);               \
            if (e != cudaSuccess) return e;                                                                                        \
        }                                                                                                                          \
        k_eb_combine<ND, CH><<<grid, threads, csmem, s>>>(p->g, p->scratch, a.flags, a.E_stat, a.B_stat, a.E_dyn, a.B_dyn, far_tens, \
                                                          p->far_nbx, far_blocks, p->far_shift);                                    \
    } while (0)
    if (chunks <= 1u) ION_COMBINE(1);
    else if (chunks <= 2u) ION_COMBINE(2);
    else if (chunks <= 4u) ION_COMBINE(4);
    else if (chunks <= 8u) ION_COMBINE(8);
    else ION_COMBINE(0);
#undef ION_COMBINE
    return cudaGetLastError();
}
cudaError_t eb_fft_launch(const EbFftPlan* p, const KArgs& a, cudaStream_t s, uint64_t* launches) {
    uint64_t n = 1u + (uint64_t)p->nsets + (p->far_handled ? 2u : 0u);  // k_eb_combine, k_eb_src per set, the far-slab kernels
    for (const EbFftPlan::Batch& b : p->batches) n += (p->streamed ? (uint64_t)p->nsets : 0u) + (b.nc ? 1u : 0u) + (b.nm ? 1u : 0u);
    *launches += n;
    return p->nd == 16 ? launch_nd<16>(p, a, s) : launch_nd<8>(p, a, s);
}
size_t eb_fft_plan_bytes(const EbFftPlan* p) { return p ? p->khat_bytes : 0; }
int eb_fft_plan_foreign_domain(const EbFftPlan* p) { return p ? p->foreign_domain : -1; }
bool eb_fft_plan_far_handled(const EbFftPlan* p) { return p && p->far_handled; }
size_t eb_fft_scratch_bytes(const EbFftPlan* p) { return p ? (size_t)6 * p->g.N * sizeof(float) : 0; }
uint32_t eb_fft_plan_tasks(const EbFftPlan* p) { return p ? p->ntasks : 0; }
uint32_t eb_fft_plan_batch(const EbFftPlan* p) { return p ? p->batch : 0; }

}  // namespace ion
