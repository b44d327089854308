// stream_collide.cuh -- the fused hot kernel of the extended-LBM MHD time step, plus update_fields and
// initialize, as templates over <velocity set, DDF storage type, MHD, TRT>.
//
// Behaviour follows /root/reference/src/kernels/sim_kernels.cl ("sim.cl"): stream_collide :465-758,
// initialize :760-832, update_fields :834-859.  Design for B200 (sm_100a):
//   * one thread per lattice cell, x fastest: every DDF access of a warp is one contiguous 128 B (FP32) / 64 B (FP16) segment per
//     direction -- the SoA layout i*N+n keeps all 2Q(+2Q+14) streams coalesced (the plain FP32 kernels use the four-cells-per-thread
//     variant with 128-bit accesses of stream_collide_v4.cuh);
//   * Esoteric-Pull in-place streaming: each (cell,slot) address is read and then written by the same thread, so no second DDF copy
//     exists and algorithmic HBM traffic is the minimum 2*Q*s bytes/cell;
//   * 3-D launch (x-chunk, y, z): no integer div/mod per cell, branch-free periodic wraps; the step parity is a template parameter,
//     so slot indices are constants and an address costs two integer instructions;
//   * all loads of a cell (19+19+7 DDFs, E, B) are pinned in program order IN FRONT of the branch on the flag byte (ptxas would sink
//     them below it) -> >50 independent requests in flight per thread, which is what saturates HBM3e at modest occupancy;
//   * MHD source terms (electron-gas LBM, D3Q7 charge advection, Lorentz force) are applied in registers; the neutral gas and the
//     electron gas share the two lanes of packed FP32 (collide_two_species, lattice.cuh);
//   * the LOD deposit is a warp-segmented shuffle reduction followed by one 16-byte vector reduction per (warp, LOD block) into one
//     of 32 private replicas (same-address reductions serialise in L2) instead of the reference's 4 same-address atomics per cell
//     (sim.cl:673-676);
//   * EQUILIBRIUM_BOUNDARIES / VOLUME_FORCE / FORCE_FIELD / UPDATE_FIELDS are warp-uniform runtime switches
//     (they do not change register pressure materially); Q, storage codec, MHD, TRT and the step parity are compile time.
// No tensor cores: the kernel is HBM-bound (153 B/cell plain, 389 B/cell MHD for D3Q19 FP32).
#pragma once
#include <cstdlib>

#include "lattice.cuh"

namespace ion {

constexpr int SC_BLOCK_MAX = 256;
// stream_collide launch shape.  Measured at 256^3 D3Q19 FP32 MHD (profiles/r1_stream_collide_ab.md): blocks of 64 threads
// reach 5.12 TB/s, 128 -> 5.07, 256 -> 4.76 (8 / 4 / 2 resident blocks per SM at 128 registers: finer blocks drain and refill
// the SM more evenly, so fewer bytes-in-flight bubbles at block boundaries).  ION_SC_MINB is a build-time experiment switch
// (register cap = 65536 / (ION_SC_BLOCK * ION_SC_MINB)).
#ifndef ION_SC_BLOCK
#define ION_SC_BLOCK 64
#endif
#ifndef ION_SC_MINB
#define ION_SC_MINB 8
#endif
// ION_SC_PACKED = 1: the MHD kernel evaluates the neutral gas and the electron gas together in packed FP32 (lattice.cuh,
// collide_two_species); 0: the scalar sequence (A/B build switch).  Results are bit-identical either way.
#ifndef ION_SC_PACKED
#define ION_SC_PACKED 1
#endif

struct LodDeposit {
    uint32_t ind;  // entry of the own finest LOD level (lod_index of the cell), 0xFFFFFFFF = nothing to deposit
    float q, ux, uy, uz;
};

// LOD deposit (sim.cl:666-677: four float atomics per cell onto the cell's finest-level LOD entry).
//   1. Lanes of a warp are consecutive x cells, so equal LOD entries form contiguous runs: each run is summed with shuffles and
//      only its head lane touches memory.
//   2. Same-address reductions serialise in L2 (~10 ns each, measured: 3.8 ms for the 256^3 kernel at depth 1 against 1.2 ms at
//      depth 3), and the blocks resident at one time sit in a handful of LOD blocks -- 20 entries at 512^3, 16 at
//      2048x2048x128 -- so with one copy of the pyramid the deposit alone took 8 ms of the 12.7 ms 512^3 kernel.  The deposit
//      therefore goes to one of lod_rep_count private replicas of the finest level, picked by block coordinates so that
//      neighbouring resident blocks use different replicas, as ONE 16-byte vector reduction (red.global.add.v4.f32) per run;
//      k_lod_fold adds the replicas into QU_lod (and clears them) right after the kernel.
__device__ __forceinline__ void lod_deposit_warp(const KArgs& a, LodDeposit d, uint32_t own_offset) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t prev = __shfl_up_sync(full, d.ind, 1);
    const bool head = (lane == 0u) || (prev != d.ind);
    const unsigned heads = __ballot_sync(full, head);
    const unsigned higher = lane == 31u ? 0u : (heads & ~((2u << lane) - 1u));
    const int run_end = higher ? (__ffs(higher) - 1) : 32;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float oq = __shfl_down_sync(full, d.q, off);
        const float ox = __shfl_down_sync(full, d.ux, off);
        const float oy = __shfl_down_sync(full, d.uy, off);
        const float oz = __shfl_down_sync(full, d.uz, off);
        if ((int)lane + off < run_end) {
            d.q += oq; d.ux += ox; d.uy += oy; d.uz += oz;
        }
    }
    if (head && d.ind != 0xFFFFFFFFu) {
        if (d.ind < a.lod_rep_entries) {
            const uint32_t rep = (blockIdx.x + 7u * blockIdx.y + 13u * blockIdx.z) & a.lod_rep_mask;
            float* p = a.lod_rep + ((size_t)rep * a.lod_rep_entries + d.ind) * 4u;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(d.q), "f"(d.ux), "f"(d.uy), "f"(d.uz) : "memory");
        } else {  // quirk Q7: an index past the own finest level but inside QU_lod is hit like in the reference
            float* p = a.QU_lod + (size_t)(d.ind + own_offset) * 4u;
            atomicAdd(p + 0, d.q);
            atomicAdd(p + 1, d.ux);
            atomicAdd(p + 2, d.uy);
            atomicAdd(p + 3, d.uz);
        }
    }
}

// Resident blocks (of 64 threads) per SM the register allocation has to allow, measured per kernel family at 256^3
// (profiles/r1_stream_collide_ab.md, r1b table): the MHD kernels keep 2Q+7+6 loaded values live -- 128 registers / 8 blocks for
// FP32 and FP16S (capping lower spills and is slower), 96 registers / 10 blocks for FP16C whose codec needs fewer temporaries.
// The plain kernels have only Q loads per thread in flight and need more warps to cover the HBM latency: FP32 SRT is best at
// 80 registers / 12 blocks (0.87 of the copy peak vs 0.83 at 16 blocks, whose 64-register cap spills), TRT and the FP16 codecs
// at 16 blocks, D3Q27 at 9.
#ifndef ION_SC_MINB_PLAIN
#define ION_SC_MINB_PLAIN 16
#endif
template <int VS, int FP, bool MHD, bool TRT> struct ScMinBlocks {
    static constexpr int value = MHD ? (FP == ION_FP16C ? 10 : ION_SC_MINB)
                                     : (VSet<VS>::Q > 19 ? 9 : (FP == ION_FP32 && !TRT ? 12 : ION_SC_MINB_PLAIN));
};

// SUBGRID_ECR helpers, sim.cl:449-461
__device__ __forceinline__ float mag_v(const float* __restrict__ V, uint64_t N, uint32_t n) {
    return sqrtf(sq(V[n]) + sq(V[N + n]) + sq(V[2ull * N + n]));
}
__device__ __forceinline__ float length3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }  // OpenCL length()

template <int VS, int FP, bool MHD, bool TRT, bool ECR, bool ODD>
__global__ void __launch_bounds__(ION_SC_BLOCK, ScMinBlocks<VS, FP, MHD, TRT>::value)
k_stream_collide(const __grid_constant__ KArgs a, const float fx, const float fy, const float fz) {
    constexpr int QQ = VSet<VS>::Q;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z + a.z_off;
    const bool inside = x < a.nx && !is_halo(a, x, y, z);  // sim.cl:485
    const Cell c = make_cell(a, inside ? x : 0u, y, z);
    const uint32_t n = c.n;
    const uint64_t N = a.N;
    auto nb = [&](int i) { return neighbor<VS>(c, i); };

    // ---- every load of the cell is issued BEFORE the flag is known (sim.cl:487 returns first): the flag byte would
    // otherwise cost one full HBM round trip during which the warp has nothing in flight.  A solid cell's DDFs are read and
    // dropped (a few % extra reads in scenes with solids, none in the fluid bulk); results are unchanged. ----
    float fhn[QQ];
    ep_load_p<FP, QQ, ODD>(fhn, a.fi, N, n, nb);  // streaming part 2, sim.cl:494
    float ehn[MHD ? QQ : 1];
    float qhn[7];
    float Bx = 0.f, By = 0.f, Bz = 0.f, Ex = 0.f, Ey = 0.f, Ez = 0.f;
    if (MHD) {
        Bx = ld_f32_pinned(a.B_dyn + n); By = ld_f32_pinned(a.B_dyn + N + n); Bz = ld_f32_pinned(a.B_dyn + 2ull * N + n);  // sim.cl:532-533
        Ex = ld_f32_pinned(a.E_dyn + n); Ey = ld_f32_pinned(a.E_dyn + N + n); Ez = ld_f32_pinned(a.E_dyn + 2ull * N + n);
        ep_load_p<FP, QQ, ODD>(ehn, a.ei, N, n, nb);                         // sim.cl:538
        ep_load_p<FP, 7, ODD>(qhn, a.fqi, N, n, [&](int i) { return neighbor7(c, i); });  // sim.cl:550
    }
    float ethn[ECR ? 7 : 1];
    float Evx = 0.f, Evy = 0.f, Evz = 0.f;
    if (ECR) {
        ep_load_p<FP, 7, ODD>(ethn, a.eti, N, n, [&](int i) { return neighbor7(c, i); });  // sim.cl:566
        Evx = a.E_var[n]; Evy = a.E_var[N + n]; Evz = a.E_var[2ull * N + n];                // sim.cl:579
    }
    // the flag byte is requested LAST: the loads above are pinned (asm volatile), so the branch on the flag cannot be scheduled
    // in front of them (nvcc otherwise duplicates the loads into both sides of that branch and the warp idles one round trip)
    const uint8_t flagsn = inside ? ld_u8_pinned(a.flags + n) : (uint8_t)ION_TYPE_S;
    const bool active = (flagsn & ION_TYPE_BO) != ION_TYPE_S;  // sim.cl:487-488 (quirk Q1: only exact TYPE_S is solid)

    LodDeposit dep;
    dep.ind = 0xFFFFFFFFu;
    dep.q = dep.ux = dep.uy = dep.uz = 0.0f;
    uint32_t lod_off = 0u;  // first entry of the own finest level inside QU_lod (multi-domain layout), sim.cl:667-670
    if (!MHD && !active) return;

    if (active) {
        const uint8_t bo = flagsn & ION_TYPE_BO;
        const bool eqb = (a.ext & ION_EXT_EQUILIBRIUM_BOUNDARIES) != 0u;
        const bool vf = (a.ext & ION_EXT_VOLUME_FORCE) != 0u;
        const bool is_e = eqb && bo == ION_TYPE_E;

        float rhon, uxn, uyn, uzn;
        float rhon_e = 0.0f, uxn_e = 0.0f, uyn_e = 0.0f, uzn_e = 0.0f;  // electron gas (MHD)
        float efx = 0.0f, efy = 0.0f, efz = 0.0f;
        if (MHD && ION_SC_PACKED) {  // gas and electron moments in the two lanes of packed FP32 (sim.cl:496-511,537-540)
            float2 rho2, m2[3];
            rho_m2<VS>(fhn, ehn, rho2, m2);
            rhon = rho2.x; uxn = m2[0].x / rho2.x; uyn = m2[1].x / rho2.x; uzn = VS == ION_D2Q9 ? 0.0f / rho2.x : m2[2].x / rho2.x;
            rhon_e = rho2.y; uxn_e = m2[0].y / rho2.y; uyn_e = m2[1].y / rho2.y; uzn_e = VS == ION_D2Q9 ? 0.0f / rho2.y : m2[2].y / rho2.y;
            if (is_e) {
                rhon = a.rho[n];
                uxn = a.u[n];
                uyn = a.u[N + n];
                uzn = a.u[2ull * N + n];
            }
        } else if (is_e) {  // sim.cl:503-507
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

        if (MHD) {
            // electron gas part 1, sim.cl:537-540
            if (!ION_SC_PACKED) rho_u<VS>(ehn, rhon_e, uxn_e, uyn_e, uzn_e);
            // gas charge advection 1, sim.cl:551-553
            float rhon_q = 0.0f;
#pragma unroll
            for (int i = 0; i < 7; i++) rhon_q += qhn[i];
            rhon_q += 1.0f;
            if (ECR) {  // sim.cl:556-629.  rel_absorbtion (sim.cl:582-585, the only double-precision expression) is computed
                        // and printed by the reference but never used, so it is not evaluated here.
                float Etn = 0.0f;  // electron temperature 1
#pragma unroll
                for (int i = 0; i < 7; i++) Etn += ethn[i];
                Etn += 1.0f;
                // ECR heating: field component perpendicular to B
                const float lb = length3(Bx, By, Bz);
                const float sc = (Evx * Bx + Evy * By + Evz * Bz) / sq(lb);
                const float Env_mag = length3(Evx - sc * Bx, Evy - sc * By, Evz - sc * Bz);
                Etn += a.keabs / a.kkbme * (rhon_e + 0.00001f) / a.kkge * sq(Env_mag);
                // drift of gyrating electrons along grad|B| (central differences with the reference's `a - b / 2.0f`)
                float gx = 0.0f, gy = 0.0f, gz = 0.0f;
                if (!(x == 0u || x == a.nx - 1u || y == 0u || y == a.ny - 1u || z == 0u || z == a.nz - 1u)) {
                    const uint32_t nxy = a.nx * a.ny;
                    gx = mag_v(a.B_dyn, N, n + 1u) - mag_v(a.B_dyn, N, n - 1u) / 2.0f;
                    gy = mag_v(a.B_dyn, N, n + a.nx) - mag_v(a.B_dyn, N, n - a.nx) / 2.0f;
                    gz = mag_v(a.B_dyn, N, n + nxy) - mag_v(a.B_dyn, N, n - nxy) / 2.0f;
                }
                const float ke_t = a.kkbme * Etn;
                const float dux = (ke_t * gx) / lb, duy = (ke_t * gy) / lb, duz = (ke_t * gz) / lb;
                uxn_e += dux;
                uyn_e += duy;
                uzn_e += duz;
                Etn -= length3(dux, duy, duz) / a.kkbme;
                // electron temperature 2
                float eteq[7];
                a_eq(Etn, uxn_e, uyn_e, uzn_e, eteq);
                if (a.ext & ION_EXT_UPDATE_FIELDS) a.Et[n] = Etn;
                const float wq = a.wq;
#pragma unroll
                for (int i = 0; i < 7; i++) ethn[i] = fmaf(1.0f - wq, ethn[i], wq * eteq[i]);
                ep_store_p<FP, 7, ODD>(ethn, a.eti, N, n, [&](int i) { return neighbor7(c, i); });
                // ionization
                const float delta_q_rho = 0.0001f * Etn;
                rhon_e += delta_q_rho;
                rhon_q += delta_q_rho;
            }
            // gas charge advection 2, sim.cl:633-637
            a.Q[n] = rhon_q - rhon_e;
            {
                float qeq[7];
                a_eq(rhon_q, uxn, uyn, uzn, qeq);
                const float wq = a.wq;
#pragma unroll
                for (int i = 0; i < 7; i++) qhn[i] = fmaf(1.0f - wq, qhn[i], wq * qeq[i]);
                ep_store_p<FP, 7, ODD>(qhn, a.fqi, N, n, [&](int i) { return neighbor7(c, i); });
            }
            // electron gas part 2, sim.cl:641-656
            const float nre = -rhon_e;
            efx = nre * (Ex + (uyn_e * Bz - uzn_e * By));
            efy = nre * (Ey + (uzn_e * Bx - uxn_e * Bz));
            efz = nre * (Ez + (uxn_e * By - uyn_e * Bx));
            const float rho2_e = 0.5f / (rhon_e * a.kkge);
            uxn_e = clampf(fmaf(efx, rho2_e, uxn_e), -ION_DEF_C, ION_DEF_C);
            uyn_e = clampf(fmaf(efy, rho2_e, uyn_e), -ION_DEF_C, ION_DEF_C);
            uzn_e = clampf(fmaf(efz, rho2_e, uzn_e), -ION_DEF_C, ION_DEF_C);
            if (!ION_SC_PACKED) {  // scalar electron relaxation; the packed build relaxes both species together below
                forcing_terms<VS>(uxn_e, uyn_e, uzn_e, efx, efy, efz, Fin);
                f_eq<VS>(rhon_e, uxn_e, uyn_e, uzn_e, feq);
#pragma unroll
                for (int i = 0; i < QQ; i++) {
                    const float Fi = Fin[i] * c_tau;
                    ehn[i] = is_e ? feq[i] : fmaf(1.0f - w, ehn[i], fmaf(w, feq[i], Fi));  // always SRT, sim.cl:649-655
                }
                ep_store_p<FP, QQ, ODD>(ehn, a.ei, N, n, nb);
            }
            // EM force on gas (pre-force gas velocity), sim.cl:660-662
            fxn += rhon_q * (Ex + uyn * Bz - uzn * By);
            fyn += rhon_q * (Ey + uzn * Bx - uxn * Bz);
            fzn += rhon_q * (Ez + uxn * By - uyn * Bx);
            // LOD construction, sim.cl:666-677
            if (a.lod_depth > 0u) {
                uint32_t off = 0u;
                if (a.dx > 1u || a.dy > 1u || a.dz > 1u) {
                    for (uint32_t d = 0u; d < a.lod_depth; d++) off += 1u << (d * (uint32_t)VSet<VS>::DIM);
                }
                // quirk Q7: on split axes the halo-inclusive division can point past the own finest level; entries
                // inside the buffer are hit like in the reference, a deposit past DEF_NUM_LOD (undefined behaviour
                // there) is dropped
                const uint32_t lc = lod_index(a, x, y, z, a.lod_depth);
                dep.ind = lc + off < a.n_lod ? lc : 0xFFFFFFFFu;
                lod_off = off;
                const float ils = 1.0f / lod_s(a, a.lod_depth);
                dep.q = rhon_q - rhon_e;
                dep.ux = uxn * ils;
                dep.uy = uyn * ils;
                dep.uz = uzn * ils;
            }
        }

        constexpr bool PACKED = MHD && ION_SC_PACKED;
        if (vf) {  // sim.cl:680-685
            const float rho2 = 0.5f / rhon;
            uxn = clampf(fmaf(fxn, rho2, uxn), -ION_DEF_C, ION_DEF_C);
            uyn = clampf(fmaf(fyn, rho2, uyn), -ION_DEF_C, ION_DEF_C);
            uzn = clampf(fmaf(fzn, rho2, uzn), -ION_DEF_C, ION_DEF_C);
            if (!PACKED) forcing_terms<VS>(uxn, uyn, uzn, fxn, fyn, fzn, Fin);
        } else {  // sim.cl:687-690
            uxn = clampf(uxn, -ION_DEF_C, ION_DEF_C);
            uyn = clampf(uyn, -ION_DEF_C, ION_DEF_C);
            uzn = clampf(uzn, -ION_DEF_C, ION_DEF_C);
            if (!PACKED) {
#pragma unroll
                for (int i = 0; i < QQ; i++) Fin[i] = 0.0f;
            }
        }

        if ((a.ext & ION_EXT_UPDATE_FIELDS) && !is_e) {  // sim.cl:694-710
            a.rho[n] = rhon;
            a.u[n] = uxn;
            a.u[N + n] = uyn;
            a.u[2ull * N + n] = uzn;
        }

        if (PACKED) {  // equilibrium, forcing and relaxation of gas + electrons, two lanes (sim.cl:641-656,712-755)
            collide_two_species<VS, TRT>(fhn, ehn, make_float2(rhon, rhon_e), make_float2(uxn, uxn_e), make_float2(uyn, uyn_e),
                                         make_float2(uzn, uzn_e), make_float2(fxn, efx), make_float2(fyn, efy), make_float2(fzn, efz), w, true /* MHD implies VOLUME_FORCE, checked at create */, is_e);
            ep_store_p<FP, QQ, ODD>(ehn, a.ei, N, n, nb);
            ep_store_p<FP, QQ, ODD>(fhn, a.fi, N, n, nb);  // streaming part 1, sim.cl:757
        } else {
        f_eq<VS>(rhon, uxn, uyn, uzn, feq);  // sim.cl:712

        if (!TRT) {  // sim.cl:714-723
#pragma unroll
            for (int i = 0; i < QQ; i++) {
                const float Fi = vf ? Fin[i] * c_tau : Fin[i];
                fhn[i] = is_e ? feq[i] : fmaf(1.0f - w, fhn[i], fmaf(w, feq[i], Fi));
            }
        } else {  // sim.cl:725-755
            const float wp = w;
            const float wm = 1.0f / (0.1875f / (1.0f / w - 0.5f) + 0.5f);
            if (vf) {
                const float c_taup = fmaf(wp, -0.25f, 0.5f), c_taum = fmaf(wm, -0.25f, 0.5f);
                float Fib[QQ];
                Fib[0] = Fin[0];
#pragma unroll
                for (int i = 1; i < QQ; i += 2) {
                    Fib[i] = Fin[i + 1];
                    Fib[i + 1] = Fin[i];
                }
#pragma unroll
                for (int i = 0; i < QQ; i++) Fin[i] = fmaf(c_taup, Fin[i] + Fib[i], c_taum * (Fin[i] - Fib[i]));
            }
            float fhb[QQ], feb[QQ];
            fhb[0] = fhn[0];
            feb[0] = feq[0];
#pragma unroll
            for (int i = 1; i < QQ; i += 2) {
                fhb[i] = fhn[i + 1];
                fhb[i + 1] = fhn[i];
                feb[i] = feq[i + 1];
                feb[i + 1] = feq[i];
            }
#pragma unroll
            for (int i = 0; i < QQ; i++) {
                fhn[i] = is_e ? feq[i]
                              : fmaf(0.5f * wp, feq[i] - fhn[i] + feb[i] - fhb[i],
                                     fmaf(0.5f * wm, feq[i] - feb[i] - fhn[i] + fhb[i], fhn[i] + Fin[i]));
            }
        }
        ep_store_p<FP, QQ, ODD>(fhn, a.fi, N, n, nb);  // streaming part 1, sim.cl:757
        }
    }

    if (MHD) {
        if (a.ext & ION_EXT_DETERMINISTIC) {  // warp-uniform; the charge deposit is Q[n], already stored
            if (dep.ind != 0xFFFFFFFFu) {
                a.lod_u[n] = dep.ux;
                a.lod_u[a.N + n] = dep.uy;
                a.lod_u[2ull * a.N + n] = dep.uz;
            }
        } else {
            lod_deposit_warp(a, dep, lod_off);
        }
    }
}

// update_fields, sim.cl:834-859
template <int VS, int FP>
__global__ void __launch_bounds__(SC_BLOCK_MAX) k_update_fields(const __grid_constant__ KArgs a, const uint64_t t) {
    constexpr int QQ = VSet<VS>::Q;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= a.nx || is_halo(a, x, y, z)) return;
    const Cell c = make_cell(a, x, y, z);
    const uint32_t n = c.n;
    if ((a.flags[n] & ION_TYPE_BO) == ION_TYPE_S) return;
    float fhn[QQ];
    ep_load<FP, QQ>(fhn, a.fi, a.N, n, t & 1ull, [&](int i) { return neighbor<VS>(c, i); });
    float rhon, uxn, uyn, uzn;
    rho_u<VS>(fhn, rhon, uxn, uyn, uzn);
    a.rho[n] = rhon;
    a.u[n] = clampf(uxn, -ION_DEF_C, ION_DEF_C);
    a.u[a.N + n] = clampf(uyn, -ION_DEF_C, ION_DEF_C);
    a.u[2ull * a.N + n] = clampf(uzn, -ION_DEF_C, ION_DEF_C);
}

// initialize, sim.cl:760-832.  The neighbour-flag scan of sim.cl:782-794 has no effect on the result (the
// inner `flagsn_bo==TYPE_S` test at :795 is always true inside the solid branch), so flags of neighbours are
// not read here; the visible end state is identical.
template <int VS, int FP, bool MHD>
__global__ void __launch_bounds__(SC_BLOCK_MAX) k_initialize(const __grid_constant__ KArgs a) {
    constexpr int QQ = VSet<VS>::Q;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= a.nx || is_halo(a, x, y, z)) return;
    const Cell c = make_cell(a, x, y, z);
    const uint32_t n = c.n;
    const uint64_t N = a.N;
    const uint8_t bo = a.flags[n] & ION_TYPE_BO;
    float ux = a.u[n], uy = a.u[N + n], uz = a.u[2ull * N + n];
    float Qn = MHD ? a.Q[n] : 0.0f;
    if (bo == ION_TYPE_S) {  // sim.cl:784-803
        ux = uy = uz = 0.0f;
        a.u[n] = 0.0f;
        a.u[N + n] = 0.0f;
        a.u[2ull * N + n] = 0.0f;
        if (MHD) {
            Qn = 0.0f;
            a.Q[n] = 0.0f;
        }
    }
    auto nb = [&](int i) { return neighbor<VS>(c, i); };
    float feq[QQ];
    f_eq<VS>(a.rho[n], ux, uy, uz, feq);
    ep_store<FP, QQ>(feq, a.fi, N, n, 1ull, nb);  // sim.cl:806
    if (MHD) {
        float qeq[7];
        a_eq(Qn, ux, uy, uz, qeq);  // sim.cl:811-814
        auto nb7 = [&](int i) { return neighbor7(c, i); };
        ep_store<FP, 7>(qeq, a.fqi, N, n, 1ull, nb7);
        a.B_dyn[n] = a.B_stat[n];  // sim.cl:816-821
        a.B_dyn[N + n] = a.B_stat[N + n];
        a.B_dyn[2ull * N + n] = a.B_stat[2ull * N + n];
        a.E_dyn[n] = a.E_stat[n];
        a.E_dyn[N + n] = a.E_stat[N + n];
        a.E_dyn[2ull * N + n] = a.E_stat[2ull * N + n];
        f_eq<VS>(0.0f, ux, uy, uz, feq);  // sim.cl:823-825 (quirk Q10: electron gas starts at rho_e = 0)
        ep_store<FP, QQ>(feq, a.ei, N, n, 1ull, nb);
        if (a.ext & ION_EXT_SUBGRID_ECR) {  // sim.cl:827-831
            a_eq(a.Et[n], ux, uy, uz, qeq);
            ep_store<FP, 7>(qeq, a.eti, N, n, 1ull, nb7);
        }
    }
}

inline dim3 cell_grid(const KArgs& a, unsigned& block, unsigned cap = SC_BLOCK_MAX) {
    unsigned b = ((a.nx + 31u) / 32u) * 32u;
    if (b > cap) b = cap;
    block = b;
    return dim3((a.nx + b - 1u) / b, a.ny, a.nz);
}
inline dim3 sc_grid(const KArgs& a, unsigned& block) {  // stream_collide: ION_SC_BLOCK threads, env ION_SC_BLOCK may lower it (A/B timing)
    static const unsigned env = getenv("ION_SC_BLOCK") ? (unsigned)atoi(getenv("ION_SC_BLOCK")) : 0u;
    const unsigned cap = (env >= 32u && env <= (unsigned)ION_SC_BLOCK) ? env : (unsigned)ION_SC_BLOCK;
    dim3 g = cell_grid(a, block, cap);
    if (a.z_cnt) g.z = a.z_cnt;  // a range of z layers (KArgs::z_off, z_cnt)
    return g;
}

// per-velocity-set launchers (one translation unit each, see sc_d*.cu)
template <int VS>
cudaError_t launch_stream_collide_vs(const KArgs& a, int fp, bool mhd, bool trt, bool ecr, uint64_t t, float fx, float fy, float fz,
                                     cudaStream_t s);
template <int VS> cudaError_t launch_update_fields_vs(const KArgs& a, int fp, uint64_t t, cudaStream_t s);
template <int VS> cudaError_t launch_initialize_vs(const KArgs& a, int fp, bool mhd, cudaStream_t s);

// four-cells-per-thread vector variant of the plain kernel (stream_collide_v4.cuh); returns false when it does not apply
template <int VS>
inline bool launch_stream_collide_v4(const KArgs& a, int fp, bool trt, uint64_t t, float fx, float fy, float fz, cudaStream_t s);

#define ION_SC_CASE(FPV, MHDV, TRTV, ECRV)                                                                 \
    if (fp == FPV && mhd == MHDV && trt == TRTV && ecr == ECRV) {                                          \
        if (t & 1ull) k_stream_collide<VS, FPV, MHDV, TRTV, ECRV, true><<<grid, block, 0, s>>>(a, fx, fy, fz);  \
        else k_stream_collide<VS, FPV, MHDV, TRTV, ECRV, false><<<grid, block, 0, s>>>(a, fx, fy, fz);         \
        return cudaGetLastError();                                                                         \
    }

#define ION_DEFINE_VS_LAUNCHERS(VSV, ALLOW_MHD)                                                                      \
    template <>                                                                                                      \
    cudaError_t launch_stream_collide_vs<VSV>(const KArgs& a, int fp, bool mhd, bool trt, bool ecr, uint64_t t,      \
                                              float fx, float fy, float fz, cudaStream_t s) {                        \
        constexpr int VS = VSV;                                                                                      \
        if (!mhd && !ecr && launch_stream_collide_v4<VSV>(a, fp, trt, t, fx, fy, fz, s)) return cudaGetLastError();  \
        unsigned block;                                                                                              \
        const dim3 grid = sc_grid(a, block);                                                                         \
        ION_SC_CASE(ION_FP32, false, false, false)                                                                   \
        ION_SC_CASE(ION_FP32, false, true, false)                                                                    \
        ION_SC_CASE(ION_FP16S, false, false, false)                                                                  \
        ION_SC_CASE(ION_FP16S, false, true, false)                                                                   \
        ION_SC_CASE(ION_FP16C, false, false, false)                                                                  \
        ION_SC_CASE(ION_FP16C, false, true, false)                                                                   \
        if (ALLOW_MHD) {                                                                                             \
            ION_SC_CASE(ION_FP32, (bool)ALLOW_MHD, false, false)                                                     \
            ION_SC_CASE(ION_FP32, (bool)ALLOW_MHD, true, false)                                                      \
            ION_SC_CASE(ION_FP16S, (bool)ALLOW_MHD, false, false)                                                    \
            ION_SC_CASE(ION_FP16S, (bool)ALLOW_MHD, true, false)                                                     \
            ION_SC_CASE(ION_FP16C, (bool)ALLOW_MHD, false, false)                                                    \
            ION_SC_CASE(ION_FP16C, (bool)ALLOW_MHD, true, false)                                                     \
            ION_SC_CASE(ION_FP32, (bool)ALLOW_MHD, false, (bool)ALLOW_MHD)                                           \
            ION_SC_CASE(ION_FP32, (bool)ALLOW_MHD, true, (bool)ALLOW_MHD)                                            \
            ION_SC_CASE(ION_FP16S, (bool)ALLOW_MHD, false, (bool)ALLOW_MHD)                                          \
            ION_SC_CASE(ION_FP16S, (bool)ALLOW_MHD, true, (bool)ALLOW_MHD)                                           \
            ION_SC_CASE(ION_FP16C, (bool)ALLOW_MHD, false, (bool)ALLOW_MHD)                                          \
            ION_SC_CASE(ION_FP16C, (bool)ALLOW_MHD, true, (bool)ALLOW_MHD)                                           \
        }                                                                                                            \
        return cudaErrorInvalidValue;                                                                                \
    }                                                                                                                \
    template <> cudaError_t launch_update_fields_vs<VSV>(const KArgs& a, int fp, uint64_t t, cudaStream_t s) {       \
        unsigned block;                                                                                              \
        const dim3 grid = cell_grid(a, block);                                                                       \
        if (fp == ION_FP32) k_update_fields<VSV, ION_FP32><<<grid, block, 0, s>>>(a, t);                             \
        else if (fp == ION_FP16S) k_update_fields<VSV, ION_FP16S><<<grid, block, 0, s>>>(a, t);                      \
        else k_update_fields<VSV, ION_FP16C><<<grid, block, 0, s>>>(a, t);                                           \
        return cudaGetLastError();                                                                                   \
    }                                                                                                                \
    template <> cudaError_t launch_initialize_vs<VSV>(const KArgs& a, int fp, bool mhd, cudaStream_t s) {            \
        unsigned block;                                                                                              \
        const dim3 grid = cell_grid(a, block);                                                                       \
        if (!mhd) {                                                                                                  \
            if (fp == ION_FP32) k_initialize<VSV, ION_FP32, false><<<grid, block, 0, s>>>(a);                        \
            else if (fp == ION_FP16S) k_initialize<VSV, ION_FP16S, false><<<grid, block, 0, s>>>(a);                 \
            else k_initialize<VSV, ION_FP16C, false><<<grid, block, 0, s>>>(a);                                      \
        } else if (ALLOW_MHD) {                                                                                      \
            if (fp == ION_FP32) k_initialize<VSV, ION_FP32, (bool)ALLOW_MHD><<<grid, block, 0, s>>>(a);              \
            else if (fp == ION_FP16S) k_initialize<VSV, ION_FP16S, (bool)ALLOW_MHD><<<grid, block, 0, s>>>(a);       \
            else k_initialize<VSV, ION_FP16C, (bool)ALLOW_MHD><<<grid, block, 0, s>>>(a);                            \
        } else {                                                                                                     \
            return cudaErrorInvalidValue;                                                                            \
        }                                                                                                            \
        return cudaGetLastError();                                                                                   \
    }

}  // namespace ion
