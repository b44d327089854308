// transfer.cu -- halo pack/unpack kernels (the device half of Lbm::communicate_field).
//
// Behaviour: /root/reference/src/kernels/sim_kernels.cl ("sim.cl") :1006-1147.  The packed layout
// transfer_buffer[b*A + a] (b = transferred DDF, a = face cell) and the Esoteric-Pull slot rules of
// sim.cl:1068/1077 are reproduced bit-exactly, so a packed face can be handed to any neighbour (same GPU,
// peer GPU over NVLink, or the reference itself).
//
// B200 mapping: one thread per face cell; for z faces (the multi-GPU slab split) `a` runs along x then y, so
// every one of the T transferred planes is read and written as contiguous, coalesced rows.  Values are moved
// as raw storage words (u16/f32) without decode/encode, like the reference's fpxx_copy.
#include "lattice.cuh"

namespace ion {

// per-face transferred directions, sim.cl:1029-1060
template <int VS> __device__ __forceinline__ int index_transfer(int side_i) {
    if (VS == ION_D2Q9) {
        constexpr int8_t t[12] = {1, 5, 7, 2, 6, 8, 3, 5, 8, 4, 6, 7};
        return t[side_i];
    } else if (VS == ION_D3Q15) {
        constexpr int8_t t[30] = {1, 7, 14, 9, 11, 2, 8, 13, 10, 12, 3, 7, 12, 9, 13, 4, 8, 11, 10, 14, 5, 7, 10, 11, 13, 6, 8, 9, 12, 14};
        return t[side_i];
    } else if (VS == ION_D3Q19) {
        constexpr int8_t t[30] = {1, 7, 13, 9, 15, 2, 8, 14, 10, 16, 3, 7, 14, 11, 17, 4, 8, 13, 12, 18, 5, 9, 16, 11, 18, 6, 10, 15, 12, 17};
        return t[side_i];
    } else {
        constexpr int8_t t[54] = {1, 7, 13, 9,  15, 19, 26, 21, 23, 2, 8,  14, 10, 16, 20, 25, 22, 24,
                                  3, 7, 14, 11, 17, 19, 24, 21, 25, 4, 8,  13, 12, 18, 20, 23, 22, 26,
                                  5, 9, 16, 11, 18, 19, 22, 23, 25, 6, 10, 15, 12, 17, 20, 21, 24, 26};
        return t[side_i];
    }
}

// index_extract_p/m, index_insert_p/m (sim.cl:1012-1027): face cell `a` -> coordinates at `layer`
__device__ __forceinline__ void face_cell(const KArgs& p, uint32_t a, uint32_t direction, uint32_t layer, uint32_t& x,
                                          uint32_t& y, uint32_t& z) {
    if (direction == 0u) { x = layer; y = a % p.ny; z = a / p.ny; }
    else if (direction == 1u) { x = a / p.nz; y = layer; z = a % p.nz; }
    else { x = a % p.nx; y = a / p.nx; z = layer; }
}
__device__ __forceinline__ uint32_t axis_size(const KArgs& p, uint32_t direction) {
    return direction == 0u ? p.nx : direction == 1u ? p.ny : p.nz;
}
__host__ __device__ inline uint32_t face_area(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t direction) {
    return direction == 0u ? ny * nz : direction == 1u ? nx * nz : nx * ny;  // get_area, domain.rs:475-482
}

// transfer_extract_fi / transfer__insert_fi (sim.cl:1063-1092); also bound to `ei` (domain.rs:340-353)
template <int VS, typename S, bool INSERT>
__global__ void k_transfer_fi(const __grid_constant__ KArgs p, S* __restrict__ fi, const uint32_t direction, const uint64_t t) {
    constexpr int T = VSet<VS>::T;
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t A = face_area(p.nx, p.ny, p.nz, direction);
    if (a >= A) return;
    const uint64_t todd = t & 1ull;
    const uint32_t L = axis_size(p, direction);
#pragma unroll
    for (int side = 0; side < 2; side++) {  // 0: positive face (transfer_buffer_p), 1: negative face (transfer_buffer_m)
        const uint32_t layer = INSERT ? (side == 0 ? L - 1u : 0u) : (side == 0 ? L - 2u : 1u);
        uint32_t x, y, z;
        face_cell(p, a, direction, layer, x, y, z);
        const Cell c = make_cell(p, x, y, z);
        S* buf = reinterpret_cast<S*>(side == 0 ? p.transfer_p : p.transfer_m);
#pragma unroll
        for (int b = 0; b < T; b++) {
            // the table lookup is per (direction, side): resolve with a small switch so that it stays in registers
            int i = 0;
#pragma unroll
            for (int dsel = 0; dsel < VSet<VS>::DIM; dsel++)
                if ((uint32_t)dsel == direction) i = index_transfer<VS>((2 * dsel + side) * T + b);
            // neighbours j[i] (extract, i odd) / j[i-1] (insert, i even) -- all needed j's are odd-indexed
            uint32_t jn = 0u;
            const int want = INSERT ? i - 1 : i;
#pragma unroll
            for (int k = 1; k < VSet<VS>::Q; k += 2)
                if (k == want) jn = neighbor<VS>(c, k);
            uint64_t index;
            if (!INSERT) {  // sim.cl:1068
                const uint32_t cell = (i & 1) ? jn : c.n;
                const uint32_t slot = todd ? ((i & 1) ? i + 1 : i - 1) : i;
                index = (uint64_t)slot * p.N + cell;
                buf[(uint64_t)b * A + a] = fi[index];
            } else {  // sim.cl:1077
                const uint32_t cell = (i & 1) ? c.n : jn;
                const uint32_t slot = todd ? i : ((i & 1) ? i + 1 : i - 1);
                index = (uint64_t)slot * p.N + cell;
                fi[index] = buf[(uint64_t)b * A + a];
            }
        }
    }
}

// transfer_extract_fqi / transfer__insert_fqi (sim.cl:1122-1147): one DDF of the D3Q7 lattice per face
template <typename S, bool INSERT>
__global__ void k_transfer_fqi(const __grid_constant__ KArgs p, S* __restrict__ fqi, const uint32_t direction, const uint64_t t) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t A = face_area(p.nx, p.ny, p.nz, direction);
    if (a >= A) return;
    const uint64_t todd = t & 1ull;
    const uint32_t L = axis_size(p, direction);
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const uint32_t layer = INSERT ? (side == 0 ? L - 1u : 0u) : (side == 0 ? L - 2u : 1u);
        uint32_t x, y, z;
        face_cell(p, a, direction, layer, x, y, z);
        const Cell c = make_cell(p, x, y, z);
        S* buf = reinterpret_cast<S*>(side == 0 ? p.transfer_p : p.transfer_m);
        const uint32_t i = 2u * direction + (uint32_t)side + 1u;  // sim.cl:1125
        uint64_t index;
        if (!INSERT) {
            const uint32_t cell = (i & 1u) ? neighbor7(c, (int)i) : c.n;
            const uint32_t slot = todd ? ((i & 1u) ? i + 1u : i - 1u) : i;
            index = (uint64_t)slot * p.N + cell;
            buf[a] = fqi[index];
        } else {
            const uint32_t cell = (i & 1u) ? c.n : neighbor7(c, (int)i - 1);
            const uint32_t slot = todd ? i : ((i & 1u) ? i + 1u : i - 1u);
            index = (uint64_t)slot * p.N + cell;
            fqi[index] = buf[a];
        }
    }
}

// transfer_extract_rho_u_flags / transfer__insert_rho_u_flags (sim.cl:1094-1119): 17 bytes per face cell
template <bool INSERT>
__global__ void k_transfer_rho_u_flags(const __grid_constant__ KArgs p, const uint32_t direction) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t A = face_area(p.nx, p.ny, p.nz, direction);
    if (a >= A) return;
    const uint32_t L = axis_size(p, direction);
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const uint32_t layer = INSERT ? (side == 0 ? L - 1u : 0u) : (side == 0 ? L - 2u : 1u);
        uint32_t x, y, z;
        face_cell(p, a, direction, layer, x, y, z);
        const uint64_t n = x + (y + (uint64_t)z * p.ny) * p.nx;
        uint8_t* raw = side == 0 ? p.transfer_p : p.transfer_m;
        float* bf = reinterpret_cast<float*>(raw);
        if (!INSERT) {
            bf[a] = p.rho[n];
            bf[A + a] = p.u[n];
            bf[2ull * A + a] = p.u[p.N + n];
            bf[3ull * A + a] = p.u[2ull * p.N + n];
            raw[16ull * A + a] = p.flags[n];
        } else {
            p.rho[n] = bf[a];
            p.u[n] = bf[A + a];
            p.u[p.N + n] = bf[2ull * A + a];
            p.u[2ull * p.N + n] = bf[3ull * A + a];
            p.flags[n] = raw[16ull * A + a];
        }
    }
}

template <int VS, bool INSERT>
static cudaError_t launch_fi_vs(const KArgs& p, void* field, bool fp32, uint32_t direction, uint64_t t, cudaStream_t s) {
    const uint32_t A = face_area(p.nx, p.ny, p.nz, direction);
    const unsigned grid = (A + 127u) / 128u;
    if (fp32) k_transfer_fi<VS, float, INSERT><<<grid, 128, 0, s>>>(p, reinterpret_cast<float*>(field), direction, t);
    else k_transfer_fi<VS, uint16_t, INSERT><<<grid, 128, 0, s>>>(p, reinterpret_cast<uint16_t*>(field), direction, t);
    return cudaGetLastError();
}

// transfer_field: IonTransferField; insert: 0 extract, 1 insert
cudaError_t launch_transfer(const KArgs& p, int vs, int fp, int transfer_field, int insert, uint32_t direction, uint64_t t,
                            cudaStream_t s) {
    const bool fp32 = fp == ION_FP32;
    const uint32_t A = face_area(p.nx, p.ny, p.nz, direction);
    const unsigned grid = (A + 127u) / 128u;
    if (transfer_field == ION_TRANSFER_FI || transfer_field == ION_TRANSFER_EI) {
        void* f = transfer_field == ION_TRANSFER_FI ? p.fi : p.ei;
        switch (vs * 2 + insert) {
            case ION_D2Q9 * 2 + 0: return launch_fi_vs<ION_D2Q9, false>(p, f, fp32, direction, t, s);
            case ION_D2Q9 * 2 + 1: return launch_fi_vs<ION_D2Q9, true>(p, f, fp32, direction, t, s);
            case ION_D3Q15 * 2 + 0: return launch_fi_vs<ION_D3Q15, false>(p, f, fp32, direction, t, s);
            case ION_D3Q15 * 2 + 1: return launch_fi_vs<ION_D3Q15, true>(p, f, fp32, direction, t, s);
            case ION_D3Q19 * 2 + 0: return launch_fi_vs<ION_D3Q19, false>(p, f, fp32, direction, t, s);
            case ION_D3Q19 * 2 + 1: return launch_fi_vs<ION_D3Q19, true>(p, f, fp32, direction, t, s);
            case ION_D3Q27 * 2 + 0: return launch_fi_vs<ION_D3Q27, false>(p, f, fp32, direction, t, s);
            case ION_D3Q27 * 2 + 1: return launch_fi_vs<ION_D3Q27, true>(p, f, fp32, direction, t, s);
        }
        return cudaErrorInvalidValue;
    }
    if (transfer_field == ION_TRANSFER_QI) {
        if (fp32) {
            if (insert) k_transfer_fqi<float, true><<<grid, 128, 0, s>>>(p, reinterpret_cast<float*>(p.fqi), direction, t);
            else k_transfer_fqi<float, false><<<grid, 128, 0, s>>>(p, reinterpret_cast<float*>(p.fqi), direction, t);
        } else {
            if (insert) k_transfer_fqi<uint16_t, true><<<grid, 128, 0, s>>>(p, reinterpret_cast<uint16_t*>(p.fqi), direction, t);
            else k_transfer_fqi<uint16_t, false><<<grid, 128, 0, s>>>(p, reinterpret_cast<uint16_t*>(p.fqi), direction, t);
        }
        return cudaGetLastError();
    }
    if (insert) k_transfer_rho_u_flags<true><<<grid, 128, 0, s>>>(p, direction);
    else k_transfer_rho_u_flags<false><<<grid, 128, 0, s>>>(p, direction);
    return cudaGetLastError();
}

}  // namespace ion
