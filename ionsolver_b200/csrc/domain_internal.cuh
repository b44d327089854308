// domain_internal.cuh -- private layout of ion_domain_t and the error helpers shared by api.cu and comm.cu.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "lattice.cuh"

namespace ion { struct EbFftPlan; }
struct ion_domain {
    IonParams params;
    int device;
    cudaStream_t stream;
    void* buf[ION_FIELD_COUNT];
    size_t bytes[ION_FIELD_COUNT];
    ion::KArgs k;
    void* lod_sources;    // scratch of update_e_b_dynamic
    uint32_t* cp_counts;  // scratch of the precompute compaction
    void* alt_p;          // spare transfer buffers: receive side of peer copies / NCCL (see ion_exchange_transfer)
    void* alt_m;
    float* lod_u;         // deterministic mode: per-cell velocity deposits (3N floats)
    float* lod_rep;       // default mode: ION_LOD_REPLICAS private copies of the finest LOD level (deposit target of stream_collide)
    float* lod_gather;    // world * n_lod_own * 4 floats, allocated on first ion_comm_exchange_lods
    cudaEvent_t ev;       // reusable ordering event (timing disabled)
    // halo stream: between ion_halo_fork and ion_halo_join the transfer kernels and the face exchange of this domain run
    // here, concurrently with whatever is queued on `stream` (update_e_b_dynamic does not touch the DDFs)
    cudaStream_t halo_stream;
    cudaEvent_t ev_fork, ev_join;
    bool halo_active;
    float ecrf;
    bool deterministic;   // ION_EXT_DETERMINISTIC: reference-ordered LOD sums and reference arithmetic in update_e_b_dynamic
    ion::EbFftPlan* eb_plan;  // polyphase-FFT update_e_b_dynamic (eb_fft.cu): static kernel spectra, built on first use
    bool eb_plan_tried;
    int precompute_mode;      // ion_domain_set_precompute_mode: 0 reference order, 1 rsqrt / FMA direct sum, 2 FFT convolution
};

#ifndef ION_LOD_REPLICAS
#define ION_LOD_REPLICAS 32u  // power of two
#endif

namespace ion {
extern std::atomic<uint64_t> g_launches;
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void set_transfer_ptrs(ion_domain* d);  // refresh KArgs after the current/spare transfer buffers were flipped
inline cudaStream_t xfer_stream(const ion_domain* d) { return d->halo_active ? d->halo_stream : d->stream; }
}  // namespace ion

#define ION_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t ion_e_ = (call);                                    \
        if (ion_e_ != cudaSuccess) return ion::cuda_fail(ion_e_, #call); \
    } while (0)
