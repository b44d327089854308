/* ionsolver_b200.h -- C ABI of libionsolver_b200.so
 *
 * Drop-in boundary for IonSolver's extended-LBM MHD time step on NVIDIA B200 (sm_100a).
 * This header replaces what the reference's Rust host obtains from the `ocl` / `ocl-macros` crates and
 * /root/reference/src/opencl.rs: one opaque device-side domain object per LbmDomain, its buffers, and one
 * enqueue function per OpenCL kernel launch site.  Every entry point cites the reference call site it
 * replaces (paths relative to /root/reference).  The cgo-style binding a maintainer adds on the Rust side is
 * shown in INTEGRATION.md.
 *
 * Contract (same as an in-order OpenCL command queue, src/lbm/domain.rs:132-135):
 *   - every ion_enqueue_* call is asynchronous and ordered on the domain's CUDA stream; ion_finish blocks;
 *   - ion_buffer_read / ion_buffer_write are blocking with respect to the host pointer (like bread!/bwrite!);
 *   - every function returns 0 on success or a non-zero IonStatus / cudaError_t value; the text of the last
 *     failure on the calling thread is available from ion_last_error_string();
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x, ion_domain_create fails.
 * Thread-compatibility: one thread per handle at a time; distinct handles may be used concurrently.
 */
#ifndef IONSOLVER_B200_H
#define IONSOLVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ION_ABI_VERSION 1

#if defined(__GNUC__)
#define ION_API __attribute__((visibility("default")))
#else
#define ION_API
#endif

/* enum discriminants are the reference's wire values (src/lbm/types.rs:17-25,55-60,76-82) */
/* D3Q27 uses the canonical weights 8/27, 2/27, 1/54, 1/216: the reference emits no DEF_WC for D3Q27 and its kernels
 * therefore never compiled for that set (domain.rs:766-768 vs sim_kernels.cl:192) */
enum IonVelocitySet { ION_D2Q9 = 0, ION_D3Q15 = 1, ION_D3Q19 = 2, ION_D3Q27 = 3 };
enum IonRelaxationTime { ION_SRT = 0, ION_TRT = 1 };
enum IonFloatType { ION_FP16S = 0, ION_FP16C = 1, ION_FP32 = 2 };
/* src/lbm/types.rs:105-111 */
enum IonTransferField { ION_TRANSFER_FI = 0, ION_TRANSFER_RHO_U_FLAGS = 1, ION_TRANSFER_EI = 2, ION_TRANSFER_QI = 3 };

/* extension switches: the `#define`s emitted at src/lbm/domain.rs:834-856 */
enum IonExt {
    ION_EXT_EQUILIBRIUM_BOUNDARIES = 1u << 0,
    ION_EXT_VOLUME_FORCE = 1u << 1,
    ION_EXT_FORCE_FIELD = 1u << 2,
    ION_EXT_MAGNETO_HYDRO = 1u << 3,
    ION_EXT_SUBGRID_ECR = 1u << 4,
    ION_EXT_UPDATE_FIELDS = 1u << 5, /* graphics_active => UPDATE_FIELDS, domain.rs:856 */
    /* No reference equivalent.  The reference's LOD deposit uses float atomics (sim_kernels.cl:673-676), so its E/B
     * fields depend on the execution order of the device and are not reproducible run to run.  With this bit the LOD
     * sums are formed in ascending cell order and update_e_b_dynamic uses the reference's exact arithmetic (sqrt,
     * cube, IEEE division, unfused sums) in the reference's loop order: every field is then bit-identical to the
     * reference kernels executed sequentially.  Needs nx, ny, nz (halo-inclusive) divisible by 2^lod_depth. */
    ION_EXT_DETERMINISTIC = 1u << 6
};

/* flag bits, src/lbm/domain.rs:819-828 (runtime values, not the IDE placeholders of sim_kernels.cl:45-54) */
#define ION_TYPE_S 0x01
#define ION_TYPE_E 0x02
#define ION_TYPE_C 0x04
#define ION_TYPE_F 0x08
#define ION_TYPE_M 0x10
#define ION_TYPE_BO 0x1F

/* Device buffers of one domain, src/lbm/domain.rs:45-74 / SURVEY.md appendix A */
enum IonField {
    ION_FIELD_FI = 0,     /* Q*n  f32|u16 */
    ION_FIELD_RHO = 1,    /* n    f32, initial 1.0 */
    ION_FIELD_U = 2,      /* 3n   f32 (x,y,z planes) */
    ION_FIELD_FLAGS = 3,  /* n    u8 */
    ION_FIELD_F = 4,      /* 3n   f32, ext_force_field */
    ION_FIELD_E_STAT = 5, /* 3n   f32, MHD */
    ION_FIELD_B_STAT = 6,
    ION_FIELD_E_DYN = 7,
    ION_FIELD_B_DYN = 8,
    ION_FIELD_FQI = 9,     /* 7n   f32|u16, MHD */
    ION_FIELD_EI = 10,     /* Q*n  f32|u16, MHD */
    ION_FIELD_Q = 11,      /* n    f32, MHD */
    ION_FIELD_QU_LOD = 12, /* 4*n_lod f32 AoS (q,ux,uy,uz), MHD */
    ION_FIELD_E_VAR = 13,  /* 3n   f32, SUBGRID_ECR */
    ION_FIELD_ETI = 14,    /* 7n   f32|u16, SUBGRID_ECR */
    ION_FIELD_ET = 15,     /* n    f32, SUBGRID_ECR */
    ION_FIELD_TRANSFER_P = 16, /* A_max*max(17,T*s) bytes, layout [b*A+a] (sim_kernels.cl:1069) */
    ION_FIELD_TRANSFER_M = 17,
    ION_FIELD_COUNT = 18
};

enum IonStatus {
    ION_OK = 0,
    ION_ERR_INVALID = 10001,     /* bad argument / params */
    ION_ERR_UNSUPPORTED = 10002, /* configuration the reference cannot build either (e.g. MHD on D2Q9) */
    ION_ERR_NO_DEVICE = 10003,   /* no sm_100 device: there is no CPU fallback */
    ION_ERR_ABSENT = 10004,      /* buffer not allocated for this configuration (Option::None in domain.rs) */
    ION_ERR_RANGE = 10005        /* offset/size outside the buffer */
};

/* The compile-time parameters the reference injects as text (get_device_defines, src/lbm/domain.rs:736-858),
 * as a plain struct.  Float members carry the exact f32 the reference prints with `{:?}`. */
typedef struct IonParams {
    uint32_t abi_version; /* ION_ABI_VERSION */
    uint32_t nx, ny, nz;  /* DEF_NX/NY/NZ: local domain size incl. halo layers (domain.rs:91-93) */
    uint32_t dx, dy, dz;  /* DEF_DX/DY/DZ: number of domains per axis */
    uint32_t di;          /* DEF_DI: index of this domain */
    int32_t ox, oy, oz;   /* DEF_OX/OY/OZ: signed origin offsets (domain.rs:101-103) */
    uint32_t velocity_set;    /* IonVelocitySet */
    uint32_t relaxation_time; /* IonRelaxationTime */
    uint32_t float_type;      /* IonFloatType */
    uint32_t ext;             /* IonExt bit mask */
    float w;                  /* DEF_W = 1/(3 nu + 1/2) (domain.rs:808) */
    float ke, kmu, kmu0, kkge, kimg, kvev, kme; /* DEF_KE..DEF_KME (domain.rs:838-844) */
    float wq;                 /* DEF_WQ (domain.rs:848) */
    float kkbme, keabs;       /* DEF_KKBME, DEF_KEABS (domain.rs:852-853) */
    uint32_t lod_depth;       /* DEF_LOD_DEPTH */
    uint32_t n_lod;           /* DEF_NUM_LOD */
    uint32_t n_lod_own;       /* DEF_NUM_LOD_OWN */
} IonParams;

typedef struct ion_domain ion_domain_t;

/* library / device ------------------------------------------------------------------------------------- */
/* replaces opencl::device_selection (src/opencl.rs:9-63): number of usable sm_100 devices */
ION_API int ion_device_count(int* count);
ION_API const char* ion_last_error_string(void);
ION_API uint32_t ion_abi_version(void);

/* domain lifetime: LbmDomain::new program build + buffer allocation + kernel binding (domain.rs:88-409) */
ION_API int ion_domain_create(const IonParams* params, int device, ion_domain_t** out);
ION_API int ion_domain_destroy(ion_domain_t* dom); /* Drop of LbmDomain */
ION_API int ion_domain_params(const ion_domain_t* dom, IonParams* out);

/* buffers: buffer!/bwrite!/bread! and Buffer::read/write with .offset()/.len() (ocl-macros; e.g. domain.rs:504-511,
 * mod.rs:460-462, setup.rs:180-302, file.rs:118-268).  Offsets and sizes are in BYTES. */
ION_API int ion_buffer_size(const ion_domain_t* dom, int field, size_t* bytes);
ION_API int ion_buffer_write(ion_domain_t* dom, int field, const void* host, size_t offset_bytes, size_t bytes);
ION_API int ion_buffer_read(ion_domain_t* dom, int field, void* host, size_t offset_bytes, size_t bytes);
/* raw device pointer (for zero-copy interop with the caller's CUDA allocator / torch tensors); NULL if absent */
ION_API int ion_buffer_device_ptr(const ion_domain_t* dom, int field, void** dptr);
/* device-to-device copy between two domains' buffers on dst's stream; the single-node replacement for the
 * host-staged std::ptr::swap exchange of mod.rs:380-384 and the LOD slice copy of mod.rs:460-462 */
/* save-and-load: buffer fields[i] is read to host_out[i], then overwritten from host_in[i]; downloads and uploads overlap on two
 * streams (file.rs:221-268 followed by :118-152 for the next state).  Pinned host memory; blocks until done. */
ION_API int ion_buffer_swap(ion_domain_t* dom, int n, const int* fields, void* const* host_out, const void* const* host_in, const size_t* bytes);
ION_API int ion_buffer_copy(ion_domain_t* dst, int dst_field, size_t dst_offset_bytes, ion_domain_t* src, int src_field,
                    size_t src_offset_bytes, size_t bytes);

/* Slice read-back: the values of one lattice plane of this domain, gathered on the device and copied to host_out.
 * The reference only DRAWS slices (graphics_field_slice, graphics_kernels.cl:669-706, selected by GraphicsConfig::slice_mode /
 * slice_x/y/z, graphics.rs:124-130); this is the data half of that view, in the same cell enumeration:
 * direction 0: a -> (index, a % ny, a / ny); 1: (a / nz, index, a % nz); 2: (a % nx, a / nx, index); halo layers included.
 * field: RHO, Q, ET (scalars), FLAGS (as float), U, F, E_STAT, B_STAT, E_DYN, B_DYN, E_VAR (component 0,1,2 or 3 = length()). */
ION_API int ion_read_slice(ion_domain_t* dom, int field, int component, uint32_t direction, uint32_t index, float* host_out);

/* kernels: one per enqueue site ---------------------------------------------------------------------------- */
ION_API int ion_enqueue_initialize(ion_domain_t* dom);                                          /* domain.rs:412-416 (ends with finish) */
ION_API int ion_enqueue_stream_collide(ion_domain_t* dom, uint64_t t, float fx, float fy, float fz); /* domain.rs:419-428 */
/* the same kernel on the z layers [z_begin, z_end) only -- boundary-layer-first scheduling: launch the two layers next to the halos,
 * start the halo exchange on the halo stream (ion_halo_fork), launch the interior; finish != 0 on the LAST range of a step (LOD fold) */
ION_API int ion_enqueue_stream_collide_range(ion_domain_t* dom, uint64_t t, float fx, float fy, float fz, uint32_t z_begin, uint32_t z_end, int finish);
ION_API int ion_enqueue_update_fields(ion_domain_t* dom, uint64_t t, float fx, float fy, float fz);  /* domain.rs:432-441 */
ION_API int ion_enqueue_update_e_b_dyn(ion_domain_t* dom);                                      /* domain.rs:443-451 */
/* no reference counterpart: 0 (default) = psi_from_mesh with the reference's arithmetic and summation order (B_stat bit-identical to the
 * reference build), 1 = the same sum with rsqrt / fused multiply-adds, four outputs per thread (~5x faster, psi within ~1e-6 relative L2),
 * 2 = psi_from_mesh and static_e_from_mesh as zero-padded FFT convolutions of the source cells with d/|d|^3 (cuFFT transforms, loaded at
 * run time; O(P^3 log P) instead of O(cells x sources); psi / E_stat within ~1e-5 relative L2; ION_ERR_ABSENT when cuFFT or memory is missing) */
ION_API int ion_domain_set_precompute_mode(ion_domain_t* dom, int mode);
/* no reference counterpart: size of the static kernel spectra of the polyphase-FFT field update (0 = direct kernels in use)
 * and the number of polyphase problems per step; valid after the first ion_enqueue_update_e_b_dyn */
ION_API int ion_domain_eb_fft_info(const ion_domain_t* dom, uint64_t* spectrum_bytes, uint32_t* tasks);
ION_API int ion_enqueue_lod_part_2_gather(ion_domain_t* dom);                                   /* domain.rs:453-462 */
ION_API int ion_enqueue_clear_qu_lod(ion_domain_t* dom);                                        /* domain.rs:464-472 */
/* transfer kernels only (the host read/write halves of domain.rs:484-543 are ion_buffer_read/write/copy on
 * ION_FIELD_TRANSFER_P/M); direction 0,1,2 = x,y,z */
ION_API int ion_enqueue_transfer_extract(ion_domain_t* dom, int transfer_field, uint32_t direction, uint64_t t); /* domain.rs:484-501 */
ION_API int ion_enqueue_transfer_insert(ion_domain_t* dom, int transfer_field, uint32_t direction, uint64_t t);  /* domain.rs:532-542 */
/* LbmDomain::voxelize_mesh_on_device (src/mesh.rs:281-343): p0/p1/p2 are 3*triangles floats, bbu the 7 floats of
 * mesh.rs:288-295 (bit-cast triangle count, bbox-2, bbox+2) */
ION_API int ion_voxelize_mesh(ion_domain_t* dom, const float* p0, const float* p1, const float* p2, uint32_t triangles,
                      const float bbu[7], uint32_t direction, uint8_t flag, float mpc_x, float mpc_y, float mpc_z,
                      uint64_t t);
ION_API int ion_enqueue_precompute_b(ion_domain_t* dom);     /* domain.rs:551-556: psi_from_mesh + static_b_from_mesh */
ION_API int ion_enqueue_precompute_e(ion_domain_t* dom);     /* domain.rs:558-567: static_e_from_mesh -> E_stat */
ION_API int ion_enqueue_precompute_e_ecr(ion_domain_t* dom); /* domain.rs:569-578: static_e_from_mesh -> E_var */
ION_API int ion_domain_set_ecr_freq(ion_domain_t* dom, float ecrf); /* kernel arg "ecrf", domain.rs:292-296 */
ION_API int ion_finish(ion_domain_t* dom);                   /* queue.finish(), mod.rs:275-279 */

/* halo exchange ---------------------------------------------------------------------------------------------
 * The reference stages every face through host memory and swaps the two host vectors of neighbouring domains
 * (std::ptr::swap of transfer_p_host / transfer_m_host, src/lbm/mod.rs:380-384, around the buffer reads/writes of
 * src/lbm/domain.rs:504-531).  Here the faces never leave device memory:
 *   ion_exchange_transfer   one process, two domains: d's transfer_p <-> dp's transfer_m.  Same device: the two
 *                           device pointers are swapped (zero copy, ordered with events); different devices: peer
 *                           copies over NVLink into the partner's spare buffer, then the spare becomes current.
 *   ion_comm_*              one process per GPU (the layout bench.py / torchrun uses): NCCL point-to-point over
 *                           NVLink between ring neighbours, and an all-gather for the LOD pyramids.
 * `bytes` is the face payload (area * bytes_per_cell of mod.rs:410-433); the packed [b*A+a] layout is untouched. */
ION_API int ion_exchange_transfer(ion_domain_t* d, ion_domain_t* dp, size_t bytes);
/* single-process LOD exchange (Lbm::communicate_qu_lods, mod.rs:436-468): copies own-pyramid entries
 * [src_entry, src_entry+entries) of `src` to entry `dst_entry` of `dst` (entries are 4 floats) */
ION_API int ion_copy_lods(ion_domain_t* dst, uint32_t dst_entry, ion_domain_t* src, uint32_t src_entry, uint32_t entries);

/* Overlap (no reference equivalent: the reference finishes every queue between extract and insert, mod.rs:379).  Between
 * ion_halo_fork and ion_halo_join the transfer kernels and face exchanges of `dom` are queued on a second stream that starts
 * after everything queued on the domain so far; ion_halo_join makes the domain's main stream wait for them.  The host layer
 * uses it to run the fi/fqi/ei halo exchange concurrently with update_e_b_dynamic, which does not touch the DDFs. */
ION_API int ion_halo_fork(ion_domain_t* dom);
ION_API int ion_halo_join(ion_domain_t* dom);

/* exchange plan (pure host arithmetic, no GPU needed; used by ion_exchange_transfer's callers, ion_comm_exchange_lods and
 * the host layer, and testable on CPU):
 *   ion_neighbor_domains   ring neighbours of domain d along `axis` (0,1,2): dp = +1 (receives d's transfer_p as its
 *                          transfer_m), dm = -1, periodic (mod.rs:386-404)
 *   ion_lod_exchange_plan  what foreign domain dc contributes to the domain described by `p` (mod.rs:448-465): entries
 *                          [src_entry, src_entry+entries) of dc's own pyramid (level max(0, depth - Chebyshev distance))
 *                          land at entry dst_entry of p's QU_lod (running offset after n_lod_own, ascending dc, self skipped).
 *                          dc == p->di yields entries = 0. */
ION_API int ion_neighbor_domains(uint32_t d_x, uint32_t d_y, uint32_t d_z, uint32_t d, uint32_t axis, uint32_t* dp, uint32_t* dm);
ION_API int ion_lod_exchange_plan(const IonParams* p, uint32_t dc, uint32_t* src_entry, uint32_t* entries, uint32_t* dst_entry);

typedef struct ion_comm ion_comm_t;
#define ION_COMM_ID_BYTES 128
ION_API int ion_comm_unique_id(uint8_t id[ION_COMM_ID_BYTES]); /* rank 0 creates, the launcher broadcasts (torch.distributed) */
ION_API int ion_comm_create(const uint8_t id[ION_COMM_ID_BYTES], int rank, int world, int device, ion_comm_t** out);
ION_API int ion_comm_destroy(ion_comm_t* comm);
/* the p<->m swap with ring neighbours that live in other processes: sends this domain's transfer_p to `rank_p`
 * (where it becomes transfer_m) and transfer_m to `rank_m` (becomes transfer_p), receives the mirror images */
ION_API int ion_comm_exchange_transfer(ion_comm_t* comm, ion_domain_t* d, int rank_p, int rank_m, size_t bytes);
/* all ranks' own LOD pyramids (n_lod_own entries each) gathered into `d`'s scratch, then the level every foreign
 * domain contributes (mod.rs:448-465) is copied into QU_lod after n_lod_own, ascending domain index */
ION_API int ion_comm_exchange_lods(ion_comm_t* comm, ion_domain_t* d);

/* instrumentation (no reference equivalent): kernels launched by this library since load, for bench.py */
ION_API uint64_t ion_kernel_launch_count(void);
/* DDF storage codec on the device (test hook; the macros load/store of domain.rs:773-784, sim_kernels.cl:79-90):
 * dir 0: count floats -> count storage words (u16 or f32), dir 1: storage words -> floats.  Host pointers. */
ION_API int ion_codec_probe(int device, int float_type, int dir, const void* host_in, void* host_out, uint64_t count);
/* FP32 issue-peak probe (FMA/s of the whole device; packed = fma.rn.f32x2): the compute roofline of update_e_b_dynamic */
ION_API int ion_measure_fma_peak(int device, int packed, double* fma_per_s);
/* the CUDA stream of a domain (cudaStream_t as void*), so callers can record CUDA events on it */
ION_API int ion_domain_stream(const ion_domain_t* dom, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* IONSOLVER_B200_H */
