/* ionsolver_b200_host.h -- C ABI of the host layer in libionsolver_b200.so
 *
 * The reference's host is Rust (LbmConfig / Lbm / LbmDomain, setup.rs, mesh.rs, file.rs).  With no Rust toolchain in
 * this environment that surface is restated in C++ (ionsolver_b200/csrc/host/) on top of the device C ABI of
 * ionsolver_b200.h, and exported here as plain C so that tests, bench.py and any other language can drive the same
 * object model: one entry point per public method of the reference (paths relative to /root/reference).
 * A Rust build keeps its own host and binds ionsolver_b200.h directly; this header is the
 * equivalent of `pub struct Lbm` for everyone else.
 *
 * All functions return 0 or an IonStatus / cudaError_t value; ion_last_error_string() has the text.
 */
#ifndef IONSOLVER_B200_HOST_H
#define IONSOLVER_B200_HOST_H

#include "ionsolver_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* LbmConfig (src/lbm/mod.rs:46-101) + Units (src/lbm/units.rs:14-27) as one POD block */
typedef struct IonLbmConfig {
    uint32_t velocity_set;    /* IonVelocitySet */
    uint32_t relaxation_time; /* IonRelaxationTime */
    uint32_t float_type;      /* IonFloatType */
    float unit_m, unit_kg, unit_s, unit_a, unit_k;
    uint32_t propellant;      /* 0 H, 1 He, 2 Ne, 3 Ar, 4 Kr, 5 Xe (units.rs:211-219) */
    uint32_t n_x, n_y, n_z;
    uint32_t d_x, d_y, d_z;
    float nu;
    float f_x, f_y, f_z;
    uint8_t ext_equilibrium_boudaries, ext_volume_force, ext_force_field, ext_magneto_hydro, ext_subgrid_ecr;
    uint8_t mhd_lod_depth;
    uint8_t graphics_active;  /* graphics_config.graphics_active: the only graphics switch that reaches the kernels */
    uint8_t deterministic;    /* B200 extension, not in the reference: ION_EXT_DETERMINISTIC (reproducible, reference-ordered E/B) */
    float ecr_freq, ecr_field_strength;
    uint64_t run_steps;
} IonLbmConfig;

typedef struct ion_lbm ion_lbm_t;

/* LbmConfig::new (mod.rs:103-135) and Units (units.rs) ------------------------------------------------------ */
ION_API void ion_lbm_config_default(IonLbmConfig* cfg);
ION_API void ion_units_set(IonLbmConfig* cfg, float lbm_length, float lbm_velocity, float lbm_rho, float lbm_charge, float lbm_temp,
                           float si_length, float si_velocity, float si_rho, float si_charge, float si_temp); /* units.rs:41-59 */
enum IonUnitFn { /* Units::<name>(v) */
    ION_UNIT_LEN_SI_LU = 0, ION_UNIT_NU_SI_LU = 1, ION_UNIT_CHARGE_SI_LU = 2, ION_UNIT_MAG_FLUX_SI_LU = 3, ION_UNIT_E_FIELD_SI_LU = 4,
    ION_UNIT_MAGNETIZATION_SI_LU = 5, ION_UNIT_TIME_LU_SI = 6, ION_UNIT_TIME_SI_LU = 7, ION_UNIT_SPEED_SI_LU = 8, ION_UNIT_LEN_LU_SI = 9,
    ION_UNIT_SPEED_LU_SI = 10, ION_UNIT_CHARGE_LU_SI = 11, ION_UNIT_MAG_FLUX_LU_SI = 12, ION_UNIT_E_FIELD_LU_SI = 13,
    /* constants (argument ignored), units.rs:149-197 */
    ION_UNIT_EPSILON_0_LU = 32, ION_UNIT_KE_LU = 33, ION_UNIT_MU_0_LU = 34, ION_UNIT_KKGE_LU = 35, ION_UNIT_KIMG_LU = 36, ION_UNIT_KVEV_LU = 37,
    ION_UNIT_KKBME_LU = 38, ION_UNIT_KEABS_LU = 39, ION_UNIT_KME_LU = 40
};
ION_API float ion_units_eval(const IonLbmConfig* cfg, int fn, float v);
/* get_device_defines (domain.rs:736-858) for domain d of the configuration, without touching a GPU */
ION_API int ion_lbm_make_params(const IonLbmConfig* cfg, uint32_t d, IonParams* out);

/* Lbm (mod.rs:152-495) -------------------------------------------------------------------------------------- */
/* Lbm::new (mod.rs:166); devices: CUDA ordinal per domain (n_devices entries, reused round-robin), NULL = automatic */
ION_API int ion_lbm_create(const IonLbmConfig* cfg, const int* devices, int n_devices, ion_lbm_t** out);
/* one process per GPU: this process builds domain `rank` of d_x*d_y*d_z == world; comm_id from ion_comm_unique_id */
ION_API int ion_lbm_create_distributed(const IonLbmConfig* cfg, int rank, int world, int device,
                                       const uint8_t comm_id[ION_COMM_ID_BYTES], ion_lbm_t** out);
ION_API int ion_lbm_destroy(ion_lbm_t* lbm);
ION_API int ion_lbm_get_config(const ion_lbm_t* lbm, IonLbmConfig* out);      /* Lbm.config (after resolution rounding, mod.rs:167-179) */
ION_API int ion_lbm_local_domains(const ion_lbm_t* lbm, uint32_t* count);     /* Lbm.domains.len() in this process */
ION_API int ion_lbm_domain(ion_lbm_t* lbm, uint32_t local_index, ion_domain_t** out, uint32_t* domain_index); /* &lbm.domains[i] (borrowed) */
ION_API int ion_lbm_initialize(ion_lbm_t* lbm);                               /* mod.rs:214 */
ION_API int ion_lbm_run(ion_lbm_t* lbm, uint64_t steps);                      /* mod.rs:235 */
ION_API int ion_lbm_do_time_step(ion_lbm_t* lbm);                             /* mod.rs:250 */
ION_API int ion_lbm_finish_queues(ion_lbm_t* lbm);                            /* mod.rs:275 */
ION_API int ion_lbm_get_time_step(const ion_lbm_t* lbm, uint64_t* t);         /* mod.rs:492 */
ION_API int ion_lbm_precompute_b(ion_lbm_t* lbm);                             /* mod.rs:284 */
ION_API int ion_lbm_precompute_e(ion_lbm_t* lbm);                             /* mod.rs:301 */
ION_API int ion_lbm_precompute_e_ecr(ion_lbm_t* lbm);                         /* mod.rs:319 */
/* communicate_fi / _rho_u_flags / _ei / _fqi (mod.rs:410-433; field = IonTransferField) and communicate_qu_lods (mod.rs:436) */
ION_API int ion_lbm_communicate_field(ion_lbm_t* lbm, int field);
ION_API int ion_lbm_communicate_qu_lods(ion_lbm_t* lbm);
ION_API int ion_lbm_set_time_step(ion_lbm_t* lbm, uint64_t t);                /* test hook: LbmDomain.t is a pub field */

/* meshes (mesh.rs:233-279) */
ION_API int ion_lbm_import_mesh(ion_lbm_t* lbm, const char* path, float scale, float ox, float oy, float oz, float rx, float ry, float rz);
ION_API int ion_lbm_import_mesh_reposition(ion_lbm_t* lbm, const char* path, float cx, float cy, float cz, float rx, float ry, float rz, float size);
enum IonModelType { ION_MODEL_SOLID = 0, ION_MODEL_MAGNET = 1, ION_MODEL_CHARGED = 2, ION_MODEL_CHARGED_ECR = 3 }; /* mesh.rs:9-14 */
ION_API int ion_lbm_voxelise_mesh(ion_lbm_t* lbm, uint32_t index, int model_type, float v0, float v1, float v2);
ION_API int ion_lbm_mesh_info(const ion_lbm_t* lbm, uint32_t index, uint32_t* triangle_number, float p_min[3], float p_max[3]);
ION_API int ion_lbm_mesh_triangles(const ion_lbm_t* lbm, uint32_t index, float* p0, float* p1, float* p2); /* 3*triangles floats each */
ION_API int ion_lbm_mesh_translate(ion_lbm_t* lbm, uint32_t index, float tx, float ty, float tz);           /* Mesh::translate, mesh.rs:120 */

/* scene helpers (setup.rs) */
ION_API int ion_lbm_set_taylor_green(ion_lbm_t* lbm, uint32_t periodicity);                            /* setup.rs:458 */
ION_API int ion_lbm_setup_velocity_field(ion_lbm_t* lbm, float vx, float vy, float vz, float density); /* setup.rs:547 */
ION_API int ion_setup_taylor_green(uint32_t n, uint32_t d_z, int velocity_set, int float_type, int graphics_active, const int* devices,
                                   int n_devices, ion_lbm_t** out);                                    /* setup.rs:92 / :115 */
ION_API int ion_setup_lid_driven_cavity(uint32_t n, const int* devices, int n_devices, ion_lbm_t** out);
/* the reference's scene functions by name: "setup_verification" (setup.rs:203), "setup_field_vis" (:244), "setup_ecr_test" (:280),
 * "setup_mesh_test" (:320), "setup_mesh_field_test" (:346), "setup_deeva_test" (:395), "setup_taylor_green" (:92), "setup_domain_test"
 * (:115).  stl_dir replaces the hard-coded "stl/" prefix; scale multiplies the lengths of the two mesh scenes (1 = 128 x 256 x 128);
 * flags for setup_deeva_test: bit 0 = ext_subgrid_ecr as in the reference (off: static E from the plates instead of E_var),
 * bit 1 = run initialize + the first time step like setup.rs:447-449 */
ION_API int ion_setup_scene(const char* name, const char* stl_dir, float scale, uint32_t flags, const int* devices, int n_devices, ion_lbm_t** out);
ION_API int ion_setup_charged_fluid(uint32_t nx, uint32_t ny, uint32_t nz, int velocity_set, int float_type, uint32_t lod_depth,
                                    const char* magnet_stl, const int* devices, int n_devices, ion_lbm_t** out); /* setup.rs:142 + :346 */

/* save / load (file.rs, FILE_LAYOUT.txt); buffers returned through `data` are released with ion_free */
ION_API int ion_lbm_encode(ion_lbm_t* lbm, int reference_compatible, uint8_t** data, size_t* len);     /* file.rs:191 */
ION_API int ion_lbm_decode(const uint8_t* data, size_t len, IonLbmConfig* cfg_inout, int reference_compatible, const int* devices,
                           int n_devices, ion_lbm_t** out);                                            /* file.rs:42 */
ION_API int ion_lbm_write_file(ion_lbm_t* lbm, const char* path);                                      /* file.rs:23 */
ION_API int ion_lbm_read_file(const char* path, IonLbmConfig* cfg_inout, ion_lbm_t** out);             /* file.rs:12 */
ION_API int ion_config_to_json(const IonLbmConfig* cfg, char** json);                                  /* file.rs:323 */
ION_API int ion_config_from_json(const char* json, IonLbmConfig* cfg);                                 /* file.rs:310 */
ION_API int ion_lbm_dump_cell(ion_lbm_t* lbm, uint32_t local_index, uint64_t cell, char** text);      /* domain.rs:584 */
/* Slice read-back and PNG (graphics.rs:124-130 slice_mode / slice_x,y,z, :328-373 PNG frames).  slice_mode 1,2,3 = X,Y,Z
 * (SliceMode), index = global cell coordinate on that axis, field / component as ion_read_slice.  The plane is assembled over the
 * local domains without halo layers: X -> out[gy + gz*Ny], Y -> out[gx + gz*Nx], Z -> out[gx + gy*Nx]; capacity in floats. */
ION_API int ion_lbm_read_slice(ion_lbm_t* lbm, int field, int component, uint32_t slice_mode, uint32_t index, float* out, size_t capacity,
                               uint32_t* width, uint32_t* height);
/* one pixel per cell, colour = iron_colormap((value - v_min) / (v_max - v_min)) (graphics_kernels.cl:412-428), second axis up */
ION_API int ion_lbm_write_slice_png(ion_lbm_t* lbm, int field, int component, uint32_t slice_mode, uint32_t index, float v_min, float v_max,
                                    const char* path);
ION_API uint32_t ion_iron_colormap(float x);                                              /* 0xRRGGBB; no GPU needed */
ION_API int ion_write_png_rgb(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height); /* no GPU needed */
ION_API void ion_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* IONSOLVER_B200_HOST_H */
