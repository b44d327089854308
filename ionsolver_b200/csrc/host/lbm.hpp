// lbm.hpp -- host side of IonSolver-B200: the reference's Rust host surface (LbmConfig / Lbm / LbmDomain / Units /
// Mesh, the setup.rs scene helpers and the .ion file format), restated in C++ because no Rust toolchain exists in
// this environment.  Everything below talks to the GPU exclusively through the C ABI of include/ionsolver_b200.h --
// the same calls the Rust host makes through `extern "C"` once src/opencl.rs and the ocl crates are removed
// (INTEGRATION.md shows that binding).  Names and argument meaning follow the reference:
//   Units          /root/reference/src/lbm/units.rs
//   VelocitySet..  /root/reference/src/lbm/types.rs
//   LbmConfig, Lbm /root/reference/src/lbm/mod.rs
//   LbmDomain      /root/reference/src/lbm/domain.rs
//   Mesh           /root/reference/src/mesh.rs
//   scene helpers  /root/reference/src/setup.rs
//   .ion / JSON    /root/reference/src/file.rs, FILE_LAYOUT.txt
#pragma once
#include <cstdint>
#include <stdexcept>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/ionsolver_b200.h"

namespace ionhost {

struct IonException : std::runtime_error {  // the reference panics (unwrap/expect); here the C-ABI error text travels up
    int code;
    IonException(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
void check(int code);  // throws IonException with ion_last_error_string() when code != 0

// ---- units.rs ------------------------------------------------------------------------------------------------
enum class Propellant { H, He, Ne, Ar, Kr, Xe };  // units.rs:211-219
struct Units {                                     // units.rs:14-27
    float m = 1.0f, kg = 1.0f, s = 1.0f, a = 1.0f, k = 1.0f;
    Propellant prop = Propellant::H;
    void set(float lbm_length, float lbm_velocity, float lbm_rho, float lbm_charge, float lbm_temp, float si_length,
             float si_velocity, float si_rho, float si_charge, float si_temp);  // units.rs:41-59
    float len_lu_si(float l) const;
    float time_lu_si(float t) const;
    float speed_lu_si(float v) const;
    float charge_lu_si(float q) const;
    float mag_flux_lu_si(float b) const;
    float e_field_lu_si(float e) const;
    float len_si_lu(float l) const;
    float time_si_lu(float t) const;
    float speed_si_lu(float v) const;
    float nu_si_lu(float nu) const;
    float charge_si_lu(float q) const;
    float mag_flux_si_lu(float b) const;
    float e_field_si_lu(float e) const;
    float magnetization_si_lu(float mg) const;
    float epsilon_0_lu() const;
    float ke_lu() const;
    float mu_0_lu() const;
    float k_charge_expansion_lu() const;
    float kkge_lu() const;
    float kimg_lu() const;
    float kveV_lu() const;
    float kkBme_lu() const;
    float keabs_lu() const;
    float kme_lu() const;
    double atom_mass() const;
};

// ---- types.rs ------------------------------------------------------------------------------------------------
enum class VelocitySet : uint8_t { D2Q9 = 0, D3Q15 = 1, D3Q19 = 2, D3Q27 = 3 };
enum class RelaxationTime : uint8_t { Srt = 0, Trt = 1 };
enum class FloatType : uint8_t { FP16S = 0, FP16C = 1, FP32 = 2 };
enum class TransferField : int { Fi = 0, RhoUFlags = 1, Ei = 2, Qi = 3 };
size_t get_transfers(VelocitySet v);                                    // types.rs:28-36
void get_set_values(VelocitySet v, uint8_t& dimensions, uint8_t& velocity_set, uint8_t& transfers);  // types.rs:37-46
size_t size_of(FloatType f);                                            // types.rs:85-92

// the one GraphicsConfig member that reaches the hot path (graphics.rs:160 default true -> UPDATE_FIELDS, domain.rs:856)
struct GraphicsConfig {
    bool graphics_active = true;
    // the remaining serde fields of graphics.rs:101-160 (camera, colour maps, keyframes ...) belong to the rasteriser,
    // which is out of scope; their JSON text is carried through read_config -> write_config untouched
    std::string passthrough_json;
};

// ---- mod.rs:46-135 -------------------------------------------------------------------------------------------
struct LbmConfig {
    VelocitySet velocity_set = VelocitySet::D2Q9;
    RelaxationTime relaxation_time = RelaxationTime::Srt;
    FloatType float_type = FloatType::FP16S;
    Units units;
    uint32_t n_x = 1, n_y = 1, n_z = 1;
    uint32_t d_x = 1, d_y = 1, d_z = 1;
    float nu = 1.0f / 6.0f;
    float f_x = 0.0f, f_y = 0.0f, f_z = 0.0f;
    bool ext_equilibrium_boudaries = false;  // (sic) spelled as in the reference
    bool ext_volume_force = false;
    bool ext_force_field = false;
    bool ext_magneto_hydro = false;
    bool ext_subgrid_ecr = false;
    uint8_t mhd_lod_depth = 4;
    float ecr_freq = 0.0f;
    float ecr_field_strength = 0.0f;
    GraphicsConfig graphics_config;
    uint64_t run_steps = 0;
    // not in the reference: reproducible, reference-ordered E/B path (ION_EXT_DETERMINISTIC in ionsolver_b200.h)
    bool deterministic = false;
};

// ---- mesh.rs ---------------------------------------------------------------------------------------------------
struct F32_3 { float x = 0, y = 0, z = 0; };
struct F32_3_3 { float xx, yx, zx, xy, yy, zy, xz, yz, zz; };
F32_3_3 construct_rotation_matrix(float rx, float ry, float rz);  // mesh.rs:52-66 (radians)
enum class ModelKind { Solid, Magnet, Charged, ChargedECR };
struct ModelType {                                                   // mesh.rs:9-14
    ModelKind kind = ModelKind::Solid;
    float magnetization[3] = {0, 0, 0};
    float charge = 0;
    static ModelType solid() { return ModelType(); }
    static ModelType magnet(float mx, float my, float mz) { ModelType t; t.kind = ModelKind::Magnet; t.magnetization[0] = mx; t.magnetization[1] = my; t.magnetization[2] = mz; return t; }
    static ModelType charged(float c) { ModelType t; t.kind = ModelKind::Charged; t.charge = c; return t; }
    static ModelType charged_ecr(float c) { ModelType t; t.kind = ModelKind::ChargedECR; t.charge = c; return t; }
};
struct Mesh {                                                        // mesh.rs:69-77
    uint32_t triangle_number = 0;
    F32_3 center, p_min, p_max;
    std::vector<F32_3> p0, p1, p2;
    void update_bounds();                                            // mesh.rs:93-108
    void scale(float scale);
    void translate(F32_3 t);
    void rotate(const F32_3_3& r);
    static Mesh read_stl_raw(const std::vector<uint8_t>& file, bool reposition, F32_3 box_size, F32_3 center,
                             const F32_3_3& rotation, float size);  // mesh.rs:175-220
};

// ---- domain.rs ---------------------------------------------------------------------------------------------------
struct Lbm;
struct LbmDomain {
    LbmConfig cfg;
    ion_domain_t* dev = nullptr;  // replaces queue + 11 kernels + ~25 buffers (domain.rs:20-80)
    int device = 0;
    uint32_t n_x = 0, n_y = 0, n_z = 0;
    uint64_t n = 0;
    int32_t o_x = 0, o_y = 0, o_z = 0;
    uint32_t d_i = 0;
    size_t n_lod = 0, n_lod_own = 0;
    float fx = 0, fy = 0, fz = 0;
    uint64_t t = 0;
    IonParams params{};

    LbmDomain() = default;
    LbmDomain(const LbmDomain&) = delete;
    LbmDomain& operator=(const LbmDomain&) = delete;
    LbmDomain(LbmDomain&& o) noexcept;
    ~LbmDomain();
    // LbmDomain::new, domain.rs:88-409
    static LbmDomain create(const LbmConfig& cfg, int device, uint32_t x, uint32_t y, uint32_t z, uint32_t i);
    // get_device_defines (domain.rs:736-858) as a parameter block; usable without a GPU
    static IonParams make_params(const LbmConfig& cfg, uint32_t x, uint32_t y, uint32_t z, uint32_t i);

    void enqueue_initialize();                 // domain.rs:412
    void enqueue_stream_collide();             // domain.rs:419
    void enqueue_stream_collide_range(uint32_t z_begin, uint32_t z_end, bool finish);  // the same kernel on a range of z layers
    void enqueue_update_fields();              // domain.rs:432
    void enqueue_update_e_b_dyn();             // domain.rs:443
    void enqueue_lod_part_2_gather();          // domain.rs:453
    void enqueue_clear_qu_lod();               // domain.rs:464
    size_t get_area(uint32_t direction) const; // domain.rs:475
    void enqueue_transfer_extract_field(TransferField field, uint32_t direction, size_t bytes_per_cell);  // domain.rs:484
    void enqueue_transfer_insert_field(TransferField field, uint32_t direction, size_t bytes_per_cell);   // domain.rs:516
    void enqueue_precompute_b();               // domain.rs:551
    void enqueue_precompute_e();               // domain.rs:558
    void enqueue_precompute_e_ecr();           // domain.rs:569
    void voxelize_mesh_on_device(const Mesh& mesh, const ModelType& ctype);  // mesh.rs:281
    void finish();                             // queue.finish()
    // bwrite!/bread! (ocl-macros): whole-buffer or ranged copies, sizes in BYTES
    void write(int field, const void* host, size_t bytes, size_t offset = 0);
    void read(int field, void* host, size_t bytes, size_t offset = 0) const;
    size_t buffer_bytes(int field) const;
    std::string dump_cell(size_t c) const;     // domain.rs:584-721 (returns the text instead of printing)
};

// ---- mod.rs:152-495 ------------------------------------------------------------------------------------------------
struct Lbm {
    std::vector<LbmDomain> domains;  // all domains (single process) or the one this rank owns (one process per GPU)
    LbmConfig config;
    std::vector<Mesh> meshes;
    bool initialized = false;
    // run the fi/fqi/ei halo exchange concurrently with update_e_b_dynamic (MHD, more than one domain); ION_NO_OVERLAP=1
    // restores the reference's strictly sequential order for A/B timing
    bool overlap_halo = !(getenv("ION_NO_OVERLAP") && atoi(getenv("ION_NO_OVERLAP")) != 0);
    // one process per GPU: rank owns domain `rank`; comm carries the halo exchange (NCCL over NVLink)
    ion_comm_t* comm = nullptr;
    int rank = 0, world = 1;

    Lbm() = default;
    Lbm(const Lbm&) = delete;
    Lbm& operator=(const Lbm&) = delete;
    ~Lbm();
    // Lbm::new, mod.rs:166-210.  `devices`: CUDA device per domain; empty = domain d on device d % device_count
    // (the reference scores OpenCL devices in opencl.rs:9-63 and falls back to one device for all domains).
    static Lbm* create(LbmConfig cfg, const std::vector<int>& devices = {});
    // same configuration, but this process only builds domain `rank` of `world` == d_x*d_y*d_z on `device`
    static Lbm* create_distributed(LbmConfig cfg, int rank, int world, int device, const uint8_t comm_id[ION_COMM_ID_BYTES]);

    void initialize();            // mod.rs:214
    void run(uint64_t steps);     // mod.rs:235
    void do_time_step();          // mod.rs:250
    void do_time_step_body();
    bool boundary_first() const;  // z slabs: boundary layers first, halo exchange overlapped with the interior update
    void finish_queues();         // mod.rs:275
    void precompute_B();          // mod.rs:284
    void precompute_E();          // mod.rs:301
    void precompute_E_ECR();      // mod.rs:319
    size_t get_d_n() const { return (size_t)config.d_x * config.d_y * config.d_z; }
    uint64_t get_time_step() const { return domains[0].t; }
    // mesh.rs:233-279
    void import_mesh(const std::string& path, float scale, float ox, float oy, float oz, float rx, float ry, float rz);
    void import_mesh_reposition(const std::string& path, float cx, float cy, float cz, float rx, float ry, float rz, float size);
    void voxelise_mesh(size_t index, const ModelType& ctype);
    // setup.rs:458-598
    void set_taylor_green(uint32_t periodicity);
    void setup_velocity_field(float vx, float vy, float vz, float density);

    // internals, named as in mod.rs:336-495
    void kernel_initialize();
    void stream_collide();
    void update_e_b_dynamic();
    void build_lods_part_2();
    void clear_qu_lod();
    void communicate_field(TransferField field, size_t bytes_per_cell);
    void communicate_fi();
    void communicate_rho_u_flags();
    void communicate_fqi();
    void communicate_ei();
    void communicate_qu_lods();
    void increment_timestep(uint32_t steps);
    void reset_timestep();
    LbmDomain* local_domain(uint32_t d);  // nullptr when domain d lives in another process
};

// ---- setup.rs scene constructors (only those the BASELINE configs derive from) ------------------------------------
Lbm* setup_taylor_green(uint32_t n, uint32_t d_z, VelocitySet vs, FloatType ft, bool graphics_active, const std::vector<int>& devices);  // setup.rs:92-113 / :115-139
Lbm* setup_lid_driven_cavity(uint32_t n, const std::vector<int>& devices);
Lbm* setup_charged_fluid(uint32_t nx, uint32_t ny, uint32_t nz, VelocitySet vs, FloatType ft, uint8_t lod_depth,
                         const std::string& magnet_stl, const std::vector<int>& devices);        // cfg2: setup.rs:142-201 + :346-393

// the reference's other scene functions (setup.rs:203-453), see setup.cpp
Lbm* setup_verification(const std::vector<int>& devices);                                         // setup.rs:203-241
Lbm* setup_field_vis(const std::vector<int>& devices);                                            // setup.rs:244-277
Lbm* setup_ecr_test(const std::vector<int>& devices);                                             // setup.rs:280-317
Lbm* setup_mesh_test(const std::string& stl_dir, const std::vector<int>& devices);                // setup.rs:320-343
Lbm* setup_mesh_field_test(const std::string& stl_dir, float scale, const std::vector<int>& devices);  // setup.rs:346-393
Lbm* setup_deeva_test(const std::string& stl_dir, float scale, bool subgrid_ecr, bool first_step, const std::vector<int>& devices);  // setup.rs:395-453
Lbm* setup_scene(const std::string& name, const std::string& stl_dir, float scale, uint32_t flags, const std::vector<int>& devices);  // setup.rs:22-64

// ---- file.rs ---------------------------------------------------------------------------------------------------
namespace file {
std::vector<uint8_t> encode(Lbm& lbm, bool reference_compatible);              // file.rs:191-306
Lbm* decode(const std::vector<uint8_t>& buffer, LbmConfig& config, bool reference_compatible, const std::vector<int>& devices);  // file.rs:42-188
void write(Lbm& lbm, const std::string& path, bool reference_compatible = false);  // file.rs:23-31
Lbm* read(const std::string& path, LbmConfig& config, bool reference_compatible = false);  // file.rs:12-20
std::string config_to_json(const LbmConfig& cfg);                              // serde_json::to_vec, file.rs:323-334
LbmConfig config_from_json(const std::string& text);                           // serde_json::from_slice, file.rs:310-321
void write_config(const std::string& path, const LbmConfig& cfg);
LbmConfig read_config(const std::string& path);
std::vector<uint8_t> read_file(const std::string& path);
}  // namespace file

// ---- slice read-back + PNG (slice.cpp; graphics.rs:124-130 slice selection, :328-373 PNG frames) --------------------
namespace slice {
void read(Lbm& lbm, int field, int component, uint32_t slice_mode, uint32_t index, std::vector<float>& out, uint32_t& w, uint32_t& h);
void write_png(Lbm& lbm, int field, int component, uint32_t slice_mode, uint32_t index, float v_min, float v_max, const std::string& path);
uint32_t iron_colormap(float x);                                                   // graphics_kernels.cl:412-428 -> 0xRRGGBB
std::vector<uint8_t> encode_png_rgb(const uint8_t* rgb, uint32_t w, uint32_t h);   // 8-bit RGB, stored deflate blocks
}  // namespace slice

}  // namespace ionhost
