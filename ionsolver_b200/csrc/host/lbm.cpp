// lbm.cpp -- Units, LbmDomain and Lbm of the host layer (see lbm.hpp for the reference files each part follows).
// Built with -ffp-contract=off: the f32 arithmetic of units.rs reaches the kernels as constants and has to
// produce the same bits as the Rust host.
#include "lbm.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace ionhost {

void check(int code) {
    if (code != 0) throw IonException(code, std::string("ionsolver_b200 error ") + std::to_string(code) + ": " + ion_last_error_string());
}

// =================================================================================================================
// units.rs
// =================================================================================================================
static inline float sq(float x) { return x * x; }
static inline double sqd(double x) { return x * x; }
static inline float cb(float x) { return x * x * x; }
static inline float to4(float x) { return x * x * x * x; }
static const float PI_F = 3.14159274101257324219f;  // std::f32::consts::PI

void Units::set(float lbm_length, float lbm_velocity, float lbm_rho, float lbm_charge, float lbm_temp, float si_length,
                float si_velocity, float si_rho, float si_charge, float si_temp) {  // units.rs:41-59
    m = si_length / lbm_length;
    kg = si_rho / lbm_rho * cb(m);
    s = m / (si_velocity / lbm_velocity);
    a = si_charge / lbm_charge / s;
    k = si_temp / lbm_temp;
}
float Units::len_lu_si(float l) const { return l * m; }
float Units::time_lu_si(float t) const { return t * s; }
float Units::speed_lu_si(float v) const { return v * (m / s); }
float Units::charge_lu_si(float q) const { return q * (a * s); }
float Units::mag_flux_lu_si(float b) const { return b * (kg / (a * sq(s))); }
float Units::e_field_lu_si(float e) const { return e * ((kg * m) / (a * cb(s))); }
float Units::len_si_lu(float l) const { return l / m; }
float Units::time_si_lu(float t) const { return t / s; }
float Units::speed_si_lu(float v) const { return v / (m / s); }
float Units::nu_si_lu(float nu) const { return nu / (sq(m) / s); }
float Units::charge_si_lu(float q) const { return q / (a * s); }
float Units::mag_flux_si_lu(float b) const { return b / (kg / (a * sq(s))); }
float Units::e_field_si_lu(float e) const { return e / ((kg * m) / (a * cb(s))); }
float Units::magnetization_si_lu(float mg) const { return mg / (a / m); }
float Units::epsilon_0_lu() const { return 8.8541878128E-12f / ((sq(a) * to4(s)) / (kg * cb(m))); }  // units.rs:149-153
float Units::ke_lu() const { return 1.0f / (4.0f * PI_F * epsilon_0_lu()); }                         // units.rs:155-161
float Units::mu_0_lu() const { return 1.256637062E-6f / ((kg * m) / (sq(a) * sq(s))); }             // units.rs:163-167
float Units::k_charge_expansion_lu() const { return 1.0f; }                                          // units.rs:169-173
float Units::kkge_lu() const { return (float)(9.1093837139E-31 / -1.602176634E-19) / (kg / (a * s)); }  // units.rs:175-177
double Units::atom_mass() const {                                                                    // units.rs:236-245
    switch (prop) {
        case Propellant::H: return 1.6735575e-27;
        case Propellant::He: return 6.6464731e-27;
        case Propellant::Ne: return 3.3509177e-26;
        case Propellant::Ar: return 6.6335209e-26;
        case Propellant::Kr: return 1.3914984e-25;
        default: return 2.1801714e-25;
    }
}
float Units::kimg_lu() const { return (float)((1.0 / (atom_mass() * 1e20)) / (double)kg); }
float Units::kveV_lu() const { return (float)(9.1093837139E-31 / (2.0 * 1.602176634E-19) / (double)kg); }
float Units::kkBme_lu() const { return (float)(-22734499.72063751808909449412 / (sqd((double)m) / (sqd((double)s) * (double)k))); }
float Units::keabs_lu() const { return (float)(1.40897016100511360652E-8 / (sqd((double)a) * sqd((double)s) / (double)kg)); }
float Units::kme_lu() const { return (float)(5.68563006E-12 / ((double)kg / ((double)a * (double)s))); }

// =================================================================================================================
// types.rs
// =================================================================================================================
size_t get_transfers(VelocitySet v) {
    switch (v) { case VelocitySet::D2Q9: return 3; case VelocitySet::D3Q15: return 5; case VelocitySet::D3Q19: return 5; default: return 9; }
}
void get_set_values(VelocitySet v, uint8_t& dimensions, uint8_t& velocity_set, uint8_t& transfers) {
    switch (v) {
        case VelocitySet::D2Q9: dimensions = 2; velocity_set = 9; transfers = 3; break;
        case VelocitySet::D3Q15: dimensions = 3; velocity_set = 15; transfers = 5; break;
        case VelocitySet::D3Q19: dimensions = 3; velocity_set = 19; transfers = 5; break;
        default: dimensions = 3; velocity_set = 27; transfers = 9; break;
    }
}
size_t size_of(FloatType f) { return f == FloatType::FP32 ? 4 : 2; }

static void get_coordinates_sl(uint64_t n, uint32_t n_x, uint32_t n_y, uint32_t& x, uint32_t& y, uint32_t& z) {  // mod.rs:31-40
    const uint64_t t = n % ((uint64_t)n_x * n_y);
    x = (uint32_t)(t % n_x);
    y = (uint32_t)(t / n_x);
    z = (uint32_t)(n / ((uint64_t)n_x * n_y));
}
static size_t ipow(size_t b, unsigned e) { size_t r = 1; while (e--) r *= b; return r; }

// =================================================================================================================
// LbmDomain (domain.rs)
// =================================================================================================================
IonParams LbmDomain::make_params(const LbmConfig& c, uint32_t x, uint32_t y, uint32_t z, uint32_t i) {
    IonParams p;
    memset(&p, 0, sizeof(p));
    p.abi_version = ION_ABI_VERSION;
    p.nx = c.n_x / c.d_x + 2u * (c.d_x > 1u);  // domain.rs:91-93
    p.ny = c.n_y / c.d_y + 2u * (c.d_y > 1u);
    p.nz = c.n_z / c.d_z + 2u * (c.d_z > 1u);
    p.dx = c.d_x; p.dy = c.d_y; p.dz = c.d_z; p.di = i;
    p.ox = (int32_t)(x * c.n_x / c.d_x) - (int32_t)(c.d_x > 1u);  // domain.rs:101-103
    p.oy = (int32_t)(y * c.n_y / c.d_y) - (int32_t)(c.d_y > 1u);
    p.oz = (int32_t)(z * c.n_z / c.d_z) - (int32_t)(c.d_z > 1u);
    uint8_t dimensions, velocity_set, transfers;
    get_set_values(c.velocity_set, dimensions, velocity_set, transfers);
    // LOD counts, domain.rs:110-126
    size_t cnt = 1;
    for (unsigned k = 0; k < c.mhd_lod_depth; k++) cnt += ipow((size_t)1 << (k + 1), dimensions);
    const size_t own = cnt;
    const uint32_t d_n = c.d_x * c.d_y * c.d_z;
    for (uint32_t d = 0; d < d_n; d++) {
        uint32_t dx, dy, dz;
        get_coordinates_sl(d, c.d_x, c.d_y, dx, dy, dz);
        const int dist = std::max(std::abs((int)z - (int)dz), std::max(std::abs((int)y - (int)dy), std::abs((int)x - (int)dx)));
        if (dist != 0) cnt += ipow((size_t)1 << std::max((int)c.mhd_lod_depth - dist, 0), dimensions);
    }
    p.velocity_set = (uint32_t)c.velocity_set;
    p.relaxation_time = (uint32_t)c.relaxation_time;
    p.float_type = (uint32_t)c.float_type;
    p.ext = (c.ext_equilibrium_boudaries ? ION_EXT_EQUILIBRIUM_BOUNDARIES : 0u) | (c.ext_volume_force ? ION_EXT_VOLUME_FORCE : 0u) |
            (c.ext_force_field ? ION_EXT_FORCE_FIELD : 0u) | (c.ext_magneto_hydro ? ION_EXT_MAGNETO_HYDRO : 0u) |
            (c.ext_subgrid_ecr ? ION_EXT_SUBGRID_ECR : 0u) | (c.graphics_config.graphics_active ? ION_EXT_UPDATE_FIELDS : 0u) |
            (c.deterministic && c.ext_magneto_hydro ? ION_EXT_DETERMINISTIC : 0u);
    p.w = 1.0f / (3.0f * c.nu + 0.5f);  // domain.rs:808
    const Units& u = c.units;            // domain.rs:838-853
    p.ke = u.ke_lu();
    p.kmu = u.mu_0_lu() / (4.0f * PI_F);
    p.kmu0 = u.mu_0_lu();
    p.kkge = u.kkge_lu();
    p.kimg = u.kimg_lu();
    p.kvev = u.kveV_lu();
    p.kme = u.kme_lu();
    p.wq = 1.0f / (2.0f * u.k_charge_expansion_lu() + 0.5f);
    p.kkbme = u.kkBme_lu();
    p.keabs = u.keabs_lu();
    p.lod_depth = c.mhd_lod_depth;
    p.n_lod = (uint32_t)cnt;
    p.n_lod_own = (uint32_t)own;
    return p;
}

LbmDomain LbmDomain::create(const LbmConfig& cfg, int device, uint32_t x, uint32_t y, uint32_t z, uint32_t i) {
    LbmDomain d;
    d.cfg = cfg;
    d.device = device;
    d.params = make_params(cfg, x, y, z, i);
    d.n_x = d.params.nx; d.n_y = d.params.ny; d.n_z = d.params.nz;
    d.n = (uint64_t)d.n_x * d.n_y * d.n_z;
    d.o_x = d.params.ox; d.o_y = d.params.oy; d.o_z = d.params.oz;
    d.d_i = i;
    d.n_lod = d.params.n_lod;
    d.n_lod_own = d.params.n_lod_own;
    d.fx = cfg.f_x; d.fy = cfg.f_y; d.fz = cfg.f_z;
    d.t = 0;
    check(ion_domain_create(&d.params, device, &d.dev));
    if (cfg.ext_subgrid_ecr) check(ion_domain_set_ecr_freq(d.dev, cfg.ecr_freq));
    return d;
}
LbmDomain::LbmDomain(LbmDomain&& o) noexcept { memcpy((void*)&params, &o.params, sizeof(params));
    cfg = o.cfg; dev = o.dev; o.dev = nullptr; device = o.device; n_x = o.n_x; n_y = o.n_y; n_z = o.n_z; n = o.n; o_x = o.o_x; o_y = o.o_y; o_z = o.o_z;
    d_i = o.d_i; n_lod = o.n_lod; n_lod_own = o.n_lod_own; fx = o.fx; fy = o.fy; fz = o.fz; t = o.t; }
LbmDomain::~LbmDomain() {
    if (dev) ion_domain_destroy(dev);
    dev = nullptr;
}

void LbmDomain::enqueue_initialize() { check(ion_enqueue_initialize(dev)); }
void LbmDomain::enqueue_stream_collide() { check(ion_enqueue_stream_collide(dev, t, fx, fy, fz)); }
void LbmDomain::enqueue_stream_collide_range(uint32_t z_begin, uint32_t z_end, bool finish) {
    check(ion_enqueue_stream_collide_range(dev, t, fx, fy, fz, z_begin, z_end, finish ? 1 : 0));
}
void LbmDomain::enqueue_update_fields() { check(ion_enqueue_update_fields(dev, t, fx, fy, fz)); }
void LbmDomain::enqueue_update_e_b_dyn() { check(ion_enqueue_update_e_b_dyn(dev)); }
void LbmDomain::enqueue_lod_part_2_gather() { check(ion_enqueue_lod_part_2_gather(dev)); }
void LbmDomain::enqueue_clear_qu_lod() { check(ion_enqueue_clear_qu_lod(dev)); }
size_t LbmDomain::get_area(uint32_t direction) const {
    const size_t a[3] = {(size_t)n_y * n_z, (size_t)n_x * n_z, (size_t)n_x * n_y};
    return a[direction];
}
void LbmDomain::enqueue_transfer_extract_field(TransferField field, uint32_t direction, size_t) {
    // kernel only: the device->host read of domain.rs:504-511 is gone, faces stay in device memory
    check(ion_enqueue_transfer_extract(dev, (int)field, direction, t));
}
void LbmDomain::enqueue_transfer_insert_field(TransferField field, uint32_t direction, size_t) {
    check(ion_enqueue_transfer_insert(dev, (int)field, direction, t));
}
void LbmDomain::enqueue_precompute_b() { check(ion_enqueue_precompute_b(dev)); }
void LbmDomain::enqueue_precompute_e() { check(ion_enqueue_precompute_e(dev)); }
void LbmDomain::enqueue_precompute_e_ecr() { check(ion_enqueue_precompute_e_ecr(dev)); }
void LbmDomain::finish() { check(ion_finish(dev)); }
void LbmDomain::write(int field, const void* host, size_t bytes, size_t offset) { check(ion_buffer_write(dev, field, host, offset, bytes)); }
void LbmDomain::read(int field, void* host, size_t bytes, size_t offset) const { check(ion_buffer_read(dev, field, host, offset, bytes)); }
size_t LbmDomain::buffer_bytes(int field) const {
    size_t b = 0;
    check(ion_buffer_size(dev, field, &b));
    return b;
}

void LbmDomain::voxelize_mesh_on_device(const Mesh& mesh, const ModelType& ctype) {  // mesh.rs:281-343
    std::vector<float> p0(mesh.triangle_number * 3u), p1(p0.size()), p2(p0.size());
    for (uint32_t i = 0; i < mesh.triangle_number; i++) {
        p0[3 * i] = mesh.p0[i].x; p0[3 * i + 1] = mesh.p0[i].y; p0[3 * i + 2] = mesh.p0[i].z;
        p1[3 * i] = mesh.p1[i].x; p1[3 * i + 1] = mesh.p1[i].y; p1[3 * i + 2] = mesh.p1[i].z;
        p2[3 * i] = mesh.p2[i].x; p2[3 * i + 1] = mesh.p2[i].y; p2[3 * i + 2] = mesh.p2[i].z;
    }
    const float x0 = mesh.p_min.x - 2.0f, y0 = mesh.p_min.y - 2.0f, z0 = mesh.p_min.z - 2.0f;
    const float x1 = mesh.p_max.x + 2.0f, y1 = mesh.p_max.y + 2.0f, z1 = mesh.p_max.z + 2.0f;
    float bbu[7];
    memcpy(&bbu[0], &mesh.triangle_number, 4);  // f32::from_bits(triangle_number)
    bbu[1] = x0; bbu[2] = y0; bbu[3] = z0; bbu[4] = x1; bbu[5] = y1; bbu[6] = z1;
    const float c[3] = {(y1 - y0) * (z1 - z0), (z1 - z0) * (x1 - x0), (x1 - x0) * (y1 - y0)};
    const uint32_t direction = (c[0] < c[1] && c[0] < c[2]) ? 0u : (c[1] < c[2] ? 1u : 2u);
    uint8_t flag = 0x01;
    float mpc[3] = {0.0f, 0.0f, 0.0f};
    switch (ctype.kind) {
        case ModelKind::Solid: flag = 0x01; break;
        case ModelKind::Magnet: flag = 0x11; break;
        case ModelKind::Charged: flag = 0x09; break;
        case ModelKind::ChargedECR: flag = 0x05; break;
    }
    if (cfg.ext_magneto_hydro) {
        if (ctype.kind == ModelKind::Magnet) {
            for (int k = 0; k < 3; k++) mpc[k] = cfg.units.magnetization_si_lu(ctype.magnetization[k]);
        } else if (ctype.kind == ModelKind::Charged || ctype.kind == ModelKind::ChargedECR) {
            mpc[0] = cfg.units.charge_si_lu(ctype.charge);
        }
    }
    check(ion_voxelize_mesh(dev, p0.data(), p1.data(), p2.data(), mesh.triangle_number, bbu, direction, flag, mpc[0], mpc[1], mpc[2], t + 1));
}

std::string LbmDomain::dump_cell(size_t c) const {  // domain.rs:584-721
    const size_t c_x = c, c_y = c + n, c_z = c + 2 * n;
    uint32_t x, y, z;
    get_coordinates_sl(c, n_x, n_y, x, y, z);
    auto rd = [&](int field, size_t idx) {
        float v = 0.0f;
        if (buffer_bytes(field)) read(field, &v, 4, idx * 4);
        return v;
    };
    auto rd3 = [&](int field, float* o) { o[0] = rd(field, c_x); o[1] = rd(field, c_y); o[2] = rd(field, c_z); };
    const float rho = rd(ION_FIELD_RHO, c);
    float u[3], es[3], bs[3], ed[3], bd[3], ev[3];
    rd3(ION_FIELD_U, u); rd3(ION_FIELD_E_STAT, es); rd3(ION_FIELD_B_STAT, bs); rd3(ION_FIELD_E_DYN, ed); rd3(ION_FIELD_B_DYN, bd); rd3(ION_FIELD_E_VAR, ev);
    uint8_t flag = 0;
    read(ION_FIELD_FLAGS, &flag, 1, c);
    const float q = rd(ION_FIELD_Q, c);
    const Units& un = cfg.units;
    char buf[2048];
    snprintf(buf, sizeof(buf),
             "Dumping cell %zu:\n    x: %u, y: %u, z: %u\n    rho:    %.9g / %.9g kg/m3\n    u:      %.9g, %.9g, %.9g / %.9g, %.9g, %.9g m/s\n"
             "    flags:  %u\n    e_stat:      %.9g, %.9g, %.9g / %.9g, %.9g, %.9g V/m\n    b_stat:      %.9g, %.9g, %.9g / %.9g, %.9g, %.9g T\n"
             "    e_dyn:  %.9g, %.9g, %.9g / %.9g, %.9g, %.9g V/m\n    b_dyn:  %.9g, %.9g, %.9g / %.9g, %.9g, %.9g T\n"
             "    e_var:  %.9g, %.9g, %.9g / %.9g, %.9g, %.9g V/m\n    charge: %.9g / %.9g As\n",
             c, x, y, z, rho, rho * (un.kg / cb(un.m)), u[0], u[1], u[2], un.speed_lu_si(u[0]), un.speed_lu_si(u[1]), un.speed_lu_si(u[2]),
             (unsigned)flag, es[0], es[1], es[2], un.e_field_lu_si(es[0]), un.e_field_lu_si(es[1]), un.e_field_lu_si(es[2]), bs[0], bs[1], bs[2],
             un.mag_flux_lu_si(bs[0]), un.mag_flux_lu_si(bs[1]), un.mag_flux_lu_si(bs[2]), ed[0], ed[1], ed[2], un.e_field_lu_si(ed[0]),
             un.e_field_lu_si(ed[1]), un.e_field_lu_si(ed[2]), bd[0], bd[1], bd[2], un.mag_flux_lu_si(bd[0]), un.mag_flux_lu_si(bd[1]),
             un.mag_flux_lu_si(bd[2]), ev[0], ev[1], ev[2], un.e_field_lu_si(ev[0]), un.e_field_lu_si(ev[1]), un.e_field_lu_si(ev[2]), q,
             un.charge_lu_si(q));
    return buf;
}

// =================================================================================================================
// Lbm (mod.rs)
// =================================================================================================================
static void round_resolution(LbmConfig& c) {  // mod.rs:167-179
    c.n_x = (c.n_x / c.d_x) * c.d_x;
    c.n_y = (c.n_y / c.d_y) * c.d_y;
    c.n_z = (c.n_z / c.d_z) * c.d_z;
}

Lbm* Lbm::create(LbmConfig cfg, const std::vector<int>& devices) {
    if (!cfg.d_x || !cfg.d_y || !cfg.d_z) throw IonException(ION_ERR_INVALID, "zero domain count");
    round_resolution(cfg);
    const uint32_t domain_numbers = cfg.d_x * cfg.d_y * cfg.d_z;
    int ndev = 0;
    check(ion_device_count(&ndev));
    if (ndev == 0) throw IonException(ION_ERR_NO_DEVICE, "no sm_100 CUDA device; ionsolver_b200 has no CPU fallback");
    Lbm* lbm = new Lbm();
    try {
        lbm->config = cfg;
        for (uint32_t d = 0; d < domain_numbers; d++) {
            const uint32_t x = (d % (cfg.d_x * cfg.d_y)) % cfg.d_x, y = (d % (cfg.d_x * cfg.d_y)) / cfg.d_x, z = d / (cfg.d_x * cfg.d_y);
            const int dev = devices.empty() ? (int)(d % (uint32_t)ndev) : devices[d % devices.size()];
            lbm->domains.push_back(LbmDomain::create(cfg, dev, x, y, z, d));
        }
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

Lbm* Lbm::create_distributed(LbmConfig cfg, int rank, int world, int device, const uint8_t comm_id[ION_COMM_ID_BYTES]) {
    round_resolution(cfg);
    const uint32_t domain_numbers = cfg.d_x * cfg.d_y * cfg.d_z;
    if ((int)domain_numbers != world) throw IonException(ION_ERR_INVALID, "one process per GPU needs d_x*d_y*d_z == world size");
    Lbm* lbm = new Lbm();
    try {
        lbm->config = cfg;
        lbm->rank = rank;
        lbm->world = world;
        const uint32_t d = (uint32_t)rank;
        const uint32_t x = (d % (cfg.d_x * cfg.d_y)) % cfg.d_x, y = (d % (cfg.d_x * cfg.d_y)) / cfg.d_x, z = d / (cfg.d_x * cfg.d_y);
        lbm->domains.push_back(LbmDomain::create(cfg, device, x, y, z, d));
        if (world > 1) check(ion_comm_create(comm_id, rank, world, device, &lbm->comm));
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

Lbm::~Lbm() {
    // A partner domain's stream (possibly on another GPU) may still hold a peer copy or an insert kernel that reads this
    // domain's transfer buffers: drain every stream, halo streams included (ion_finish joins them), before anything is freed.
    for (auto& d : domains) {
        if (d.dev) ion_finish(d.dev);  // no throw from a destructor
    }
    domains.clear();
    if (comm) ion_comm_destroy(comm);
}

LbmDomain* Lbm::local_domain(uint32_t d) {
    if (world == 1) return d < domains.size() ? &domains[d] : nullptr;
    return (int)d == rank ? &domains[0] : nullptr;
}

void Lbm::initialize() {  // mod.rs:214-231
    increment_timestep(1);  // the communicate calls at initialization need an odd time step
    communicate_rho_u_flags();
    kernel_initialize();
    communicate_rho_u_flags();
    communicate_fi();
    if (config.ext_magneto_hydro) {
        communicate_fqi();
        communicate_ei();
        communicate_qu_lods();
        update_e_b_dynamic();
    }
    finish_queues();
    reset_timestep();
    initialized = true;
}

void Lbm::run(uint64_t steps) {  // mod.rs:235-245
    if (!initialized) initialize();
    for (uint64_t i = 0; i < steps; i++) do_time_step();
}

void Lbm::do_time_step() {  // mod.rs:250-272
    try {
        do_time_step_body();
    } catch (...) {
        // leave no halo stream forked and nothing queued that refers to another domain's buffers before the error propagates
        for (auto& d : domains)
            if (d.dev) ion_finish(d.dev);
        throw;
    }
}

// z-slab decompositions (d_x = d_y = 1) with at least four layers per slab take the boundary-first schedule
bool Lbm::boundary_first() const {
    if (!overlap_halo || config.d_x != 1u || config.d_y != 1u || config.d_z < 2u) return false;
    for (const auto& d : domains)
        if (d.n_z < 4u) return false;
    static const bool off = getenv("ION_NO_BOUNDARY_FIRST") && atoi(getenv("ION_NO_BOUNDARY_FIRST")) != 0;
    return !off;
}

void Lbm::do_time_step_body() {
    const bool mhd = config.ext_magneto_hydro;
    if (mhd) clear_qu_lod();
    if (boundary_first()) {
        // Same operations as mod.rs:250-272, scheduled for overlap.  (1) stream_collide on the two layers next to the halos: they
        // write every DDF the neighbours need (the extract kernels read layers 0/1 and nz-2/nz-1 only, sim_kernels.cl:1063-1069).
        // (2) The halo stream of every domain is forked: extract -> exchange over NVLink -> insert of fi (and rho/u/flags, fqi, ei) run
        // there.  (3) stream_collide on the interior layers runs meanwhile on the main stream -- insert only writes the halo cells'
        // own slots of this step's parity, which no thread touches before the next step -- followed by the LOD exchange and
        // update_e_b_dynamic, which never touch a DDF.  (4) join.
        for (auto& d : domains) {
            d.enqueue_stream_collide_range(1u, 2u, false);
            d.enqueue_stream_collide_range(d.n_z - 2u, d.n_z - 1u, false);
        }
        for (auto& d : domains) check(ion_halo_fork(d.dev));
        if (config.graphics_config.graphics_active) communicate_rho_u_flags();
        communicate_fi();
        if (mhd) {
            communicate_fqi();
            communicate_ei();
        }
        for (auto& d : domains) d.enqueue_stream_collide_range(2u, d.n_z - 2u, true);
        if (mhd) {
            build_lods_part_2();
            communicate_qu_lods();
            update_e_b_dynamic();
        }
        for (auto& d : domains) check(ion_halo_join(d.dev));
        if (mhd && world == 1) finish_queues();  // see below: cross-stream LOD copies of a single process
        increment_timestep(1);
        return;
    }
    stream_collide();
    if (config.graphics_config.graphics_active) communicate_rho_u_flags();
    if (mhd && get_d_n() > 1 && overlap_halo) {
        // Same operations as the reference sequence below, re-ordered so that they overlap: the LOD pyramids are gathered
        // and exchanged first (update_e_b_dynamic needs only them), then the fi / fqi / ei halo exchange runs on the
        // domains' halo streams while update_e_b_dynamic -- which never touches a DDF -- runs on the main streams.
        build_lods_part_2();
        communicate_qu_lods();
        for (auto& d : domains) check(ion_halo_fork(d.dev));
        communicate_fi();
        communicate_fqi();
        communicate_ei();
        update_e_b_dynamic();
        for (auto& d : domains) check(ion_halo_join(d.dev));
    } else {
        communicate_fi();
        if (mhd) {
            if (get_d_n() > 1) build_lods_part_2();
            communicate_fqi();
            communicate_ei();
            communicate_qu_lods();
            update_e_b_dynamic();
        }
    }
    // mod.rs:267-270 blocks here (`finish_queues`) in the single-domain and MHD cases.  On an in-order CUDA stream
    // the next step is ordered behind this one anyway, so the host does not stall; callers that need completion
    // call finish_queues() (run() does not need it, reads through ion_buffer_read synchronise by themselves).
    // Exception: several domains in ONE process exchange LOD pyramids by cross-stream copies, and the next step's
    // clear_qu_lod of a fast domain must not overtake a slow neighbour's copy -- keep the reference's barrier there.
    if (mhd && world == 1 && get_d_n() > 1) finish_queues();
    increment_timestep(1);
}

void Lbm::finish_queues() {
    for (auto& d : domains) d.finish();
}
void Lbm::precompute_B() {  // mod.rs:284-297
    if (!config.ext_magneto_hydro) return;
    for (auto& d : domains) d.enqueue_precompute_b();
    finish_queues();
}
void Lbm::precompute_E() {
    if (!config.ext_magneto_hydro) return;
    for (auto& d : domains) d.enqueue_precompute_e();
    finish_queues();
}
void Lbm::precompute_E_ECR() {
    if (!(config.ext_magneto_hydro && config.ext_subgrid_ecr)) return;
    for (auto& d : domains) d.enqueue_precompute_e_ecr();
    finish_queues();
}
void Lbm::kernel_initialize() { for (auto& d : domains) d.enqueue_initialize(); }
void Lbm::stream_collide() { for (auto& d : domains) d.enqueue_stream_collide(); }
void Lbm::update_e_b_dynamic() { for (auto& d : domains) d.enqueue_update_e_b_dyn(); }
void Lbm::build_lods_part_2() { for (auto& d : domains) d.enqueue_lod_part_2_gather(); }
void Lbm::clear_qu_lod() { for (auto& d : domains) d.enqueue_clear_qu_lod(); }

void Lbm::communicate_field(TransferField field, size_t bytes_per_cell) {  // mod.rs:371-407
    const uint32_t dxyz[3] = {config.d_x, config.d_y, config.d_z};
    const uint32_t d_x = config.d_x, d_y = config.d_y, d_z = config.d_z;
    const uint32_t d_n = (uint32_t)get_d_n();
    for (uint32_t axis = 0; axis < 3; axis++) {
        if (dxyz[axis] <= 1) continue;
        for (auto& d : domains) d.enqueue_transfer_extract_field(field, axis, bytes_per_cell);
        // the reference synchronises every queue here and swaps host vectors; the device-resident exchange is
        // ordered with events / NCCL stream semantics instead
        for (uint32_t d = 0; d < d_n; d++) {
            uint32_t dp, dm;  // ring neighbours, mod.rs:386-404
            check(ion_neighbor_domains(d_x, d_y, d_z, d, axis, &dp, &dm));
            if (world == 1) {
                const size_t bytes = domains[d].get_area(axis) * bytes_per_cell;
                check(ion_exchange_transfer(domains[d].dev, domains[dp].dev, bytes));
            } else if ((int)d == rank) {
                const size_t bytes = domains[0].get_area(axis) * bytes_per_cell;
                check(ion_comm_exchange_transfer(comm, domains[0].dev, (int)dp, (int)dm, bytes));
            }
        }
        for (auto& d : domains) d.enqueue_transfer_insert_field(field, axis, bytes_per_cell);
    }
}
void Lbm::communicate_fi() { communicate_field(TransferField::Fi, size_of(config.float_type) * get_transfers(config.velocity_set)); }
void Lbm::communicate_rho_u_flags() { communicate_field(TransferField::RhoUFlags, 17); }
void Lbm::communicate_fqi() { communicate_field(TransferField::Qi, size_of(config.float_type) * 1); }
void Lbm::communicate_ei() { communicate_field(TransferField::Ei, size_of(config.float_type) * get_transfers(config.velocity_set)); }

void Lbm::communicate_qu_lods() {  // mod.rs:436-468
    const uint32_t d_n = (uint32_t)get_d_n();
    if (d_n <= 1) return;
    if (world > 1) {
        check(ion_comm_exchange_lods(comm, domains[0].dev));
        return;
    }
    for (uint32_t d = 0; d < d_n; d++) {
        for (uint32_t dc = 0; dc < d_n; dc++) {
            uint32_t src = 0, cnt = 0, dst = 0;  // which pyramid level of dc lands where in d's QU_lod (mod.rs:448-465)
            check(ion_lod_exchange_plan(&domains[d].params, dc, &src, &cnt, &dst));
            if (cnt) check(ion_copy_lods(domains[d].dev, dst, domains[dc].dev, src, cnt));
        }
    }
}

void Lbm::increment_timestep(uint32_t steps) { for (auto& d : domains) d.t += steps; }
void Lbm::reset_timestep() { for (auto& d : domains) d.t = 0; }

}  // namespace ionhost
