// host_api.cpp -- extern "C" face of the host layer (include/ionsolver_b200_host.h).  Exceptions of the C++ host
// become status codes + ion_last_error_string(), exactly like the device ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../../include/ionsolver_b200_host.h"
#include "lbm.hpp"

namespace ion {
int fail(int code, const char* fmt, ...);  // api.cu: stores the message for ion_last_error_string()
}
using namespace ionhost;

struct ion_lbm {
    Lbm* lbm;
};

namespace {
LbmConfig to_cpp(const IonLbmConfig& c) {
    LbmConfig o;
    o.velocity_set = (VelocitySet)c.velocity_set;
    o.relaxation_time = (RelaxationTime)c.relaxation_time;
    o.float_type = (FloatType)c.float_type;
    o.units.m = c.unit_m; o.units.kg = c.unit_kg; o.units.s = c.unit_s; o.units.a = c.unit_a; o.units.k = c.unit_k;
    o.units.prop = (Propellant)c.propellant;
    o.n_x = c.n_x; o.n_y = c.n_y; o.n_z = c.n_z;
    o.d_x = c.d_x; o.d_y = c.d_y; o.d_z = c.d_z;
    o.nu = c.nu;
    o.f_x = c.f_x; o.f_y = c.f_y; o.f_z = c.f_z;
    o.ext_equilibrium_boudaries = c.ext_equilibrium_boudaries; o.ext_volume_force = c.ext_volume_force; o.ext_force_field = c.ext_force_field;
    o.ext_magneto_hydro = c.ext_magneto_hydro; o.ext_subgrid_ecr = c.ext_subgrid_ecr;
    o.mhd_lod_depth = c.mhd_lod_depth;
    o.graphics_config.graphics_active = c.graphics_active;
    o.deterministic = c.deterministic != 0;
    o.ecr_freq = c.ecr_freq; o.ecr_field_strength = c.ecr_field_strength;
    o.run_steps = c.run_steps;
    return o;
}
void to_c(const LbmConfig& c, IonLbmConfig* o) {
    memset(o, 0, sizeof(*o));
    o->velocity_set = (uint32_t)c.velocity_set; o->relaxation_time = (uint32_t)c.relaxation_time; o->float_type = (uint32_t)c.float_type;
    o->unit_m = c.units.m; o->unit_kg = c.units.kg; o->unit_s = c.units.s; o->unit_a = c.units.a; o->unit_k = c.units.k;
    o->propellant = (uint32_t)c.units.prop;
    o->n_x = c.n_x; o->n_y = c.n_y; o->n_z = c.n_z;
    o->d_x = c.d_x; o->d_y = c.d_y; o->d_z = c.d_z;
    o->nu = c.nu;
    o->f_x = c.f_x; o->f_y = c.f_y; o->f_z = c.f_z;
    o->ext_equilibrium_boudaries = c.ext_equilibrium_boudaries; o->ext_volume_force = c.ext_volume_force; o->ext_force_field = c.ext_force_field;
    o->ext_magneto_hydro = c.ext_magneto_hydro; o->ext_subgrid_ecr = c.ext_subgrid_ecr;
    o->mhd_lod_depth = c.mhd_lod_depth;
    o->graphics_active = c.graphics_config.graphics_active;
    o->deterministic = c.deterministic;
    o->ecr_freq = c.ecr_freq; o->ecr_field_strength = c.ecr_field_strength;
    o->run_steps = c.run_steps;
}
int validate(const IonLbmConfig* c) {
    if (!c) return ion::fail(ION_ERR_INVALID, "NULL config");
    if (c->velocity_set > 3 || c->relaxation_time > 1 || c->float_type > 2 || c->propellant > 5) return ion::fail(ION_ERR_INVALID, "enum value out of range in IonLbmConfig");
    if (!c->d_x || !c->d_y || !c->d_z) return ion::fail(ION_ERR_INVALID, "zero domain count");
    return 0;
}
std::vector<int> devs(const int* d, int n) { return (d && n > 0) ? std::vector<int>(d, d + n) : std::vector<int>(); }
char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
int wrap_new(Lbm* l, ion_lbm_t** out) {
    ion_lbm* h = new ion_lbm{l};
    *out = h;
    return 0;
}
}  // namespace

#define ION_TRY(body)                                                       \
    try { body; return 0; }                                                 \
    catch (const IonException& e) { return ion::fail(e.code, "%s", e.what()); } \
    catch (const std::exception& e) { return ion::fail(ION_ERR_INVALID, "%s", e.what()); } \
    catch (...) { return ion::fail(ION_ERR_INVALID, "unknown exception"); }
#define ION_NEED(l) if (!(l) || !(l)->lbm) return ion::fail(ION_ERR_INVALID, "NULL lbm")

extern "C" {

void ion_lbm_config_default(IonLbmConfig* cfg) { if (cfg) to_c(LbmConfig(), cfg); }
void ion_units_set(IonLbmConfig* c, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8, float a9) {
    if (!c) return;
    Units u;
    u.set(a0, a1, a2, a3, a4, a5, a6, a7, a8, a9);
    c->unit_m = u.m; c->unit_kg = u.kg; c->unit_s = u.s; c->unit_a = u.a; c->unit_k = u.k;
}
float ion_units_eval(const IonLbmConfig* c, int fn, float v) {
    if (!c) return 0.0f;
    const Units u = to_cpp(*c).units;
    switch (fn) {
        case ION_UNIT_LEN_SI_LU: return u.len_si_lu(v);
        case ION_UNIT_NU_SI_LU: return u.nu_si_lu(v);
        case ION_UNIT_CHARGE_SI_LU: return u.charge_si_lu(v);
        case ION_UNIT_MAG_FLUX_SI_LU: return u.mag_flux_si_lu(v);
        case ION_UNIT_E_FIELD_SI_LU: return u.e_field_si_lu(v);
        case ION_UNIT_MAGNETIZATION_SI_LU: return u.magnetization_si_lu(v);
        case ION_UNIT_TIME_LU_SI: return u.time_lu_si(v);
        case ION_UNIT_TIME_SI_LU: return u.time_si_lu(v);
        case ION_UNIT_SPEED_SI_LU: return u.speed_si_lu(v);
        case ION_UNIT_LEN_LU_SI: return u.len_lu_si(v);
        case ION_UNIT_SPEED_LU_SI: return u.speed_lu_si(v);
        case ION_UNIT_CHARGE_LU_SI: return u.charge_lu_si(v);
        case ION_UNIT_MAG_FLUX_LU_SI: return u.mag_flux_lu_si(v);
        case ION_UNIT_E_FIELD_LU_SI: return u.e_field_lu_si(v);
        case ION_UNIT_EPSILON_0_LU: return u.epsilon_0_lu();
        case ION_UNIT_KE_LU: return u.ke_lu();
        case ION_UNIT_MU_0_LU: return u.mu_0_lu();
        case ION_UNIT_KKGE_LU: return u.kkge_lu();
        case ION_UNIT_KIMG_LU: return u.kimg_lu();
        case ION_UNIT_KVEV_LU: return u.kveV_lu();
        case ION_UNIT_KKBME_LU: return u.kkBme_lu();
        case ION_UNIT_KEABS_LU: return u.keabs_lu();
        case ION_UNIT_KME_LU: return u.kme_lu();
    }
    return 0.0f;
}
int ion_lbm_make_params(const IonLbmConfig* c, uint32_t d, IonParams* out) {
    int r = validate(c);
    if (r) return r;
    if (!out) return ion::fail(ION_ERR_INVALID, "NULL out");
    LbmConfig cfg = to_cpp(*c);
    cfg.n_x = (cfg.n_x / cfg.d_x) * cfg.d_x; cfg.n_y = (cfg.n_y / cfg.d_y) * cfg.d_y; cfg.n_z = (cfg.n_z / cfg.d_z) * cfg.d_z;
    if (d >= cfg.d_x * cfg.d_y * cfg.d_z) return ion::fail(ION_ERR_INVALID, "domain index out of range");
    const uint32_t x = (d % (cfg.d_x * cfg.d_y)) % cfg.d_x, y = (d % (cfg.d_x * cfg.d_y)) / cfg.d_x, z = d / (cfg.d_x * cfg.d_y);
    *out = LbmDomain::make_params(cfg, x, y, z, d);
    return 0;
}

int ion_lbm_create(const IonLbmConfig* c, const int* devices, int n_devices, ion_lbm_t** out) {
    int r = validate(c);
    if (r) return r;
    if (!out) return ion::fail(ION_ERR_INVALID, "NULL out");
    *out = nullptr;
    ION_TRY(wrap_new(Lbm::create(to_cpp(*c), devs(devices, n_devices)), out))
}
int ion_lbm_create_distributed(const IonLbmConfig* c, int rank, int world, int device, const uint8_t comm_id[ION_COMM_ID_BYTES], ion_lbm_t** out) {
    int r = validate(c);
    if (r) return r;
    if (!out) return ion::fail(ION_ERR_INVALID, "NULL out");
    *out = nullptr;
    ION_TRY(wrap_new(Lbm::create_distributed(to_cpp(*c), rank, world, device, comm_id), out))
}
int ion_lbm_destroy(ion_lbm_t* l) {
    if (!l) return 0;
    delete l->lbm;
    delete l;
    return 0;
}
int ion_lbm_get_config(const ion_lbm_t* l, IonLbmConfig* out) { ION_NEED(l); if (!out) return ion::fail(ION_ERR_INVALID, "NULL out"); to_c(l->lbm->config, out); return 0; }
int ion_lbm_local_domains(const ion_lbm_t* l, uint32_t* count) { ION_NEED(l); if (!count) return ion::fail(ION_ERR_INVALID, "NULL out"); *count = (uint32_t)l->lbm->domains.size(); return 0; }
int ion_lbm_domain(ion_lbm_t* l, uint32_t i, ion_domain_t** out, uint32_t* domain_index) {
    ION_NEED(l);
    if (!out || i >= l->lbm->domains.size()) return ion::fail(ION_ERR_INVALID, "domain index out of range");
    *out = l->lbm->domains[i].dev;
    if (domain_index) *domain_index = l->lbm->domains[i].d_i;
    return 0;
}
int ion_lbm_initialize(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->initialize()) }
int ion_lbm_run(ion_lbm_t* l, uint64_t steps) { ION_NEED(l); ION_TRY(l->lbm->run(steps)) }
int ion_lbm_do_time_step(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->do_time_step()) }
int ion_lbm_finish_queues(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->finish_queues()) }
int ion_lbm_get_time_step(const ion_lbm_t* l, uint64_t* t) { ION_NEED(l); if (!t) return ion::fail(ION_ERR_INVALID, "NULL out"); *t = l->lbm->get_time_step(); return 0; }
int ion_lbm_set_time_step(ion_lbm_t* l, uint64_t t) { ION_NEED(l); for (auto& d : l->lbm->domains) d.t = t; return 0; }
int ion_lbm_precompute_b(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->precompute_B()) }
int ion_lbm_precompute_e(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->precompute_E()) }
int ion_lbm_precompute_e_ecr(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->precompute_E_ECR()) }
int ion_lbm_communicate_field(ion_lbm_t* l, int field) {
    ION_NEED(l);
    ION_TRY({
        switch (field) {
            case ION_TRANSFER_FI: l->lbm->communicate_fi(); break;
            case ION_TRANSFER_RHO_U_FLAGS: l->lbm->communicate_rho_u_flags(); break;
            case ION_TRANSFER_EI: l->lbm->communicate_ei(); break;
            case ION_TRANSFER_QI: l->lbm->communicate_fqi(); break;
            default: throw IonException(ION_ERR_INVALID, "unknown transfer field");
        }
    })
}
int ion_lbm_communicate_qu_lods(ion_lbm_t* l) { ION_NEED(l); ION_TRY(l->lbm->communicate_qu_lods()) }

int ion_lbm_import_mesh(ion_lbm_t* l, const char* path, float scale, float ox, float oy, float oz, float rx, float ry, float rz) {
    ION_NEED(l);
    if (!path) return ion::fail(ION_ERR_INVALID, "NULL path");
    ION_TRY(l->lbm->import_mesh(path, scale, ox, oy, oz, rx, ry, rz))
}
int ion_lbm_import_mesh_reposition(ion_lbm_t* l, const char* path, float cx, float cy, float cz, float rx, float ry, float rz, float size) {
    ION_NEED(l);
    if (!path) return ion::fail(ION_ERR_INVALID, "NULL path");
    ION_TRY(l->lbm->import_mesh_reposition(path, cx, cy, cz, rx, ry, rz, size))
}
int ion_lbm_voxelise_mesh(ion_lbm_t* l, uint32_t index, int model_type, float v0, float v1, float v2) {
    ION_NEED(l);
    ModelType t;
    switch (model_type) {
        case ION_MODEL_SOLID: t = ModelType::solid(); break;
        case ION_MODEL_MAGNET: t = ModelType::magnet(v0, v1, v2); break;
        case ION_MODEL_CHARGED: t = ModelType::charged(v0); break;
        case ION_MODEL_CHARGED_ECR: t = ModelType::charged_ecr(v0); break;
        default: return ion::fail(ION_ERR_INVALID, "unknown model type %d", model_type);
    }
    ION_TRY(l->lbm->voxelise_mesh(index, t))
}
int ion_lbm_mesh_info(const ion_lbm_t* l, uint32_t index, uint32_t* tn, float p_min[3], float p_max[3]) {
    ION_NEED(l);
    if (index >= l->lbm->meshes.size()) return ion::fail(ION_ERR_INVALID, "mesh index out of range");
    const Mesh& m = l->lbm->meshes[index];
    if (tn) *tn = m.triangle_number;
    if (p_min) { p_min[0] = m.p_min.x; p_min[1] = m.p_min.y; p_min[2] = m.p_min.z; }
    if (p_max) { p_max[0] = m.p_max.x; p_max[1] = m.p_max.y; p_max[2] = m.p_max.z; }
    return 0;
}
int ion_lbm_mesh_triangles(const ion_lbm_t* l, uint32_t index, float* p0, float* p1, float* p2) {
    ION_NEED(l);
    if (index >= l->lbm->meshes.size() || !p0 || !p1 || !p2) return ion::fail(ION_ERR_INVALID, "bad mesh query");
    const Mesh& m = l->lbm->meshes[index];
    for (uint32_t i = 0; i < m.triangle_number; i++) {
        p0[3 * i] = m.p0[i].x; p0[3 * i + 1] = m.p0[i].y; p0[3 * i + 2] = m.p0[i].z;
        p1[3 * i] = m.p1[i].x; p1[3 * i + 1] = m.p1[i].y; p1[3 * i + 2] = m.p1[i].z;
        p2[3 * i] = m.p2[i].x; p2[3 * i + 1] = m.p2[i].y; p2[3 * i + 2] = m.p2[i].z;
    }
    return 0;
}
int ion_lbm_mesh_translate(ion_lbm_t* l, uint32_t index, float tx, float ty, float tz) {
    ION_NEED(l);
    if (index >= l->lbm->meshes.size()) return ion::fail(ION_ERR_INVALID, "mesh index out of range");
    F32_3 t; t.x = tx; t.y = ty; t.z = tz;
    l->lbm->meshes[index].translate(t);
    return 0;
}

int ion_lbm_set_taylor_green(ion_lbm_t* l, uint32_t periodicity) { ION_NEED(l); if (!periodicity) return ion::fail(ION_ERR_INVALID, "periodicity 0"); ION_TRY(l->lbm->set_taylor_green(periodicity)) }
int ion_lbm_setup_velocity_field(ion_lbm_t* l, float vx, float vy, float vz, float density) { ION_NEED(l); ION_TRY(l->lbm->setup_velocity_field(vx, vy, vz, density)) }
int ion_setup_taylor_green(uint32_t n, uint32_t d_z, int vs, int ft, int graphics_active, const int* devices, int n_devices, ion_lbm_t** out) {
    if (!out || !n || !d_z || vs < 0 || vs > 3 || ft < 0 || ft > 2) return ion::fail(ION_ERR_INVALID, "bad argument");
    *out = nullptr;
    ION_TRY(wrap_new(setup_taylor_green(n, d_z, (VelocitySet)vs, (FloatType)ft, graphics_active != 0, devs(devices, n_devices)), out))
}
int ion_setup_lid_driven_cavity(uint32_t n, const int* devices, int n_devices, ion_lbm_t** out) {
    if (!out || n < 3) return ion::fail(ION_ERR_INVALID, "bad argument");
    *out = nullptr;
    ION_TRY(wrap_new(setup_lid_driven_cavity(n, devs(devices, n_devices)), out))
}
int ion_setup_charged_fluid(uint32_t nx, uint32_t ny, uint32_t nz, int vs, int ft, uint32_t lod_depth, const char* magnet_stl, const int* devices,
                            int n_devices, ion_lbm_t** out) {
    if (!out || !nx || !ny || !nz || vs < 1 || vs > 3 || ft < 0 || ft > 2 || lod_depth > 4) return ion::fail(ION_ERR_INVALID, "bad argument");
    *out = nullptr;
    ION_TRY(wrap_new(setup_charged_fluid(nx, ny, nz, (VelocitySet)vs, (FloatType)ft, (uint8_t)lod_depth, magnet_stl ? magnet_stl : "", devs(devices, n_devices)), out))
}

int ion_setup_scene(const char* name, const char* stl_dir, float scale, uint32_t flags, const int* devices, int n_devices, ion_lbm_t** out) {
    if (!out || !name) return ion::fail(ION_ERR_INVALID, "NULL argument");
    *out = nullptr;
    ION_TRY(wrap_new(setup_scene(name, stl_dir ? stl_dir : "stl", scale, flags, devs(devices, n_devices)), out))
}

int ion_lbm_encode(ion_lbm_t* l, int reference_compatible, uint8_t** data, size_t* len) {
    ION_NEED(l);
    if (!data || !len) return ion::fail(ION_ERR_INVALID, "NULL out");
    ION_TRY({
        const std::vector<uint8_t> b = file::encode(*l->lbm, reference_compatible != 0);
        *data = (uint8_t*)malloc(b.size() ? b.size() : 1);
        if (!*data) throw IonException(ION_ERR_INVALID, "out of host memory");
        memcpy(*data, b.data(), b.size());
        *len = b.size();
    })
}
int ion_lbm_decode(const uint8_t* data, size_t len, IonLbmConfig* cfg, int reference_compatible, const int* devices, int n_devices, ion_lbm_t** out) {
    if (!data || !cfg || !out) return ion::fail(ION_ERR_INVALID, "NULL argument");
    *out = nullptr;
    LbmConfig c = to_cpp(*cfg);
    struct WriteBack {  // decode fills `config` in place as it parses (file.rs:54-104), also when it fails later
        LbmConfig& c; IonLbmConfig* o; Lbm* lbm = nullptr;
        ~WriteBack() { to_c(lbm ? lbm->config : c, o); }
    } wb{c, cfg};
    ION_TRY({
        wb.lbm = file::decode(std::vector<uint8_t>(data, data + len), c, reference_compatible != 0, devs(devices, n_devices));
        wrap_new(wb.lbm, out);
    })
}
int ion_lbm_write_file(ion_lbm_t* l, const char* path) { ION_NEED(l); if (!path) return ion::fail(ION_ERR_INVALID, "NULL path"); ION_TRY(file::write(*l->lbm, path)) }
int ion_lbm_read_file(const char* path, IonLbmConfig* cfg, ion_lbm_t** out) {
    if (!path || !cfg || !out) return ion::fail(ION_ERR_INVALID, "NULL argument");
    *out = nullptr;
    ION_TRY({
        LbmConfig c = to_cpp(*cfg);
        Lbm* lbm = file::read(path, c);
        to_c(lbm->config, cfg);
        wrap_new(lbm, out);
    })
}
int ion_config_to_json(const IonLbmConfig* cfg, char** json) {
    int r = validate(cfg);
    if (r) return r;
    if (!json) return ion::fail(ION_ERR_INVALID, "NULL out");
    ION_TRY(*json = dup_string(file::config_to_json(to_cpp(*cfg))))
}
int ion_config_from_json(const char* json, IonLbmConfig* cfg) {
    if (!json || !cfg) return ion::fail(ION_ERR_INVALID, "NULL argument");
    ION_TRY(to_c(file::config_from_json(json), cfg))
}
int ion_lbm_dump_cell(ion_lbm_t* l, uint32_t local_index, uint64_t cell, char** text) {
    ION_NEED(l);
    if (!text || local_index >= l->lbm->domains.size() || cell >= l->lbm->domains[local_index].n) return ion::fail(ION_ERR_INVALID, "bad dump_cell query");
    ION_TRY(*text = dup_string(l->lbm->domains[local_index].dump_cell(cell)))
}
int ion_lbm_read_slice(ion_lbm_t* l, int field, int component, uint32_t slice_mode, uint32_t index, float* out, size_t capacity,
                       uint32_t* width, uint32_t* height) {
    ION_NEED(l);
    if (!out || !width || !height) return ion::fail(ION_ERR_INVALID, "NULL argument");
    ION_TRY({
        std::vector<float> plane;
        slice::read(*l->lbm, field, component, slice_mode, index, plane, *width, *height);
        if (plane.size() > capacity) throw IonException(ION_ERR_RANGE, "slice needs " + std::to_string(plane.size()) + " floats");
        memcpy(out, plane.data(), plane.size() * sizeof(float));
    })
}
int ion_lbm_write_slice_png(ion_lbm_t* l, int field, int component, uint32_t slice_mode, uint32_t index, float v_min, float v_max,
                            const char* path) {
    ION_NEED(l);
    if (!path) return ion::fail(ION_ERR_INVALID, "NULL path");
    ION_TRY(slice::write_png(*l->lbm, field, component, slice_mode, index, v_min, v_max, path))
}
uint32_t ion_iron_colormap(float x) { return slice::iron_colormap(x); }
int ion_write_png_rgb(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb) return ion::fail(ION_ERR_INVALID, "NULL argument");
    ION_TRY({
        const std::vector<uint8_t> png = slice::encode_png_rgb(rgb, width, height);
        FILE* f = fopen(path, "wb");
        if (!f) throw IonException(ION_ERR_INVALID, std::string("cannot open \"") + path + "\" for writing");
        const size_t n = fwrite(png.data(), 1, png.size(), f);
        fclose(f);
        if (n != png.size()) throw IonException(ION_ERR_INVALID, "short write");
    })
}
void ion_free(void* p) { free(p); }

}  // extern "C"
