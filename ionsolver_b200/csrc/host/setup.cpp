// setup.cpp -- scene helpers of /root/reference/src/setup.rs: the two field initialisers every scene uses
// (set_taylor_green :458-543, setup_velocity_field :547-598) and constructors for the BASELINE.json configurations,
// which the reference expresses by editing setup() (it has no CLI).
#include <cmath>
#include <cstring>

#include "lbm.hpp"

namespace ionhost {

static const float PI_F = 3.14159274101257324219f;

// setup.rs:458-543.  f32 arithmetic in source order; the density line keeps the reference's missing parentheses
// (SURVEY quirk Q11): rho = 1 - A^2*3/4*cos(4 pi x/a) + cos(4 pi y/b).
void Lbm::set_taylor_green(uint32_t periodicity) {
    const uint32_t nx = config.n_x, ny = config.n_y, nz = config.n_z;
    const float A = 0.25f;
    const float a = (float)nx / (float)periodicity, b = (float)ny / (float)periodicity, c = (float)nz / (float)periodicity;
    const uint32_t dx = config.d_x, dy = config.d_y, dz = config.d_z;
    const uint64_t dsx = nx / dx + (dx > 1u) * 2u, dsy = ny / dy + (dy > 1u) * 2u, dsz = nz / dz + (dz > 1u) * 2u;
    const uint64_t dtotal = dsx * dsy * dsz;
    for (uint32_t d = 0; d < dx * dy * dz; d++) {
        LbmDomain* dom = local_domain(d);
        if (!dom) continue;
        const uint32_t x = (d % (dx * dy)) % dx, y = (d % (dx * dy)) / dx, z = d / (dx * dy);
        std::vector<float> u(dtotal * 3, 0.0f), rho(dtotal, 0.0f);
        for (uint64_t zi = 0; zi < dsz; zi++)
            for (uint64_t yi = 0; yi < dsy; yi++)
                for (uint64_t xi = 0; xi < dsx; xi++) {
                    if (((xi == 0 || xi == dsx - 1) && dx > 1) || ((yi == 0 || yi == dsy - 1) && dy > 1) || ((zi == 0 || zi == dsz - 1) && dz > 1)) continue;
                    const uint64_t dn = zi * dsx * dsy + yi * dsx + xi;
                    const uint64_t gx = xi - (dx > 1u) + (uint64_t)x * (dsx - (dx > 1u) * 2u);
                    const uint64_t gy = yi - (dy > 1u) + (uint64_t)y * (dsy - (dy > 1u) * 2u);
                    const uint64_t gz = zi - (dz > 1u) + (uint64_t)z * (dsz - (dz > 1u) * 2u);
                    const float fx = (float)gx + 0.5f - 0.5f * (float)nx;
                    const float fy = (float)gy + 0.5f - 0.5f * (float)ny;
                    const float fz = (float)gz + 0.5f - 0.5f * (float)nz;
                    u[dn] = A * cosf(2.0f * PI_F * fx / a) * sinf(2.0f * PI_F * fy / b) * sinf(2.0f * PI_F * fz / c);
                    u[dn + dtotal] = -A * sinf(2.0f * PI_F * fx / a) * cosf(2.0f * PI_F * fy / b) * sinf(2.0f * PI_F * fz / c);
                    u[dn + dtotal * 2] = A * sinf(2.0f * PI_F * fx / a) * sinf(2.0f * PI_F * fy / b) * cosf(2.0f * PI_F * fz / c);
                    rho[dn] = 1.0f - (A * A) * 3.0f / 4.0f * cosf(4.0f * PI_F * fx / a) + cosf(4.0f * PI_F * fy / b);
                }
        dom->write(ION_FIELD_U, u.data(), u.size() * 4);
        dom->write(ION_FIELD_RHO, rho.data(), rho.size() * 4);
        dom->finish();
    }
}

// setup.rs:547-598
void Lbm::setup_velocity_field(float vx, float vy, float vz, float density) {
    const uint32_t dx = config.d_x, dy = config.d_y, dz = config.d_z;
    const uint64_t dsx = config.n_x / dx + (dx > 1u) * 2u, dsy = config.n_y / dy + (dy > 1u) * 2u, dsz = config.n_z / dz + (dz > 1u) * 2u;
    const uint64_t dtotal = dsx * dsy * dsz;
    for (uint32_t d = 0; d < dx * dy * dz; d++) {
        LbmDomain* dom = local_domain(d);
        if (!dom) continue;
        std::vector<float> u(dtotal * 3, 0.0f), rho(dtotal, 0.0f);
        for (uint64_t zi = 0; zi < dsz; zi++)
            for (uint64_t yi = 0; yi < dsy; yi++)
                for (uint64_t xi = 0; xi < dsx; xi++) {
                    if (((xi == 0 || xi == dsx - 1) && dx > 1) || ((yi == 0 || yi == dsy - 1) && dy > 1) || ((zi == 0 || zi == dsz - 1) && dz > 1)) continue;
                    const uint64_t dn = zi * dsx * dsy + yi * dsx + xi;
                    u[dn] = vx; u[dn + dtotal] = vy; u[dn + dtotal * 2] = vz;
                    rho[dn] = density;
                }
        dom->write(ION_FIELD_U, u.data(), u.size() * 4);
        dom->write(ION_FIELD_RHO, rho.data(), rho.size() * 4);
        dom->finish();
    }
}

// setup_taylor_green (setup.rs:92-113) / setup_domain_test (setup.rs:115-139), size and split as parameters
Lbm* setup_taylor_green(uint32_t n, uint32_t d_z, VelocitySet vs, FloatType ft, bool graphics_active, const std::vector<int>& devices) {
    LbmConfig cfg;
    cfg.n_x = cfg.n_y = cfg.n_z = n;
    cfg.d_z = d_z;
    cfg.nu = cfg.units.nu_si_lu(0.1f);
    cfg.velocity_set = vs;
    cfg.float_type = ft;
    cfg.graphics_config.graphics_active = graphics_active;
    Lbm* lbm = Lbm::create(cfg, devices);
    lbm->set_taylor_green(1);
    return lbm;
}

// BASELINE config 1 (ii), not shipped by the reference: solid walls x=0, x=n-1, y=0, y=n-1, z=0 and a
// TYPE_E lid at z=n-1 moving with u=(0.1,0,0), rho=1; equilibrium boundaries on, nu=0.01.
Lbm* setup_lid_driven_cavity(uint32_t n, const std::vector<int>& devices) {
    LbmConfig cfg;
    cfg.n_x = cfg.n_y = cfg.n_z = n;
    cfg.nu = 0.01f;
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.float_type = FloatType::FP32;
    cfg.ext_equilibrium_boudaries = true;
    cfg.graphics_config.graphics_active = false;
    Lbm* lbm = Lbm::create(cfg, devices);
    const uint64_t N = (uint64_t)n * n * n;
    std::vector<uint8_t> flags(N, 0);
    std::vector<float> u(3 * N, 0.0f);
    for (uint32_t z = 0; z < n; z++)
        for (uint32_t y = 0; y < n; y++)
            for (uint32_t x = 0; x < n; x++) {
                const uint64_t i = x + (y + (uint64_t)z * n) * n;
                if (z == n - 1) { flags[i] = ION_TYPE_E; u[i] = 0.1f; }
                else if (x == 0 || x == n - 1 || y == 0 || y == n - 1 || z == 0) flags[i] = ION_TYPE_S;
            }
    lbm->domains[0].write(ION_FIELD_FLAGS, flags.data(), N);
    lbm->domains[0].write(ION_FIELD_U, u.data(), 3 * N * 4);
    return lbm;
}

// BASELINE config 2: charged fluid (setup_bfield_spin, setup.rs:142-201: units, Q = 0.002 per cell, u = (0.1, 0.01, 0))
// in the static field of the voxelised disk magnet (setup_mesh_field_test, setup.rs:346-393: import_mesh_reposition +
// voxelise_mesh(Magnet{(0,1e6,0)}) + precompute_B).  Lengths scale with nx/128.  magnet_stl empty = no magnet.
Lbm* setup_charged_fluid(uint32_t nx, uint32_t ny, uint32_t nz, VelocitySet vs, FloatType ft, uint8_t lod_depth,
                         const std::string& magnet_stl, const std::vector<int>& devices) {
    LbmConfig cfg;
    cfg.units.set((float)nx, 1.0f, 1.0f, 1.0f, 1.0f, 0.1f, 1.0f, 1.2250f, 0.0000000001f, 1.0f);
    cfg.n_x = nx; cfg.n_y = ny; cfg.n_z = nz;
    cfg.nu = cfg.units.nu_si_lu(1.48E-5f);
    cfg.velocity_set = vs;
    cfg.float_type = ft;
    cfg.mhd_lod_depth = lod_depth;
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.graphics_config.graphics_active = false;
    Lbm* lbm = Lbm::create(cfg, devices);
    if (!magnet_stl.empty()) {
        lbm->import_mesh_reposition(magnet_stl, 0.5f * (float)nx + 0.1f, 0.5f * (float)ny + 0.1f, 0.5f * (float)nz, 0.0f, 0.0f, 0.0f,
                                    0.5f * (float)nx - 1.0f);
        lbm->voxelise_mesh(0, ModelType::magnet(0.0f, 1000000.0f, 0.0f));
        lbm->precompute_B();
    }
    for (auto& d : lbm->domains) {
        std::vector<float> charge(d.n, 0.002f);
        d.write(ION_FIELD_Q, charge.data(), charge.size() * 4);
    }
    lbm->setup_velocity_field(0.1f, 0.01f, 0.0f, 1.0f);
    return lbm;
}

// ---------------------------------------------------------------------------------------------------------------
// The reference's remaining scene functions (setup.rs), selectable by name.  Each keeps the reference's units, sizes, mesh
// placements and model types; what belongs to the rasteriser (camera, colours, streamline settings) has no effect on the path.
// `stl_dir` replaces the hard-coded "stl/" prefix.  `scale` multiplies all lengths of the two mesh scenes (1 = the reference's
// 128 x 256 x 128; BASELINE cfg3 uses 2); `subgrid_ecr` overrides setup_deeva_test's ext_subgrid_ecr = true, which overflows
// within two steps in the reference itself -- pass false for runs, true for the literal scene.
// ---------------------------------------------------------------------------------------------------------------
static bool file_exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (f) fclose(f);
    return f != nullptr;
}

Lbm* setup_verification(const std::vector<int>& devices) {  // setup.rs:203-241: 1 C moving at 1 m/s in cell 0
    LbmConfig cfg;
    cfg.units.set(1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.5f, 1.0f, 1.0f, 10.0f, 1.0f);
    cfg.n_x = cfg.n_y = cfg.n_z = 128;
    cfg.nu = cfg.units.nu_si_lu(1.48E-5f);
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.mhd_lod_depth = 4;
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.graphics_config.graphics_active = true;
    Lbm* lbm = Lbm::create(cfg, devices);
    const uint64_t N = (uint64_t)cfg.n_x * cfg.n_y * cfg.n_z;
    std::vector<float> charge(N, 0.0f), vel(3 * N, 0.0f), rho(N, 1.0f);
    charge[0] = 1.0f;
    vel[0] = 1.0f;
    lbm->domains[0].write(ION_FIELD_Q, charge.data(), N * 4);
    lbm->domains[0].write(ION_FIELD_U, vel.data(), 3 * N * 4);
    lbm->domains[0].write(ION_FIELD_RHO, rho.data(), N * 4);
    return lbm;
}

Lbm* setup_field_vis(const std::vector<int>& devices) {  // setup.rs:244-277 (its magnets are commented out in the reference)
    LbmConfig cfg;
    cfg.units.set(128.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.1f, 1.0f, 1.2250f, 0.0000000001f, 1.0f);
    cfg.n_x = cfg.n_y = cfg.n_z = 256;
    cfg.nu = cfg.units.nu_si_lu(1.48E-5f);
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.mhd_lod_depth = 4;
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.graphics_config.graphics_active = true;
    return Lbm::create(cfg, devices);
}

Lbm* setup_ecr_test(const std::vector<int>& devices) {  // setup.rs:280-317
    LbmConfig cfg;
    cfg.units.set(1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.01f, 10000.0f, 10E-8f, 0.0000000000001f, 50000.0f);
    cfg.n_x = cfg.n_y = cfg.n_z = 10;
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.ecr_freq = cfg.units.time_lu_si(2.45E9f);
    cfg.mhd_lod_depth = 1;
    Lbm* lbm = Lbm::create(cfg, devices);
    std::vector<float> b(3000, 0.0f);
    const float by = cfg.units.mag_flux_si_lu(0.0875f);
    for (int i = 1000; i < 2000; i++) b[i] = by;
    lbm->domains[0].write(ION_FIELD_B_STAT, b.data(), b.size() * 4);
    return lbm;
}

Lbm* setup_mesh_test(const std::string& stl_dir, const std::vector<int>& devices) {  // setup.rs:320-343
    LbmConfig cfg;
    cfg.n_x = 128; cfg.n_y = 256; cfg.n_z = 128;
    cfg.nu = cfg.units.nu_si_lu(0.05f);
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.graphics_config.graphics_active = true;
    Lbm* lbm = Lbm::create(cfg, devices);
    try {
        size_t next = 0;
        // stl/cow.stl is git-ignored in the reference (.gitignore:1) and absent from most checkouts: optional here
        if (file_exists(stl_dir + "/cow.stl")) {
            lbm->import_mesh_reposition(stl_dir + "/cow.stl", 64.0f, 128.0f, 64.0f, 0.0f, 0.0f, 0.0f, 200.0f);
            F32_3 off;
            off.z = 128.0f - lbm->meshes[0].p_max.z;
            lbm->meshes[0].translate(off);
            lbm->voxelise_mesh(next++, ModelType::solid());
        }
        lbm->import_mesh_reposition(stl_dir + "/ring-magnet.stl", 64.5f, 10.0f, 64.0f, 0.0f, 0.0f, 0.0f, 127.0f);
        lbm->voxelise_mesh(next, ModelType::magnet(0.0f, 0.0f, 0.0f));
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

Lbm* setup_mesh_field_test(const std::string& stl_dir, float scale, const std::vector<int>& devices) {  // setup.rs:346-393
    LbmConfig cfg;
    cfg.units.set(128.0f * scale, 1.0f, 1.0f, 1.0f, 1.0f, 0.1f, 1.0f, 1.2250f, 0.0000000001f, 1.0f);
    cfg.n_x = (uint32_t)(128.0f * scale); cfg.n_y = (uint32_t)(256.0f * scale); cfg.n_z = (uint32_t)(128.0f * scale);
    cfg.nu = cfg.units.nu_si_lu(0.05f);
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.ecr_freq = cfg.units.time_lu_si(2.45E9f);
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.graphics_config.graphics_active = true;
    Lbm* lbm = Lbm::create(cfg, devices);
    try {
        lbm->import_mesh_reposition(stl_dir + "/disk-magnet.stl", 64.1f * scale, 246.1f * scale, 64.0f * scale, 0.0f, 0.0f, 0.0f, 127.0f * scale);
        lbm->import_mesh(stl_dir + "/ring-magnet.stl", 1.0f, 64.1f * scale, 64.1f * scale, 64.0f * scale, 0.0f, 0.0f, 0.0f);
        lbm->voxelise_mesh(0, ModelType::magnet(0.0f, 1000000.0f, 0.0f));
        lbm->voxelise_mesh(1, ModelType::magnet(0.0f, 1000000.0f, 0.0f));
        // lbm.precompute_B() is commented out in the reference (setup.rs:388)
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

Lbm* setup_deeva_test(const std::string& stl_dir, float scale, bool subgrid_ecr, bool first_step, const std::vector<int>& devices) {  // setup.rs:395-453
    LbmConfig cfg;
    cfg.units.set(128.0f * scale, 1.0f, 1.0f, 1.0f, 1.0f, 0.1f, 1.0f, 10e-8f, 1.0f, 50000.0f);
    cfg.n_x = (uint32_t)(128.0f * scale); cfg.n_y = (uint32_t)(256.0f * scale); cfg.n_z = (uint32_t)(128.0f * scale);
    cfg.nu = cfg.units.nu_si_lu(0.05f);
    cfg.velocity_set = VelocitySet::D3Q19;
    cfg.ecr_freq = cfg.units.time_lu_si(2.45E9f);
    cfg.ext_volume_force = true;
    cfg.ext_magneto_hydro = true;
    cfg.ext_subgrid_ecr = subgrid_ecr;
    cfg.graphics_config.graphics_active = true;
    Lbm* lbm = Lbm::create(cfg, devices);
    try {
        const float c = 64.0f * scale;
        lbm->import_mesh(stl_dir + "/deeva_disk_magnet.stl", 1.0f, 64.001f * scale, 0.0f, c, 0.0f, 0.0f, 0.0f);
        lbm->import_mesh(stl_dir + "/deeva_inlet.stl", 1.0f, c, 0.0f, c, 0.0f, 0.0f, 0.0f);
        lbm->import_mesh(stl_dir + "/deeva_quartz_tube.stl", 1.0f, 64.001f * scale, 0.0f, c, 0.0f, 0.0f, 0.0f);
        lbm->import_mesh(stl_dir + "/deeva_ring_magnet.stl", 1.0f, 64.001f * scale, -0.5f * scale, c, 0.0f, 0.0f, 0.0f);
        lbm->import_mesh(stl_dir + "/deeva_e_plate1.stl", 1.0f, c, 0.0f, c, 0.0f, 0.0f, 0.0f);
        lbm->import_mesh(stl_dir + "/deeva_e_plate2.stl", 1.0f, c, 0.0f, c, 0.0f, 0.0f, 0.0f);
        lbm->voxelise_mesh(0, ModelType::magnet(0.0f, 1000000.0f, 0.0f));
        lbm->voxelise_mesh(1, ModelType::solid());
        lbm->voxelise_mesh(2, ModelType::solid());
        lbm->voxelise_mesh(3, ModelType::magnet(0.0f, 500000.0f, 0.0f));
        // 2000 V over 0.05 m, 0.0015 m^2: Q = C V = 5.3e-10 C over 2432 plate cells (setup.rs:433-440)
        lbm->voxelise_mesh(4, ModelType::charged_ecr(0.00000000000021844213f / 2.0f));
        lbm->voxelise_mesh(5, ModelType::charged_ecr(-0.00000000000021844213f / 2.0f));
        lbm->precompute_B();
        if (subgrid_ecr) lbm->precompute_E_ECR();  // E_var only exists with ext_subgrid_ecr (domain.rs:200-211)
        else lbm->precompute_E();                   // "static E from plates" for BASELINE cfg3
        if (first_step) {
            lbm->initialize();
            lbm->do_time_step();
        }
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

// dispatch by the reference's function name
Lbm* setup_scene(const std::string& name, const std::string& stl_dir, float scale, uint32_t flags, const std::vector<int>& devices) {
    if (!(scale > 0.0f)) scale = 1.0f;
    if (name == "setup_verification") return setup_verification(devices);
    if (name == "setup_field_vis") return setup_field_vis(devices);
    if (name == "setup_ecr_test") return setup_ecr_test(devices);
    if (name == "setup_mesh_test") return setup_mesh_test(stl_dir, devices);
    if (name == "setup_mesh_field_test") return setup_mesh_field_test(stl_dir, scale, devices);
    if (name == "setup_deeva_test") return setup_deeva_test(stl_dir, scale, (flags & 1u) != 0, (flags & 2u) != 0, devices);
    if (name == "setup_taylor_green") return setup_taylor_green(256, 1, VelocitySet::D3Q19, FloatType::FP16S, true, devices);  // setup.rs:92-113
    if (name == "setup_domain_test") return setup_taylor_green(256, 2, VelocitySet::D3Q19, FloatType::FP16S, true, devices);   // setup.rs:115-139
    throw IonException(ION_ERR_INVALID, "unknown scene \"" + name + "\" (setup.rs:22-64)");
}

}  // namespace ionhost
