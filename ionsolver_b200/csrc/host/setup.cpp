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

}  // namespace ionhost
