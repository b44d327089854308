"""Scratch GPU check: CUDA kernels vs the shim-compiled reference kernels (oracle/_ref). Run with --build-only on the
CPU box first so that the reference libraries exist and travel to the GPU box."""
import sys, os, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_host as rh
build_only = "--build-only" in sys.argv

def params_from(cfg, g):
    from ionsolver_b200 import capi
    p = capi.IonParams()
    p.abi_version = 1
    p.nx, p.ny, p.nz = g.n_x, g.n_y, g.n_z
    p.dx, p.dy, p.dz, p.di = cfg.d_x, cfg.d_y, cfg.d_z, g.d_i
    p.ox, p.oy, p.oz = g.o_x, g.o_y, g.o_z
    p.velocity_set = capi.VELOCITY_SETS[cfg.velocity_set]
    p.relaxation_time = capi.RELAXATION_TIMES[cfg.relaxation_time]
    p.float_type = capi.FLOAT_TYPES[cfg.float_type]
    p.ext = (1 * cfg.ext_equilibrium_boudaries | 2 * cfg.ext_volume_force | 4 * cfg.ext_force_field | 8 * cfg.ext_magneto_hydro
             | 16 * cfg.ext_subgrid_ecr | 32 * cfg.graphics_active)
    f32 = np.float32
    p.w = float(f32(1.0) / f32(f32(3.0) * f32(cfg.nu) + f32(0.5)))
    u = cfg.units
    p.ke = float(u.ke_lu()); p.kmu0 = float(u.mu_0_lu()); p.kmu = float(f32(u.mu_0_lu() / f32(f32(4.0) * f32(np.pi))))
    p.kkge = float(u.kkge_lu()); p.kimg = float(u.kimg_lu()); p.kvev = float(u.kveV_lu()); p.kme = float(u.kme_lu())
    p.wq = float(f32(f32(1.0) / f32(f32(f32(2.0) * u.k_charge_expansion_lu()) + f32(0.5))))
    p.kkbme = float(u.kkBme_lu()); p.keabs = float(u.keabs_lu())
    p.lod_depth, p.n_lod, p.n_lod_own = cfg.mhd_lod_depth, g.n_lod, g.n_lod_own
    return p

def cmp(name, a, b, exact=True):
    a = np.asarray(a); b = np.asarray(b)
    if exact:
        same = np.array_equal(a, b, equal_nan=True) if a.dtype.kind == 'f' else np.array_equal(a, b)
        nd = int((a != b).sum()) if not same else 0
        md = float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if not same else 0.0
        print(f"   {name:8s} exact={same} ndiff={nd}/{a.size} maxabs={md:.3e}")
        return same
    den = np.linalg.norm(b.astype(np.float64)) + 1e-300
    r = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / den
    print(f"   {name:8s} rel-L2={r:.3e}")
    return r

def run(cfg, steps, seed=1):
    print("CONFIG", cfg.velocity_set, cfg.float_type, cfg.relaxation_time, "n", cfg.n_x, cfg.n_y, cfg.n_z,
          "ext", cfg.ext_equilibrium_boudaries, cfg.ext_volume_force, cfg.ext_force_field, cfg.ext_magneto_hydro, "lod", cfg.mhd_lod_depth, flush=True)
    ref = rh.RefLbm(cfg, threads=8)
    if build_only:
        return
    from ionsolver_b200 import capi
    d = ref.domains[0]
    g = d.g
    rng = np.random.default_rng(seed)
    n = g.n
    d.rho[:] = (1.0 + 0.05 * rng.standard_normal(n)).astype(np.float32)
    d.u[:] = (0.05 * rng.standard_normal(3 * n)).astype(np.float32)
    fl = np.zeros(n, np.uint8)
    r = rng.random(n)
    fl[r < 0.05] = 0x01
    if cfg.ext_equilibrium_boudaries:
        fl[(r >= 0.05) & (r < 0.08)] = 0x02
    fl[(r >= 0.08) & (r < 0.09)] = 0x11  # magnet-flagged cells are fluid to stream_collide (Q1)
    d.flags[:] = fl
    if cfg.ext_force_field:
        d.f[:] = (1e-4 * rng.standard_normal(3 * n)).astype(np.float32)
    if cfg.ext_magneto_hydro:
        d.qc[:] = (0.002 + 0.0005 * rng.standard_normal(n)).astype(np.float32)
        d.b_stat[:] = (1e-3 * rng.standard_normal(3 * n)).astype(np.float32)
        d.e_stat[:] = (1e-3 * rng.standard_normal(3 * n)).astype(np.float32)
    dom = capi.Domain(params_from(cfg, g))
    dom.write(capi.FIELD_RHO, d.rho); dom.write(capi.FIELD_U, d.u); dom.write(capi.FIELD_FLAGS, d.flags)
    if cfg.ext_force_field: dom.write(capi.FIELD_F, d.f)
    if cfg.ext_magneto_hydro:
        dom.write(capi.FIELD_Q, d.qc); dom.write(capi.FIELD_B_STAT, d.b_stat); dom.write(capi.FIELD_E_STAT, d.e_stat)
    # --- initialize
    ref.initialize()
    dom.enqueue_initialize()
    mhd = cfg.ext_magneto_hydro
    if mhd:
        dom.enqueue_update_e_b_dyn()
    ok = cmp("fi@init", dom.read(capi.FIELD_FI), d.fi)
    if mhd:
        cmp("ei@init", dom.read(capi.FIELD_EI), d.ei); cmp("fqi@init", dom.read(capi.FIELD_FQI), d.fqi)
        cmp("E@init", dom.read(capi.FIELD_E_DYN), d.e_dyn); cmp("B@init", dom.read(capi.FIELD_B_DYN), d.b_dyn)
    # --- kernel-level check for MHD: one stream_collide from identical state, then update_e_b from identical LODs
    fx, fy, fz = cfg.f_x, cfg.f_y, cfg.f_z
    for t in range(steps):
        if mhd:
            dom.enqueue_clear_qu_lod()
        dom.enqueue_stream_collide(t, fx, fy, fz)
        if mhd:
            dom.enqueue_update_e_b_dyn()
        ref.do_time_step()
        if mhd and t == 0:
            print("  after step 1:")
            cmp("fi", dom.read(capi.FIELD_FI), d.fi); cmp("ei", dom.read(capi.FIELD_EI), d.ei); cmp("fqi", dom.read(capi.FIELD_FQI), d.fqi)
            cmp("Q", dom.read(capi.FIELD_Q), d.qc)
            cmp("lod", dom.read(capi.FIELD_QU_LOD), d.qu_lod, exact=False)
            cmp("E", dom.read(capi.FIELD_E_DYN), d.e_dyn, exact=False); cmp("B", dom.read(capi.FIELD_B_DYN), d.b_dyn, exact=False)
    dom.finish()
    print(f"  after {steps} steps:")
    ex = not mhd
    cmp("fi", dom.read(capi.FIELD_FI), d.fi, exact=ex)
    if cfg.graphics_active:
        cmp("rho", dom.read(capi.FIELD_RHO), d.rho, exact=ex); cmp("u", dom.read(capi.FIELD_U), d.u, exact=ex)
    if mhd:
        cmp("Q", dom.read(capi.FIELD_Q), d.qc, exact=False)
        cmp("E", dom.read(capi.FIELD_E_DYN), d.e_dyn, exact=False); cmp("B", dom.read(capi.FIELD_B_DYN), d.b_dyn, exact=False)
    # update_fields kernel
    dom.enqueue_update_fields(steps); d.t = steps; d.enqueue_update_fields()
    cmp("rho(uf)", dom.read(capi.FIELD_RHO), d.rho, exact=ex); cmp("u(uf)", dom.read(capi.FIELD_U), d.u, exact=ex)
    dom.close()

C = rh.RefConfig
cases = [
    C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=24, n_z=20, nu=0.1, graphics_active=True),
    C(velocity_set="D3Q19", float_type="FP32", relaxation_time="TRT", n_x=33, n_y=17, n_z=9, nu=0.02, ext_equilibrium_boudaries=True,
      ext_volume_force=True, ext_force_field=True, f_x=1e-4, f_y=-2e-4, f_z=3e-4, graphics_active=True),
    C(velocity_set="D3Q19", float_type="FP16S", n_x=32, n_y=16, n_z=16, nu=0.05, ext_volume_force=True, f_x=1e-4, graphics_active=True),
    C(velocity_set="D3Q19", float_type="FP16C", relaxation_time="TRT", n_x=32, n_y=16, n_z=16, nu=0.05, graphics_active=True),
    C(velocity_set="D3Q27", float_type="FP32", n_x=20, n_y=16, n_z=12, nu=0.05, ext_equilibrium_boudaries=True, graphics_active=True),
    C(velocity_set="D3Q15", float_type="FP16S", relaxation_time="TRT", n_x=20, n_y=16, n_z=12, nu=0.05, ext_volume_force=True, f_z=1e-4),
    C(velocity_set="D2Q9", float_type="FP32", n_x=40, n_y=30, n_z=1, nu=0.05, ext_volume_force=True, f_x=1e-4, graphics_active=True),
    C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=32, n_z=32, nu=0.05, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3, graphics_active=True),
    C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=32, n_z=32, nu=0.05, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=2),
    C(velocity_set="D3Q27", float_type="FP16C", n_x=16, n_y=16, n_z=16, nu=0.05, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=1, graphics_active=True),
]
for cfg in cases:
    if cfg.ext_magneto_hydro:
        cfg.units.set(32.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    run(cfg, 5)
