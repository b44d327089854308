"""Shared test matrix: the configurations every parity test, the golden-vector generator and build() iterate over.

A case is an `oracle.ref_host.RefConfig` (the oracle's statement of LbmConfig).  `to_lbm_config` turns it into the
product's `ionsolver_b200.lbm.LbmConfig`; `fill_inputs` produces the seeded synthetic state both sides start from.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_host as rh  # noqa: E402

C = rh.RefConfig
STL_DIR = os.path.join(ROOT, "tests", "golden", "stl")


def _mhd(c, length=32.0, weak=False):
    """setup_bfield_spin-style units (setup.rs:144); `weak` scales the charge unit so that the Coulomb coupling stays
    small and the dynamics are well conditioned for multi-step comparisons."""
    c.units.set(length, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-7 if weak else 1e-10, 1.0)
    return c


def single_domain_cases():
    """(name, RefConfig): every (velocity set x collision x storage x extension) family the reference can build,
    on odd sizes that exercise the periodic wrap."""
    return [
        ("d3q19_fp32_srt", C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=24, n_z=20, nu=0.1, graphics_active=True)),
        ("d3q19_fp32_trt_eqb_ff", C(velocity_set="D3Q19", float_type="FP32", relaxation_time="TRT", n_x=33, n_y=17, n_z=9, nu=0.02,
                                    ext_equilibrium_boudaries=True, ext_volume_force=True, ext_force_field=True, f_x=1e-4, f_y=-2e-4,
                                    f_z=3e-4, graphics_active=True)),
        ("d3q19_fp16s_srt_vf", C(velocity_set="D3Q19", float_type="FP16S", n_x=32, n_y=16, n_z=16, nu=0.05, ext_volume_force=True,
                                 f_x=1e-4, graphics_active=True)),
        ("d3q19_fp16c_trt", C(velocity_set="D3Q19", float_type="FP16C", relaxation_time="TRT", n_x=32, n_y=16, n_z=16, nu=0.05,
                              graphics_active=True)),
        ("d3q27_fp32_srt_eqb", C(velocity_set="D3Q27", float_type="FP32", n_x=20, n_y=16, n_z=12, nu=0.05,
                                 ext_equilibrium_boudaries=True, graphics_active=True)),
        ("d3q15_fp16s_trt_vf", C(velocity_set="D3Q15", float_type="FP16S", relaxation_time="TRT", n_x=20, n_y=16, n_z=12, nu=0.05,
                                 ext_volume_force=True, f_z=1e-4)),
        ("d2q9_fp32_srt_vf", C(velocity_set="D2Q9", float_type="FP32", n_x=40, n_y=30, n_z=1, nu=0.05, ext_volume_force=True,
                               f_x=1e-4, graphics_active=True)),
        ("d3q27_fp16c_trt_vf_big", C(velocity_set="D3Q27", float_type="FP16C", relaxation_time="TRT", n_x=300, n_y=5, n_z=3, nu=0.03,
                                     ext_volume_force=True, f_y=1e-4, graphics_active=True)),
    ]


def mhd_cases():
    return [
        ("mhd_d3q19_fp32_lod3", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=32, n_z=32, nu=0.05, ext_volume_force=True,
                                      ext_magneto_hydro=True, mhd_lod_depth=3, graphics_active=True))),
        ("mhd_d3q19_fp32_lod4", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=16, n_z=48, nu=0.05, ext_volume_force=True,
                                      ext_magneto_hydro=True, mhd_lod_depth=4, graphics_active=True))),
        ("mhd_d3q19_fp32_lod2", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=32, n_z=32, nu=0.05, ext_volume_force=True,
                                      ext_magneto_hydro=True, mhd_lod_depth=2))),
        ("mhd_d3q27_fp16c_lod1", _mhd(C(velocity_set="D3Q27", float_type="FP16C", n_x=16, n_y=16, n_z=16, nu=0.05, ext_volume_force=True,
                                       ext_magneto_hydro=True, mhd_lod_depth=1, graphics_active=True), 16.0)),
        ("mhd_d3q19_fp16s_lod2_eqb", _mhd(C(velocity_set="D3Q19", float_type="FP16S", n_x=24, n_y=16, n_z=20, nu=0.05, ext_volume_force=True,
                                           ext_equilibrium_boudaries=True, ext_magneto_hydro=True, mhd_lod_depth=2, graphics_active=True), 24.0)),
    ]


def multi_domain_cases():
    """Split lattices; the halo layout, flags in halos and (without MHD) every field have to stay bit-exact."""
    return [
        ("z2_d3q19_fp32", C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=12, n_z=20, d_z=2, nu=0.05, graphics_active=True)),
        ("z3_d3q19_fp16s_trt", C(velocity_set="D3Q19", float_type="FP16S", relaxation_time="TRT", n_x=16, n_y=8, n_z=18, d_z=3, nu=0.05,
                                 ext_volume_force=True, f_x=1e-4)),
        ("x2y2_d3q27_fp32", C(velocity_set="D3Q27", float_type="FP32", n_x=12, n_y=16, n_z=6, d_x=2, d_y=2, nu=0.05, graphics_active=True)),
        ("x2y2z2_d3q15_fp16c", C(velocity_set="D3Q15", float_type="FP16C", n_x=12, n_y=8, n_z=8, d_x=2, d_y=2, d_z=2, nu=0.05)),
        ("x2_d2q9_fp32", C(velocity_set="D2Q9", float_type="FP32", n_x=24, n_y=18, n_z=1, d_x=2, nu=0.05, graphics_active=True)),
    ]


def multi_domain_mhd_cases():
    """Halo-inclusive local sizes are multiples of 2^depth: otherwise the reference's LOD deposit writes outside
    QU_lod (quirk Q7, undefined behaviour) and there is nothing to compare against."""
    return [
        ("mhd_z2_d3q19_fp32_lod2", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=28, d_z=2, nu=0.05,
                                         ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=2, graphics_active=True))),
        ("mhd_z4_d3q19_fp16s_lod1", _mhd(C(velocity_set="D3Q19", float_type="FP16S", n_x=8, n_y=8, n_z=24, d_z=4, nu=0.05,
                                          ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=1), 8.0)),
        # depth 3 / 4: the tiled and packed update_e_b_dynamic kernels with foreign-domain pyramids (level depth-1 of the
        # neighbour slab: 64 / 512 sources); local size incl. halos 16^3
        ("mhd_z2_d3q19_fp32_lod3", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=28, d_z=2, nu=0.05,
                                         ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3), 16.0)),
        ("mhd_z2_d3q19_fp32_lod4", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=28, d_z=2, nu=0.05,
                                         ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4, graphics_active=True), 16.0)),
        ("mhd_z3_d3q19_fp32_lod4", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=16, n_z=42, d_z=3, nu=0.05,
                                         ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 32.0)),
        # the template combinations of BASELINE.json's scaling configurations on block-aligned small lattices:
        # cfg4 = D3Q27 x FP16S x MHD x z split x depth 4 (D3Q27 with canonical weights, quirk Q3), cfg5 = D3Q19 x FP16C x MHD x z split x depth 4
        ("mhd_z2_d3q27_fp16s_lod4", _mhd(C(velocity_set="D3Q27", float_type="FP16S", n_x=16, n_y=16, n_z=28, d_z=2, nu=0.05,
                                          ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 16.0)),
        # (weak coupling: with the strong-coupling units one cell reaches rho_e = 0 at step 5 and stores NaN DDFs, which the
        # FP16C bit formula launders into a finite code that depends on the NaN's sign and payload, i.e. on the hardware --
        # tests/test_oracle.py::test_parity_scenes_never_store_nan keeps every scene inside the domain where parity is defined)
        ("mhd_z2_d3q19_fp16c_lod4", _mhd(C(velocity_set="D3Q19", float_type="FP16C", n_x=16, n_y=16, n_z=28, d_z=2, nu=0.05,
                                          ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4, graphics_active=True), 16.0, weak=True)),
    ]


def ragged_mhd_cases():
    """z slabs whose halo-inclusive height is NOT a multiple of 2^depth -- what every benchmark slab looks like (256 + 2 layers):
    lod_index overflows for the top layers (quirk Q7; the oracle drops deposits past the end of QU_lod), and update_e_b_dynamic
    has cells whose z block index is 2^depth (an extra z window of the polyphase FFT path).  Default mode only: the
    deterministic mode needs block-aligned slabs."""
    return [
        ("mhd_z2_d3q19_fp32_lod3_ragged", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=32, d_z=2, nu=0.05,
                                                ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3), 16.0)),
        ("mhd_z2_d3q19_fp32_lod4_ragged", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=32, n_y=16, n_z=64, d_z=2, nu=0.05,
                                                ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 32.0)),
        ("mhd_z3_d3q19_fp16s_lod4_ragged", _mhd(C(velocity_set="D3Q19", float_type="FP16S", n_x=16, n_y=32, n_z=48, d_z=3, nu=0.05,
                                                 ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 32.0, weak=True)),
        # tall slabs (256 layers incl. halos): the third slab sees the first one 289+ cells away, far enough for the Taylor path of
        # the far slabs (eb_fft.cu far_set: R >= 40 x 6.06); its neighbour's pyramid goes through the FFT as a second source set
        ("mhd_z3_d3q19_fp32_lod4_tall", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=762, d_z=3, nu=0.05,
                                              ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 16.0)),
        # 128-layer slabs: the first slab is 145+ cells from the third, which selects the 4^3 Taylor blocks (R >= 25 x 2.6)
        ("mhd_z3_d3q19_fp32_lod4_tall128", _mhd(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=16, n_z=378, d_z=3, nu=0.05,
                                                 ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4), 16.0)),
    ]


def drift_scene(float_type="FP32", depth=4, n=(32, 16, 16), velocity_set="D3Q19"):
    """A WELL-POSED MHD scene for N-step comparisons (returns cfg; fill with `fill_drift_inputs`).

    With every unit set the reference's scenes use, the electron gas is driven bang-bang: its acceleration per step is
    E * 0.5 / KKGE with KKGE ~ -1e-14 LU (units.rs:175-177), so any field above ~1e-15 LU saturates u_e at the +-c_s clamp with
    the SIGN of the force (sim_kernels.cl:643-646) and a last-bit difference in E flips cells -- no tolerance survives a few
    steps, in the reference itself (its LOD sums are float atomics).  Here the velocity unit is 1e8 m/s per lattice unit, which
    scales KE down to ~8e-17, and there are no static fields: the self-consistent E accelerates the electrons by ~3e-4 c per
    step -- coupled, smooth, finite for hundreds of steps in FP32, FP16S and FP16C (checked on the oracle)."""
    c = C(velocity_set=velocity_set, float_type=float_type, n_x=n[0], n_y=n[1], n_z=n[2], nu=0.05, ext_volume_force=True,
          ext_magneto_hydro=True, mhd_lod_depth=depth, graphics_active=True)
    c.units.set(float(n[0]), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0e8, 1.2250, 1e-10, 1.0)
    return c


def fill_drift_inputs(lbm, cfg, seed=21):
    """Smooth seeded rho / u, no solids, no static fields, a net charge of 0.002 per cell modulated along x and y."""
    fill_inputs(lbm, cfg, seed=seed, smooth=True)
    for d in lbm.domains:
        n = d.g.n
        i = np.arange(n)
        x, y = i % d.g.n_x, (i // d.g.n_x) % d.g.n_y
        d.flags[:] = 0
        d.qc[:] = (RHO_E0 + 0.002 * (1.0 + 0.3 * np.sin(2 * np.pi * x / d.g.n_x) * np.cos(2 * np.pi * y / d.g.n_y))).astype(np.float32)
        d.e_stat[:] = 0
        d.b_stat[:] = 0


def taylor_green_numpy(n):
    """setup.rs:458-543 in numpy float32 (one domain)."""
    f32 = np.float32
    pif, A = f32(np.pi), f32(0.25)
    g = np.arange(n, dtype=np.float32)
    f = (g + f32(0.5) - f32(0.5) * f32(n)).astype(np.float32)
    a = f32(n)
    arg2 = (f32(2.0) * pif * f / a).astype(np.float32)
    arg4 = (f32(4.0) * pif * f / a).astype(np.float32)
    c2, s2, c4 = np.cos(arg2), np.sin(arg2), np.cos(arg4)
    Z, Y, X = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ux = A * c2[X] * s2[Y] * s2[Z]
    uy = -A * s2[X] * c2[Y] * s2[Z]
    uz = A * s2[X] * s2[Y] * c2[Z]
    rho = f32(1.0) - (A * A) * f32(3.0) / f32(4.0) * c4[X] + c4[Y]
    return np.concatenate([ux.ravel(), uy.ravel(), uz.ravel()]).astype(np.float32), rho.ravel().astype(np.float32)


def _ecr(c, length=16.0):
    """Units for the SUBGRID_ECR cases.  The extension is experimental in the reference (sim.cl:556-629: debug printfs, a
    placeholder ionisation term, `a - b / 2.0f` in grad_mag_v) and with setup_deeva_test's units DEF_KKBME is ~ -1e12, so the
    electron drift term overflows within two steps.  A velocity unit of 1e6 m/s per lattice unit brings DEF_KKBME to -2.3e-5
    and keeps every field finite over the 8 recorded steps; the arithmetic exercised is the same."""
    c.units.set(length, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0e6, 1.2250, 1e-10, 1.0)
    return c


def ecr_cases():
    """SUBGRID_ECR: single domain only -- the reference never communicates `eti` (types.rs:105-111)."""
    return [
        ("ecr_d3q19_fp32", _ecr(C(velocity_set="D3Q19", float_type="FP32", n_x=16, n_y=12, n_z=8, nu=0.05, ext_volume_force=True,
                                  ext_magneto_hydro=True, ext_subgrid_ecr=True, mhd_lod_depth=2, graphics_active=True, ecr_freq=2.0e-4))),
        ("ecr_d3q27_fp16c_trt", _ecr(C(velocity_set="D3Q27", float_type="FP16C", relaxation_time="TRT", n_x=16, n_y=16, n_z=16, nu=0.05,
                                       ext_volume_force=True, ext_magneto_hydro=True, ext_subgrid_ecr=True, mhd_lod_depth=3,
                                       ecr_freq=2.0e-4))),
    ]


def extra_oracle_cases():
    """Oracle-only pins (CPU, live reference build): combinations of velocity set / collision / storage / extensions that the
    golden matrix above does not contain, so that the C restatement is compared with the reference's own kernels on them too."""
    return [
        ("x_d3q15_fp32_srt_eqb_ff_vf", C(velocity_set="D3Q15", float_type="FP32", n_x=19, n_y=11, n_z=7, nu=0.03, ext_equilibrium_boudaries=True,
                                         ext_volume_force=True, ext_force_field=True, f_x=2e-4, f_z=-1e-4, graphics_active=True)),
        ("x_d2q9_fp16c_trt", C(velocity_set="D2Q9", float_type="FP16C", relaxation_time="TRT", n_x=28, n_y=18, n_z=1, nu=0.04, graphics_active=True)),
        ("x_d3q27_fp32_trt_ff", C(velocity_set="D3Q27", float_type="FP32", relaxation_time="TRT", n_x=12, n_y=10, n_z=8, nu=0.06,
                                  ext_volume_force=True, ext_force_field=True, f_y=1e-4)),
        ("x_mhd_d3q19_fp32_trt_lod2", _mhd(C(velocity_set="D3Q19", float_type="FP32", relaxation_time="TRT", n_x=16, n_y=16, n_z=16, nu=0.05,
                                             ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=2, graphics_active=True), 16.0, weak=True)),
        ("x_mhd_d3q15_fp16c_lod2", _mhd(C(velocity_set="D3Q15", float_type="FP16C", n_x=16, n_y=12, n_z=8, nu=0.05, ext_volume_force=True,
                                          ext_magneto_hydro=True, mhd_lod_depth=2), 16.0)),
        ("x_y2_d3q19_fp16c_trt", C(velocity_set="D3Q19", float_type="FP16C", relaxation_time="TRT", n_x=10, n_y=16, n_z=6, d_y=2, nu=0.05,
                                   ext_volume_force=True, f_x=1e-4, graphics_active=True)),
    ]


def all_cases():
    return single_domain_cases() + mhd_cases() + multi_domain_cases() + multi_domain_mhd_cases() + ecr_cases()


def fill_inputs(lbm, cfg, seed=1, smooth=False):
    """Seeded synthetic state written into an oracle RefLbm (numpy buffers).  Halo cells are filled too: the
    initial communicate_rho_u_flags overwrites them, which is part of what is tested."""
    rng = np.random.default_rng(seed)
    amp = 0.01 if smooth else 0.05
    for d in lbm.domains:
        n = d.g.n
        d.rho[:] = (1.0 + amp * rng.standard_normal(n)).astype(np.float32)
        d.u[:] = (amp * rng.standard_normal(3 * n)).astype(np.float32)
        fl = np.zeros(n, np.uint8)
        r = rng.random(n)
        fl[r < 0.05] = 0x01
        if cfg.ext_equilibrium_boudaries:
            fl[(r >= 0.05) & (r < 0.08)] = 0x02
        fl[(r >= 0.08) & (r < 0.09)] = 0x11  # magnet-flagged cells are fluid to stream_collide (quirk Q1)
        d.flags[:] = fl
        if cfg.ext_force_field:
            d.f[:] = (1e-4 * rng.standard_normal(3 * n)).astype(np.float32)
        if cfg.ext_magneto_hydro:
            d.qc[:] = (RHO_E0 + 0.002 + 0.0005 * rng.standard_normal(n)).astype(np.float32)
            d.b_stat[:] = (1e-3 * rng.standard_normal(3 * n)).astype(np.float32)
            d.e_stat[:] = (1e-3 * rng.standard_normal(3 * n)).astype(np.float32)
        if cfg.ext_subgrid_ecr:  # a nearly uniform guide field, a weak oscillating field and a temperature around 1
            d.b_stat[:] = (1e-5 * rng.standard_normal(3 * n)).astype(np.float32)
            d.b_stat[n:2 * n] += np.float32(0.01)
            d.e_var[:] = (1e-13 * rng.standard_normal(3 * n)).astype(np.float32)
            d.et[:] = (1.0 + 0.01 * rng.standard_normal(n)).astype(np.float32)


# The reference initialises the electron gas at density 0 (initialize: calculate_f_eq(0.0, ...), quirk Q10); its first
# stream_collide then divides by rho_e = 0 in some cells and the NaN floods every field within three steps.  The MHD
# scenes therefore give the electron gas a finite density after initialize(), the way a user of the reference has to:
# by writing the `ei` buffer.  Gas charge (fill_inputs) is RHO_E0 + 0.002 so that the net charge stays small.
RHO_E0 = 0.1
_W = {"D3Q15": [2 / 9] + [1 / 9] * 6 + [1 / 72] * 8, "D3Q19": [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12,
      "D3Q27": [8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8}


def electron_gas_at_rest(ref_domain, cfg, rho_e=RHO_E0):
    """Stored `ei` DDFs of a resting electron gas of density rho_e: DDF-shifted equilibrium f_i = w_i (rho_e - 1)."""
    w = np.asarray(_W[cfg.velocity_set], np.float32)
    ei = (w[:, None] * np.float32(rho_e - 1.0) * np.ones((len(w), ref_domain.g.n), np.float32)).astype(np.float32).ravel()
    return ei if cfg.float_type == "FP32" else ref_domain.codec(ei, 0)


def seed_electron_gas(ref_lbm, gpu_lbm=None):
    """Call after initialize() on both sides."""
    cfg = ref_lbm.config
    for i, rd in enumerate(ref_lbm.domains):
        stored = electron_gas_at_rest(rd, cfg)
        rd.ei[:] = stored
        if gpu_lbm is not None:
            gpu_lbm.domains[i].write(FIELD_OF["ei"], stored)


VS = {"D2Q9": 0, "D3Q15": 1, "D3Q19": 2, "D3Q27": 3}
RT = {"SRT": 0, "TRT": 1}
FT = {"FP16S": 0, "FP16C": 1, "FP32": 2}


def to_lbm_config(cfg, deterministic=False):
    """RefConfig -> product LbmConfig (same field names as mod.rs:46-101)."""
    from ionsolver_b200 import lbm as L
    u = cfg.units
    return L.LbmConfig(
        velocity_set=VS[cfg.velocity_set], relaxation_time=RT[cfg.relaxation_time], float_type=FT[cfg.float_type],
        units=L.Units(float(u.m), float(u.kg), float(u.s), float(u.a), float(u.k)),
        n_x=cfg.n_x, n_y=cfg.n_y, n_z=cfg.n_z, d_x=cfg.d_x, d_y=cfg.d_y, d_z=cfg.d_z, nu=float(np.float32(cfg.nu)),
        f_x=cfg.f_x, f_y=cfg.f_y, f_z=cfg.f_z, ext_equilibrium_boudaries=cfg.ext_equilibrium_boudaries,
        ext_volume_force=cfg.ext_volume_force, ext_force_field=cfg.ext_force_field, ext_magneto_hydro=cfg.ext_magneto_hydro,
        ext_subgrid_ecr=cfg.ext_subgrid_ecr, mhd_lod_depth=cfg.mhd_lod_depth, ecr_freq=cfg.ecr_freq,
        graphics_config=L.GraphicsConfig(cfg.graphics_active), deterministic=deterministic)


# oracle buffer name -> product field id (include/ionsolver_b200.h enum IonField)
FIELD_OF = {"fi": 0, "rho": 1, "u": 2, "flags": 3, "f": 4, "e_stat": 5, "b_stat": 6, "e_dyn": 7, "b_dyn": 8, "fqi": 9, "ei": 10,
            "qc": 11, "qu_lod": 12, "e_var": 13, "eti": 14, "et": 15, "transfer_p": 16, "transfer_m": 17}


def upload_inputs(ref_lbm, gpu_lbm):
    """Copy the oracle's input state (rho, u, flags, F, Q, static fields) into the product's domains."""
    cfg = ref_lbm.config
    for rd, gd in zip(ref_lbm.domains, gpu_lbm.domains):
        for name in ("rho", "u", "flags"):
            gd.write(FIELD_OF[name], getattr(rd, name))
        if cfg.ext_force_field:
            gd.write(FIELD_OF["f"], rd.f)
        if cfg.ext_magneto_hydro:
            for name in ("qc", "b_stat", "e_stat") + (("e_var", "et") if cfg.ext_subgrid_ecr else ()):
                gd.write(FIELD_OF[name], getattr(rd, name))
