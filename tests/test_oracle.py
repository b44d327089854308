"""Pins for the oracle itself (CPU only).

1. oracle/lbm_oracle.c (the plain-C restatement) reproduces, bit for bit, the golden vectors that
   tests/golden/make_golden.py recorded from the reference's OWN kernel source compiled for the host.
2. Where /root/reference exists (this container, not the GPU box) the restatement is also compared live against that
   reference build, including the voxeliser on the reference's STL assets.
"""
import os

import numpy as np
import pytest

import cases
from conftest import have_reference
from oracle import port, ref_host as rh
from oracle_util import check_against_golden, same_bits, sha, buffer_names

ALL = cases.all_cases()


@pytest.mark.parametrize("name,cfg", ALL, ids=[c[0] for c in ALL])
def test_port_matches_reference_golden(name, cfg, golden):
    g = golden["cases"][name]
    lbm = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(lbm, cfg)
    for d, gi in zip(lbm.domains, g["inputs"]):  # the seeded inputs themselves are pinned
        for n in ("rho", "u", "flags"):
            assert sha(getattr(d, n)) == gi[n], f"input {n} differs from the generator's: numpy RNG drift?"
    lbm.initialize()
    bad = check_against_golden(lbm, cfg, g["after_initialize"], "after initialize")
    if cfg.ext_magneto_hydro:
        cases.seed_electron_gas(lbm)
    for _ in range(g["steps"]):
        lbm.do_time_step()
    bad += check_against_golden(lbm, cfg, g["after_steps"], f"after {g['steps']} steps")
    assert not bad, bad


@pytest.mark.skipif(not have_reference(), reason="/root/reference absent: live reference build unavailable")
@pytest.mark.parametrize("name,cfg", ALL[::3], ids=[c[0] for c in ALL[::3]])
def test_port_matches_live_reference_build(name, cfg):
    a = rh.RefLbm(cfg, threads=1)
    b = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(a, cfg, seed=5)
    cases.fill_inputs(b, cfg, seed=5)
    for lbm in (a, b):
        lbm.initialize()
        if cfg.ext_magneto_hydro:
            cases.seed_electron_gas(lbm)
        lbm.run(3)
    for da, db in zip(a.domains, b.domains):
        for n in buffer_names(cfg):
            assert same_bits(getattr(da, n), getattr(db, n)), f"domain {da.g.d_i} buffer {n}"


EXTRA = cases.extra_oracle_cases()


@pytest.mark.skipif(not have_reference(), reason="/root/reference absent: live reference build unavailable")
@pytest.mark.parametrize("name,cfg", EXTRA, ids=[c[0] for c in EXTRA])
def test_port_matches_live_reference_build_on_extra_combinations(name, cfg):
    """MHD with TRT, D3Q15 / D2Q9 with the compressed codecs, D3Q27 TRT with a force field, a y-split: the restatement against
    the reference's kernels compiled on the spot, every buffer bit for bit after initialize + 4 steps."""
    a = rh.RefLbm(cfg, threads=1)
    b = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(a, cfg, seed=9)
    cases.fill_inputs(b, cfg, seed=9)
    for lbm in (a, b):
        lbm.initialize()
        if cfg.ext_magneto_hydro:
            cases.seed_electron_gas(lbm)
        lbm.run(4)
    for da, db in zip(a.domains, b.domains):
        for n in buffer_names(cfg):
            assert same_bits(getattr(da, n), getattr(db, n)), f"domain {da.g.d_i} buffer {n}"


def test_codecs_against_golden(golden):
    g = golden["codecs"]
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.uniform(-2.0, 2.0, 200000), rng.standard_normal(100000) * 1e-3, rng.standard_normal(100000) * 1e-6,
        np.array([0.0, -0.0, 1.0, -1.0, 1.99951168, 2.0, 6.10351562e-5, 2.98023224e-8, 1e-9, 65504.0 / 32768.0, 3.0])]).astype(np.float32)
    assert sha(x) == g["floats_sha256"]
    for ft in ("FP16S", "FP16C"):
        d = port.PortDomain(rh.RefConfig(velocity_set="D3Q19", float_type=ft, n_x=4, n_y=4, n_z=4), 0, 0, 0, 0)
        codes = np.arange(65536, dtype=np.uint16)
        if ft == "FP16S":
            codes = codes[np.isfinite(codes.view(np.float16))]
        assert sha(d.codec(codes, 1)) == g[ft]["decode_all_codes"]["sha256"]
        assert sha(d.codec(x, 0)) == g[ft]["encode_floats"]["sha256"]


def test_fp16c_round_trip_properties():
    """Format facts stated at sim_kernels.cl:79: range +-1.99951168, smallest denormal 2.98023224e-8; decode(encode(x)) is
    idempotent and monotone."""
    d = port.PortDomain(rh.RefConfig(velocity_set="D3Q19", float_type="FP16C", n_x=4, n_y=4, n_z=4), 0, 0, 0, 0)
    codes = np.arange(65536, dtype=np.uint16)
    vals = d.codec(codes, 1)
    assert np.isfinite(vals).all()
    assert np.float32(vals.max()) == np.float32(1.99951168) and np.float32(vals.min()) == np.float32(-1.99951168)
    pos = vals[:32768]
    assert (np.diff(pos) > 0).all()
    assert np.float32(pos[1]) == np.float32(2.98023224e-8)
    assert np.array_equal(d.codec(vals, 0), codes | ((vals == 0) & (codes == 0x8000)) * 0)  # every code survives a round trip


def test_neighbors_against_golden(golden):
    for vs, rec in golden["neighbors"].items():
        nx, ny, nz = rec["dims"]
        d = port.PortDomain(rh.RefConfig(velocity_set=vs, float_type="FP32", n_x=nx, n_y=ny, n_z=nz), 0, 0, 0, 0)
        tab = np.stack([d.neighbors(n) for n in range(d.g.n)]).astype(np.uint32)
        assert sha(tab) == rec["sha256"]
        stored = np.load(os.path.join(os.path.dirname(__file__), "golden", f"neighbors_{vs}.npy"))
        assert np.array_equal(tab, stored)
        # structural property: direction i+1 is the inverse of direction i (sim_kernels.cl:326-347)
        for i in range(1, tab.shape[1], 2):
            assert np.array_equal(tab[tab[:, i], i + 1], np.arange(d.g.n))


def voxel_scene(backend):
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=40, n_z=44, nu=0.05, ext_volume_force=True,
                       ext_magneto_hydro=True, mhd_lod_depth=2)
    cfg.units.set(48.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    return cfg, rh.RefLbm(cfg, threads=0, backend=backend)


def test_voxelizer_and_static_fields_against_golden(golden):
    g = golden["voxelize"]
    cfg, lbm = voxel_scene("port")
    kinds = {"Magnet": "Magnet", "Solid": "Solid", "Charged": "Charged", "ChargedECR": "ChargedECR"}
    for i, (f, kind, val, origin) in enumerate(g["config"]["meshes"]):
        data = open(os.path.join(cases.STL_DIR, f), "rb").read()
        lbm.import_mesh(data, 1.0, origin[0], origin[1], origin[2], 0.0, 0.0, 0.0)
        lbm.voxelise_mesh(i, kinds[kind], tuple(val) if isinstance(val, list) else val)
        assert sha(lbm.domains[0].flags) == g[f]["flags_after"], f"flags after voxelising {f}"
    lbm.precompute_B()
    lbm.precompute_E()
    d = lbm.domains[0]
    assert sha(d.e_dyn[: (cfg.n_x + 2) * (cfg.n_y + 2) * (cfg.n_z + 2)]) == g["psi"]["sha256"]
    assert sha(d.b_stat) == g["b_stat"]["sha256"]
    assert sha(d.e_stat) == g["e_stat"]["sha256"]


@pytest.mark.skipif(not have_reference(), reason="/root/reference absent: its STL assets are not copied into this repository")
def test_voxelizer_on_reference_stl_assets():
    """Flags after voxelising the reference's own thruster STLs (scene of setup_deeva_test, setup.rs:395-446, at half
    resolution): restatement vs reference build, bit-exact."""
    stl = "/root/reference/stl"
    res = {}
    for backend in ("ref", "port"):
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=64, n_y=128, n_z=64, nu=0.05, ext_volume_force=True,
                           ext_magneto_hydro=True, mhd_lod_depth=2)
        cfg.units.set(64.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 10e-8, 1.0, 50000.0)
        lbm = rh.RefLbm(cfg, threads=0, backend=backend)
        parts = [("deeva_disk_magnet.stl", (32.001, 0.0, 32.0), "Magnet", (0.0, 1000000.0, 0.0)),
                 ("deeva_inlet.stl", (32.0, 0.0, 32.0), "Solid", None),
                 ("deeva_quartz_tube.stl", (32.001, 0.0, 32.0), "Solid", None),
                 ("deeva_ring_magnet.stl", (32.001, -0.5, 32.0), "Magnet", (0.0, 500000.0, 0.0)),
                 ("deeva_e_plate1.stl", (32.0, 0.0, 32.0), "ChargedECR", 1.0e-13),
                 ("deeva_e_plate2.stl", (32.0, 0.0, 32.0), "ChargedECR", -1.0e-13),
                 ("disk-magnet.stl", (32.1, 100.1, 32.0), "Magnet", (0.0, 1000000.0, 0.0)),
                 ("ring-magnet.stl", (32.1, 115.1, 32.0), "Charged", 2.0e-13)]
        hist = []
        for i, (f, o, kind, val) in enumerate(parts):
            lbm.import_mesh(open(os.path.join(stl, f), "rb").read(), 1.0, o[0], o[1], o[2], 0.0, 0.0, 0.0)
            lbm.voxelise_mesh(i, kind, val)
            hist.append(lbm.domains[0].flags.copy())
        res[backend] = (hist, lbm.domains[0].b_dyn.copy())
    for i, (x, y) in enumerate(zip(res["ref"][0], res["port"][0])):
        assert np.array_equal(x, y), f"flags differ after mesh {i}"
        assert (x != 0).sum() > 0
    assert same_bits(res["ref"][1], res["port"][1])


def test_parity_scenes_never_store_nan():
    """Every MHD parity scene stays finite: no NaN is ever handed to a DDF encoder over the recorded steps (oracle hook
    ora_nan_stores).  The reference launders NaN through the FP16C bit formula into a finite code that depends on the NaN's sign
    and payload -- x86's default NaN 0xFFC00000 becomes -1.5, NVIDIA's 0x7FFFFFFF becomes -0 -- so a scene that stores NaN has no
    hardware-independent reference result."""
    from oracle import port
    ragged = {n for n, _ in cases.ragged_mhd_cases()}  # compared after ONE step (tests/test_gpu_parity.py::test_multi_domain_mhd)
    for name, cfg in cases.all_cases() + cases.ragged_mhd_cases():
        if not cfg.ext_magneto_hydro or cfg.n_x * cfg.n_y * cfg.n_z > 40000:
            continue
        lbm = rh.RefLbm(cfg, threads=1, backend="port")
        cases.fill_inputs(lbm, cfg)
        lbm.initialize()
        cases.seed_electron_gas(lbm)
        port.nan_stores(reset=True)
        for _ in range(1 if name in ragged else 8):
            lbm.do_time_step()
        assert port.nan_stores(reset=True) == 0, name


def test_mass_and_charge_are_conserved_by_the_oracle():
    """Size-independent property the GPU tests reuse at full size: a periodic box without solids conserves sum(rho)."""
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=24, n_y=20, n_z=16, nu=0.05, graphics_active=True)
    lbm = rh.RefLbm(cfg, threads=0, backend="port")
    rng = np.random.default_rng(3)
    d = lbm.domains[0]
    d.rho[:] = (1.0 + 0.02 * rng.standard_normal(d.g.n)).astype(np.float32)
    d.u[:] = (0.02 * rng.standard_normal(3 * d.g.n)).astype(np.float32)
    m0 = d.rho.astype(np.float64).sum()
    lbm.run(20)
    assert abs(d.rho.astype(np.float64).sum() - m0) / m0 < 1e-6
