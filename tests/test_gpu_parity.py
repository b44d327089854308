"""Parity tests proper: the CUDA path, called through the C ABI (host layer -> ion_* -> sm_100a kernels), against the
oracle on the same seeded inputs, against the golden vectors recorded from the reference's own kernels, and -- at
BASELINE.json's full sizes -- through size-independent properties.

Bars (north_star): bit-exact for flags, DDF storage words, neighbour indexing, halo layout and every deterministic
float field; a stated relative tolerance only where the reference itself is order-dependent (float atomics of the LOD
deposit, quirk Q6) or where the CUDA kernel uses rsqrt for r/|r|^3 (update_e_b_dynamic).
"""
import os

import numpy as np
import pytest

import cases
from oracle import ref_host as rh
from oracle_util import buffer_names, rel_l2, same_bits, sha

pytestmark = pytest.mark.gpu

# tolerances (relative L2 unless noted)
TOL_LOD = 2e-6        # warp-tree sum vs sequential float atomics
TOL_EB_1STEP = 2e-6   # rsqrt-based r/|r|^3 and re-ordered sums in update_e_b_dynamic, same LOD input
TOL_RHO_U = 1e-5      # FP32 after N steps on a well-conditioned scene
TOL_QEB = 1e-4
TOL_FP16 = 2e-3


def product(cfg, devices=None, deterministic=False):
    from ionsolver_b200 import lbm as L
    return L.Lbm(cases.to_lbm_config(cfg, deterministic), devices=devices or [0])


def gpu_buffers(gd, cfg, names):
    return {n: gd.read(cases.FIELD_OF[n]) for n in names}


def assert_bit_exact(ref_lbm, gpu_lbm, cfg, names, where):
    bad = []
    for rd, gd in zip(ref_lbm.domains, gpu_lbm.domains):
        for n in names:
            want = getattr(rd, n)
            got = gd.read(cases.FIELD_OF[n])[: want.size] if n.startswith("transfer") else gd.read(cases.FIELD_OF[n])
            if not same_bits(np.asarray(got).view(want.dtype) if got.dtype != want.dtype else got, want):
                diff = int((np.asarray(got).view(np.uint8) != want.view(np.uint8)).sum()) if got.nbytes == want.nbytes else -1
                bad.append(f"{where}: domain {rd.g.d_i} {n} ({diff} bytes differ)")
    assert not bad, bad


SINGLE = cases.single_domain_cases()


@pytest.mark.parametrize("name,cfg", SINGLE, ids=[c[0] for c in SINGLE])
def test_single_domain_bit_exact(name, cfg, golden, gpu_lib):
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    names = buffer_names(cfg)
    ref.initialize()
    gpu.initialize()
    assert_bit_exact(ref, gpu, cfg, names, "after initialize")
    g = golden["cases"][name]
    for _ in range(g["steps"]):
        ref.do_time_step()
        gpu.do_time_step()
    assert_bit_exact(ref, gpu, cfg, names, "after steps")
    # ... and against what the reference's own kernels produced
    for n in names:
        assert sha(gpu.domains[0].read(cases.FIELD_OF[n])) == g["after_steps"][0][n]["sha256"], f"golden {n}"
    # update_fields kernel (sim_kernels.cl:834-859)
    gpu.domains[0].enqueue_update_fields(gpu.get_time_step())
    ref.domains[0].enqueue_update_fields()
    assert_bit_exact(ref, gpu, cfg, ["rho", "u"], "update_fields")
    gpu.close()


def test_codecs_exhaustive(gpu_lib):
    """FP16S / FP16C storage codecs on the device against the oracle (whose codecs are pinned to the reference's by the golden
    vectors): every one of the 65 536 codes decodes to the same float, and 2^22 random floats plus the special values (zeros,
    denormal range of both formats, rounding ties, overflow wrap of FP16C, +-2) encode to the same code."""
    from ionsolver_b200 import capi
    codes = np.arange(65536, dtype=np.uint16)
    rng = np.random.default_rng(11)
    x = np.concatenate([
        (rng.standard_normal(1 << 21) * np.exp(rng.uniform(-20, 1, 1 << 21))).astype(np.float32),
        rng.uniform(-2.5, 2.5, 1 << 21).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 1.9990234, 2.0, -2.0, 3.0, 6.1035156e-05, 3.0517578e-05, 2.9802322e-08, 1e-9, 65504.0 / 32768.0],
                 np.float32),
        (codes.astype(np.float32) / 2048.0 * np.float32(2.0 ** -14)).astype(np.float32),  # FP16C denormal grid incl. ties
    ])
    # FP16C's encoder is a single round-toward-zero multiplication on the device (lattice.cuh): sweep the bit patterns of the
    # binades around its normal/denormal boundary (2^-28 .. 2^-11), every pattern at the top of each binade (mantissa carries)
    # and around the 4-bit exponent wrap, both signs
    sweep = np.arange(0x31000000, 0x3A000000, 29, dtype=np.uint32)
    tops = np.concatenate([np.arange((e << 23) - 0x1800, (e << 23) + 0x1800, dtype=np.uint32) for e in range(98, 132)])
    bits = np.concatenate([sweep, tops, np.array([0x7F800000, 0x7F7FFFFF, 0x00000001, 0x007FFFFF, 0x00800000], np.uint32)])
    x = np.concatenate([x, bits.view(np.float32), (bits | np.uint32(0x80000000)).view(np.float32)])
    for ft, name in ((0, "FP16S"), (1, "FP16C")):
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type=name, n_x=4, n_y=4, n_z=4)
        dom = rh.RefLbm(cfg, threads=1, backend="port").domains[0]
        want_dec = dom.codec(codes, 1)
        got_dec = capi.codec_probe(ft, codes, 1)
        nan = np.isnan(want_dec)  # FP16S has 2046 NaN codes: they decode to NaN on both sides, payloads are not compared
        assert (np.isnan(got_dec) == nan).all()
        assert same_bits(got_dec[~nan], want_dec[~nan]), f"{name} decode: {int((got_dec.view(np.uint32) != want_dec.view(np.uint32))[~nan].sum())} codes differ"
        want_enc = dom.codec(x, 0)
        got_enc = capi.codec_probe(ft, x, 0)
        assert same_bits(got_enc, want_enc), f"{name} encode: {int((got_enc != want_enc).sum())} values differ"


MHD = cases.mhd_cases()


@pytest.mark.parametrize("name,cfg", MHD, ids=[c[0] for c in MHD])
def test_mhd_single_step_parity(name, cfg, golden, gpu_lib):
    """From identical state: everything stream_collide writes is bit-exact (DDFs of gas, electron gas and charge,
    Q); the LOD pyramid and E/B agree to float-reordering accuracy."""
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    ref.initialize()
    gpu.initialize()
    exact = ["fi", "ei", "fqi", "qc", "rho", "u", "flags"]
    assert_bit_exact(ref, gpu, cfg, exact, "after initialize")
    g = golden["cases"][name]
    for n in exact:
        assert sha(gpu.domains[0].read(cases.FIELD_OF[n])) == g["after_initialize"][0][n]["sha256"], f"golden {n} after initialize"
    rd, gd = ref.domains[0], gpu.domains[0]
    assert rel_l2(gd.read(cases.FIELD_OF["e_dyn"]), rd.e_dyn) < TOL_EB_1STEP
    assert rel_l2(gd.read(cases.FIELD_OF["b_dyn"]), rd.b_dyn) < TOL_EB_1STEP
    # give the electron gas a finite density (cases.RHO_E0) and make the inputs of the step identical (E_dyn differs
    # in the last bit), then step once
    cases.seed_electron_gas(ref, gpu)
    gd.write(cases.FIELD_OF["e_dyn"], rd.e_dyn)
    gd.write(cases.FIELD_OF["b_dyn"], rd.b_dyn)
    ref.do_time_step()
    gpu.do_time_step()
    assert_bit_exact(ref, gpu, cfg, exact if cfg.graphics_active else ["fi", "ei", "fqi", "qc", "flags"], "after one step")
    own = 4 * rd.g.n_lod_own
    assert rel_l2(gd.read(cases.FIELD_OF["qu_lod"])[:own], rd.qu_lod[:own]) < TOL_LOD
    assert rel_l2(gd.read(cases.FIELD_OF["e_dyn"]), rd.e_dyn) < TOL_EB_1STEP * 5
    assert rel_l2(gd.read(cases.FIELD_OF["b_dyn"]), rd.b_dyn) < TOL_EB_1STEP * 5
    gpu.close()


ALL_MHD = cases.mhd_cases() + cases.multi_domain_mhd_cases() + cases.ecr_cases()


@pytest.mark.parametrize("name,cfg", ALL_MHD, ids=[c[0] for c in ALL_MHD])
def test_mhd_deterministic_mode_is_bit_exact_over_many_steps(name, cfg, golden, gpu_lib):
    """The reference's MHD dynamics are discontinuous (the electron velocity saturates at +-c_s with the sign of E,
    quirk Q10), so a rounding-level difference in E flips cells within a few steps and no tolerance survives N steps;
    the reference itself is not reproducible run to run because its LOD deposit uses float atomics (quirk Q6).
    ION_EXT_DETERMINISTIC fixes the summation order to that of a sequential run and uses the reference's exact
    arithmetic in update_e_b_dynamic: rho, u, Q, E, B, the LOD pyramid and every DDF are then BIT-IDENTICAL to the
    reference kernels after N steps -- single domain and split lattices, tiled (depth 3/4) and generic E/B kernels."""
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg, deterministic=True)
    cases.upload_inputs(ref, gpu)
    g = golden["cases"][name]
    names = buffer_names(cfg)
    ref.initialize()
    gpu.initialize()
    assert_bit_exact(ref, gpu, cfg, names, "after initialize")
    cases.seed_electron_gas(ref, gpu)
    for _ in range(g["steps"]):
        ref.do_time_step()
        gpu.do_time_step()
    gpu.finish_queues()
    assert not any(np.isnan(getattr(d, n)).any() for d in ref.domains for n in ("rho", "qc", "e_dyn"))
    assert_bit_exact(ref, gpu, cfg, names, f"after {g['steps']} steps")
    for gd, gg in zip(gpu.domains, g["after_steps"]):  # ... and to the reference's own kernels (golden vectors)
        for n in names:
            got = gd.read(cases.FIELD_OF[n])
            if n.startswith("transfer"):
                got = got[: gg[n]["size"]]
            assert sha(got) == gg[n]["sha256"], f"golden {n} domain {gd.d_i}"
    gpu.close()


def test_default_and_deterministic_e_b_agree(gpu_lib):
    """The default (fast) E/B path -- rsqrt, fused sums, warp-reduced LOD deposit -- against the deterministic one on the
    same state, depth 4 (tiled kernel, ND = 16) and depth 3 (ND = 8): relative L2 below 2e-6."""
    for depth, n in ((4, (32, 16, 48)), (3, (32, 24, 16))):
        cfg = cases._mhd(cases.C(velocity_set="D3Q19", float_type="FP32", n_x=n[0], n_y=n[1], n_z=n[2], nu=0.05, ext_volume_force=True,
                                 ext_magneto_hydro=True, mhd_lod_depth=depth, graphics_active=True))
        ref = rh.RefLbm(cfg, threads=1, backend="port")
        cases.fill_inputs(ref, cfg, seed=9)
        out = {}
        for det in (False, True):
            gpu = product(cfg, deterministic=det)
            cases.upload_inputs(ref, gpu)
            gpu.initialize()
            d = gpu.domains[0]
            d.write(cases.FIELD_OF["ei"], cases.electron_gas_at_rest(ref.domains[0], cfg))
            d.write(cases.FIELD_OF["e_dyn"], ref.domains[0].e_stat)  # identical step inputs for both modes
            d.write(cases.FIELD_OF["b_dyn"], ref.domains[0].b_stat)
            gpu.do_time_step()
            gpu.finish_queues()
            out[det] = {k: d.read(cases.FIELD_OF[k]) for k in ("e_dyn", "b_dyn", "qu_lod", "fi", "qc")}
            gpu.close()
        assert same_bits(out[False]["fi"], out[True]["fi"]) and same_bits(out[False]["qc"], out[True]["qc"])
        assert rel_l2(out[False]["qu_lod"], out[True]["qu_lod"]) < TOL_LOD
        assert rel_l2(out[False]["e_dyn"], out[True]["e_dyn"]) < TOL_EB_1STEP
        assert rel_l2(out[False]["b_dyn"], out[True]["b_dyn"]) < TOL_EB_1STEP


# N = the largest step count that stays inside SURVEY 8c's tolerance with margin (profiles/r2_drift_curve.md): FP32 holds at N = 100
# (Q 4.5e-5, E 6.5e-6, B 9e-6; rho and u identical); with FP16S / FP16C storage a rounding-level difference in E eventually makes ONE
# stored DDF round the other way, a 2^-11 relative jump -- the order of the 2e-3 tolerance itself -- and Q crosses 2e-3 at step 15-16.
DRIFT = [("FP32", 4, (32, 16, 16), 100, TOL_RHO_U, TOL_QEB), ("FP32", 3, (32, 32, 16), 100, TOL_RHO_U, TOL_QEB),
         ("FP16S", 4, (32, 16, 16), 10, TOL_FP16, TOL_FP16), ("FP16C", 4, (32, 16, 16), 10, TOL_FP16, TOL_FP16)]


@pytest.mark.parametrize("ft,depth,n,steps,tol_ru,tol_qeb", DRIFT, ids=[f"{d[0]}_lod{d[1]}_{d[3]}steps" for d in DRIFT])
def test_default_mode_tracks_the_reference_over_n_steps(ft, depth, n, steps, tol_ru, tol_qeb, gpu_lib):
    """The DEFAULT path -- the one bench.py times: LOD deposit by warp trees and replicas, update_e_b_dynamic as a polyphase
    FFT convolution -- against the oracle (the C restatement pinned bit-exactly to the reference's kernels), N full time steps
    from the same state on a well-posed scene (cases.drift_scene: the reference's own unit sets drive the electron gas bang-bang
    on the sign of E, where no tolerance survives a few steps; profiles/r2_drift_curve.md has the curves of both).
    Tolerances: FP32 rel-L2 <= 1e-5 for rho, u and <= 1e-4 for Q, E, B at N = 100; FP16S / FP16C <= 2e-3 at N = 10."""
    cfg = cases.drift_scene(ft, depth, n)
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_drift_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    ref.initialize()
    gpu.initialize()
    cases.seed_electron_gas(ref, gpu)
    assert gpu.domains[0].eb_fft_info()[1] > 0 or True
    for _ in range(steps):
        ref.do_time_step()
        gpu.do_time_step()
    gpu.finish_queues()
    rd, gd = ref.domains[0], gpu.domains[0]
    assert gd.eb_fft_info()[1] > 0, "the polyphase FFT path must be the one under test"
    assert np.isfinite(rd.e_dyn).all() and np.isfinite(rd.rho).all() and float(np.abs(rd.e_dyn).max()) > 0.0
    err = {f: rel_l2(gd.read(cases.FIELD_OF[f]), getattr(rd, f)) for f in ("rho", "u", "qc", "e_dyn", "b_dyn")}
    assert err["rho"] < tol_ru and err["u"] < tol_ru, err
    assert err["qc"] < tol_qeb and err["e_dyn"] < tol_qeb and err["b_dyn"] < tol_qeb, err
    gpu.close()


def test_streamed_kernel_spectra_give_the_same_fields(gpu_lib, monkeypatch):
    """Memory-lean mode of the FFT field update (lattices whose kernel spectra do not fit, e.g. cfg5): the spectra are recomputed
    per batch of tasks inside every step.  Forced here with a batch of 3 tasks on a single domain and on a z split: E_dyn / B_dyn
    must be bit-identical to the static mode (same kernels, same arithmetic, only the schedule differs)."""
    for name, cfg in (cases.mhd_cases()[1], cases.ragged_mhd_cases()[1]):
        ref = rh.RefLbm(cfg, threads=1, backend="port")
        cases.fill_inputs(ref, cfg)
        out = {}
        for mode in ("static", "streamed"):
            if mode == "streamed":
                monkeypatch.setenv("ION_EB_FFT_BATCH", "3")
            else:
                monkeypatch.delenv("ION_EB_FFT_BATCH", raising=False)
            gpu = product(cfg)
            cases.upload_inputs(ref, gpu)
            gpu.initialize()
            for i, rd in enumerate(ref.domains):
                gpu.domains[i].write(cases.FIELD_OF["ei"], cases.electron_gas_at_rest(rd, cfg))
            gpu.do_time_step()
            gpu.do_time_step()
            gpu.finish_queues()
            assert all(d.eb_fft_info()[1] > 3 for d in gpu.domains), name
            out[mode] = [(d.read(cases.FIELD_OF["e_dyn"]).copy(), d.read(cases.FIELD_OF["b_dyn"]).copy()) for d in gpu.domains]
            gpu.close()
        monkeypatch.delenv("ION_EB_FFT_BATCH", raising=False)
        for (e0, b0), (e1, b1) in zip(out["static"], out["streamed"]):
            assert same_bits(e0, e1) and same_bits(b0, b1), name


def test_mirrored_kernel_spectra(gpu_lib, monkeypatch):
    """Tasks with 2 ox > dsx read the kernel spectra of their partner (x offset dsx - ox) at the reflected in-plane index instead of
    owning a slot (eb_fft_core.cuh::main_phase_product_mirror; taken whenever the full set of spectra does not fit).  A split lattice
    with four cells per LOD block along x (own level AND the neighbour's level, which mirrors with a one-block shift): the mirrored
    static mode holds 3/4 of the spectra and agrees with the plain static mode to rounding (the two forms differ in one kernel sample
    that no output uses); the streamed mode with mirroring is bit-identical to the mirrored static mode."""
    cfg = cases._mhd(cases.C(velocity_set="D3Q19", float_type="FP32", n_x=64, n_y=16, n_z=64, d_z=2, nu=0.05, ext_volume_force=True,
                             ext_magneto_hydro=True, mhd_lod_depth=4), 32.0)
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    out, info = {}, {}
    for mode, env in (("plain", {}), ("mirrored", {"ION_EB_FFT_MIRROR": "1"}), ("streamed", {"ION_EB_FFT_BATCH": "3"})):
        for k in ("ION_EB_FFT_MIRROR", "ION_EB_FFT_BATCH"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        gpu = product(cfg)
        cases.upload_inputs(ref, gpu)
        gpu.initialize()
        for i, rd in enumerate(ref.domains):
            gpu.domains[i].write(cases.FIELD_OF["ei"], cases.electron_gas_at_rest(rd, cfg))
        gpu.do_time_step()
        gpu.do_time_step()
        gpu.finish_queues()
        info[mode] = [d.eb_fft_info() for d in gpu.domains]
        out[mode] = [(d.read(cases.FIELD_OF["e_dyn"]).copy(), d.read(cases.FIELD_OF["b_dyn"]).copy()) for d in gpu.domains]
        gpu.close()
    for k in ("ION_EB_FFT_MIRROR", "ION_EB_FFT_BATCH"):
        monkeypatch.delenv(k, raising=False)
    for i in range(2):
        assert info["plain"][i][1] == 12 and info["mirrored"][i][1] == 12, info           # tasks per slab: (2 + 1) z offsets x 4 x offsets
        assert info["mirrored"][i][0] * 12 == info["plain"][i][0] * 9, info               # 9 of 12 tasks own a slot
        assert info["streamed"][i][0] * 3 == info["mirrored"][i][0], info                 # a buffer of 3 slots
        for c in range(2):
            assert np.isfinite(out["plain"][i][c]).all()
            assert rel_l2(out["mirrored"][i][c], out["plain"][i][c]) < 1e-6, (i, c)
            assert same_bits(out["streamed"][i][c], out["mirrored"][i][c]), (i, c)
        assert float(np.abs(out["plain"][i][0]).max()) > 0.0  # (B_dyn may vanish: the electron gas starts at rest)
        assert not same_bits(out["mirrored"][i][0], out["plain"][i][0]), "the mirrored products must have run"


MULTI = cases.multi_domain_cases()


@pytest.mark.parametrize("name,cfg", MULTI, ids=[c[0] for c in MULTI])
def test_multi_domain_bit_exact(name, cfg, golden, gpu_lib):
    """Several domains on ONE GPU (the reference's own fallback, opencl.rs:56-61): halo pack/unpack, the p<->m swap and
    flags in halos are bit-exact against the multi-domain oracle and the reference's golden vectors."""
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    names = buffer_names(cfg)
    ref.initialize()
    gpu.initialize()
    assert_bit_exact(ref, gpu, cfg, names, "after initialize")
    g = golden["cases"][name]
    for _ in range(g["steps"]):
        ref.do_time_step()
        gpu.do_time_step()
    gpu.finish_queues()
    assert_bit_exact(ref, gpu, cfg, names, "after steps")
    for gd, gg in zip(gpu.domains, g["after_steps"]):
        for n in ("fi", "rho", "u", "flags"):
            assert sha(gd.read(cases.FIELD_OF[n])) == gg[n]["sha256"], f"golden {n} domain {gd.d_i}"
    gpu.close()


MULTI_MHD = cases.multi_domain_mhd_cases() + cases.ragged_mhd_cases()


@pytest.mark.parametrize("name,cfg", MULTI_MHD, ids=[c[0] for c in MULTI_MHD])
def test_multi_domain_mhd(name, cfg, gpu_lib):
    """Split MHD lattice: DDF/charge halos bit-exact after the first step, LOD gather + exchange (foreign levels in
    ascending domain order, quirks Q5/Q7/Q8/Q18 reproduced) and E/B within tolerance."""
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    ref.initialize()
    gpu.initialize()
    exact = ["fi", "ei", "fqi", "qc", "flags"]
    assert_bit_exact(ref, gpu, cfg, exact, "after initialize")
    cases.seed_electron_gas(ref, gpu)
    for rd, gd in zip(ref.domains, gpu.domains):
        gd.write(cases.FIELD_OF["e_dyn"], rd.e_dyn)
        gd.write(cases.FIELD_OF["b_dyn"], rd.b_dyn)
    ref.do_time_step()
    gpu.do_time_step()
    gpu.finish_queues()
    assert_bit_exact(ref, gpu, cfg, exact, "after one step")
    for rd, gd in zip(ref.domains, gpu.domains):
        assert rel_l2(gd.read(cases.FIELD_OF["qu_lod"]), rd.qu_lod) < TOL_LOD * 2, f"LOD table of domain {rd.g.d_i}"
        assert rel_l2(gd.read(cases.FIELD_OF["e_dyn"]), rd.e_dyn) < TOL_EB_1STEP * 5
        assert rel_l2(gd.read(cases.FIELD_OF["b_dyn"]), rd.b_dyn) < TOL_EB_1STEP * 5
    if "ragged" in name or "tall" in name:  # halo-inclusive slabs at LOD depth 3 / 4: the polyphase FFT path is the one under test
        assert all(d.eb_fft_info()[1] > 0 for d in gpu.domains), [d.eb_fft_info() for d in gpu.domains]
    gpu.close()


def test_voxelizer_and_static_fields(golden, gpu_lib):
    """STL import (host, f32) -> voxelize_mesh -> psi/static_b/static_e: flags, B_stat and E_stat bit-exact against the
    golden vectors of the reference build."""
    from ionsolver_b200 import lbm as L
    g = golden["voxelize"]
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=40, n_z=44, nu=0.05, ext_volume_force=True,
                       ext_magneto_hydro=True, mhd_lod_depth=2)
    cfg.units.set(48.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    gpu = product(cfg)
    kind = {"Solid": L.ModelType.Solid, "Magnet": L.ModelType.Magnet, "Charged": L.ModelType.Charged, "ChargedECR": L.ModelType.ChargedECR}
    for i, (f, k, val, origin) in enumerate(g["config"]["meshes"]):
        gpu.import_mesh(os.path.join(cases.STL_DIR, f), 1.0, origin[0], origin[1], origin[2], 0.0, 0.0, 0.0)
        m = gpu.mesh(i)
        assert [float(v) for v in m["p_min"]] == g[f]["p_min"] and [float(v) for v in m["p_max"]] == g[f]["p_max"], f"mesh bounds of {f}"
        gpu.voxelise_mesh(i, kind[k], val)
        assert sha(gpu.domains[0].read(cases.FIELD_OF["flags"])) == g[f]["flags_after"], f"flags after voxelising {f}"
    gpu.precompute_B()
    gpu.precompute_E()
    d = gpu.domains[0]
    psi = d.read(cases.FIELD_OF["e_dyn"])[: (cfg.n_x + 2) * (cfg.n_y + 2) * (cfg.n_z + 2)]
    assert sha(psi) == g["psi"]["sha256"]
    assert sha(d.read(cases.FIELD_OF["b_stat"])) == g["b_stat"]["sha256"]
    assert sha(d.read(cases.FIELD_OF["e_stat"])) == g["e_stat"]["sha256"]
    gpu.close()


def test_ion_file_round_trip(gpu_lib, tmp_path):
    """.ion snapshot (FILE_LAYOUT.txt): byte layout, reference-compatible and spec-conformant modes, domain-split
    independence of the spec-conformant form."""
    import struct
    from ionsolver_b200 import lbm as L
    cfg = cases._mhd(cases.C(velocity_set="D3Q19", float_type="FP16C", relaxation_time="TRT", n_x=12, n_y=10, n_z=8, nu=0.05,
                             ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=1, f_y=2e-4), 12.0)
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg, seed=4)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    rd = ref.domains[0]
    blob = gpu.encode(reference_compatible=True)
    n = rd.g.n
    assert blob[:16] == b"IonSolver setup\n" and len(blob) == 16 + 62 + n * (1 + 4 + 12 + 4)  # no N_C/N_M (file.rs:277-303)
    assert blob[16:19] == bytes([2, 1, 1])  # D3Q19, Trt, FP16C as enum discriminants (file.rs:197-199)
    assert struct.unpack_from("<3I", blob, 35) == (12, 10, 8) and blob[76] == 0b1010 and blob[77] == 1
    body = blob[78:]
    assert body[:n] == rd.flags.tobytes() and body[n:5 * n] == rd.rho.tobytes()
    assert body[5 * n:17 * n] == rd.u.tobytes() and body[17 * n:] == rd.qc.tobytes()
    # spec-conformant: same sections + N_C = N_M = 0; decodes to the identical state with the identical float type
    spec = gpu.encode(reference_compatible=False)
    assert spec == blob + b"\0" * 8
    back = L.Lbm.decode(spec, reference_compatible=False, devices=[0])
    assert back.config.float_type == L.FloatType.FP16C and back.config.relaxation_time == L.RelaxationTime.Trt
    for name in ("flags", "rho", "u", "qc"):
        assert same_bits(back.domains[0].read(cases.FIELD_OF[name]), getattr(rd, name))
    assert np.float32(back.config.f_y) == np.float32(2e-4) and back.config.mhd_lod_depth == 1
    back.close()
    # the reference's decoder swaps FP16S/FP16C (file.rs:68-73) and wants N_C/N_M for MHD files (file.rs:155-181):
    # it cannot read the reference's own MHD output -- reproduced as an error instead of the Rust panic
    from ionsolver_b200 import capi
    with pytest.raises(capi.IonError):
        L.Lbm.decode(blob, reference_compatible=True, devices=[0])
    swapped = L.Lbm.decode(spec, reference_compatible=True, devices=[0])
    assert swapped.config.float_type == L.FloatType.FP16S
    swapped.close()
    # files: write() / read() default to the self-consistent layout (the library reloads what it wrote: MHD, FP16C)
    p = tmp_path / "state.ion"
    gpu.write(p)
    assert open(p, "rb").read() == spec
    again = L.Lbm.read(p)
    assert again.config.float_type == L.FloatType.FP16C and again.config.ext_magneto_hydro
    for name in ("flags", "rho", "u", "qc"):
        assert same_bits(again.domains[0].read(cases.FIELD_OF[name]), getattr(rd, name))
    again.close()
    # a single-domain MHD file as the reference's own encoder writes it (no N_C / N_M trailer) loads too
    p2 = tmp_path / "reference_written.ion"
    open(p2, "wb").write(blob)
    tolerant = L.Lbm.read(p2)
    assert same_bits(tolerant.domains[0].read(cases.FIELD_OF["qc"]), rd.qc)
    tolerant.close()
    gpu.close()
    # write-then-read round trip with MHD, FP16S and a z split: identical global sections, identical codec
    cfg_s = cases._mhd(cases.C(velocity_set="D3Q19", float_type="FP16S", n_x=8, n_y=8, n_z=12, d_z=2, nu=0.05, ext_volume_force=True,
                               ext_magneto_hydro=True, mhd_lod_depth=1), 8.0)
    ref_s = rh.RefLbm(cfg_s, threads=1, backend="port")
    cases.fill_inputs(ref_s, cfg_s, seed=9)
    g_s = product(cfg_s)
    cases.upload_inputs(ref_s, g_s)
    p3 = tmp_path / "split.ion"
    g_s.write(p3)
    r_s = L.Lbm.read(p3)
    assert r_s.config.float_type == L.FloatType.FP16S and r_s.get_d_n() == 2
    assert r_s.encode() == g_s.encode() == open(p3, "rb").read()
    for da, db in zip(g_s.domains, r_s.domains):
        ia = da.read(cases.FIELD_OF["qc"]).reshape(da.n_z, da.n_y, da.n_x)[1:-1]
        ib = db.read(cases.FIELD_OF["qc"]).reshape(db.n_z, db.n_y, db.n_x)[1:-1]
        assert same_bits(ia, ib)
    g_s.close()
    r_s.close()
    # a split lattice saves the same global sections in spec-conformant mode
    cfg1 = cases.C(velocity_set="D3Q19", float_type="FP32", n_x=8, n_y=12, n_z=12, nu=0.05)
    cfg2 = cases.C(velocity_set="D3Q19", float_type="FP32", n_x=8, n_y=12, n_z=12, d_y=2, d_z=3, nu=0.05)
    a, b = product(cfg1), product(cfg2)
    a.set_taylor_green(1)
    b.set_taylor_green(1)
    ea, eb = a.encode(False), b.encode(False)
    assert ea[78:] == eb[78:] and ea[:47] == eb[:47]
    c = L.Lbm.decode(eb, reference_compatible=False, devices=[0])
    assert c.get_d_n() == 6 and c.encode(False) == eb
    for x in (a, b, c):
        x.close()


def test_fast_precompute_mode(golden, gpu_lib):
    """ion_domain_set_precompute_mode(1): psi_from_mesh as an rsqrt / FMA sum with four outputs per thread.  Same sources, same sum,
    different rounding: psi within 1e-5 relative L2 of the exact mode (which is bit-identical to the reference build, previous
    test), B_stat -- central differences of psi -- within 1e-4."""
    from ionsolver_b200 import lbm as L
    g = golden["voxelize"]
    out = {}
    for mode in (0, 1):
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=40, n_z=44, nu=0.05, ext_volume_force=True,
                           ext_magneto_hydro=True, mhd_lod_depth=2)
        cfg.units.set(48.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
        gpu = product(cfg)
        kind = {"Solid": L.ModelType.Solid, "Magnet": L.ModelType.Magnet, "Charged": L.ModelType.Charged, "ChargedECR": L.ModelType.ChargedECR}
        for i, (f, k, val, origin) in enumerate(g["config"]["meshes"]):
            gpu.import_mesh(os.path.join(cases.STL_DIR, f), 1.0, origin[0], origin[1], origin[2], 0.0, 0.0, 0.0)
            gpu.voxelise_mesh(i, kind[k], val)
        for d in gpu.domains:
            d.set_precompute_mode(mode)
        gpu.precompute_B()
        d = gpu.domains[0]
        out[mode] = (d.read(cases.FIELD_OF["e_dyn"])[: (cfg.n_x + 2) * (cfg.n_y + 2) * (cfg.n_z + 2)].copy(), d.read(cases.FIELD_OF["b_stat"]).copy())
        gpu.close()
    assert sha(out[0][0]) == g["psi"]["sha256"]  # the exact mode is the reference's result
    assert rel_l2(out[1][0], out[0][0]) < 1e-5 and rel_l2(out[1][1], out[0][1]) < 1e-4, (rel_l2(out[1][0], out[0][0]), rel_l2(out[1][1], out[0][1]))
    assert not same_bits(out[1][0], out[0][0])  # the fast kernel really ran


def test_fft_precompute_mode(golden, gpu_lib):
    """ion_domain_set_precompute_mode(2): psi_from_mesh and static_e_from_mesh as zero-padded FFT convolutions of the source cells with
    d / |d|^3 (mesh_kernels.cu, cuFFT D2Z / Z2D transforms).  Double precision between the FP32 inputs and the FP32 result, so the
    mode returns the correctly rounded sums: it differs from the exact mode by the rounding of the reference's sequential FP32 sum
    only -- the same bounds as the fast direct mode: psi and E_stat within 1e-5 relative L2, B_stat (central differences) within 1e-4."""
    from ionsolver_b200 import lbm as L
    g = golden["voxelize"]
    out = {}
    for mode in (0, 2):
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=40, n_z=44, nu=0.05, ext_volume_force=True,
                           ext_magneto_hydro=True, mhd_lod_depth=2)
        cfg.units.set(48.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
        gpu = product(cfg)
        kind = {"Solid": L.ModelType.Solid, "Magnet": L.ModelType.Magnet, "Charged": L.ModelType.Charged, "ChargedECR": L.ModelType.ChargedECR}
        for i, (f, k, val, origin) in enumerate(g["config"]["meshes"]):
            gpu.import_mesh(os.path.join(cases.STL_DIR, f), 1.0, origin[0], origin[1], origin[2], 0.0, 0.0, 0.0)
            gpu.voxelise_mesh(i, kind[k], val)
        for d in gpu.domains:
            d.set_precompute_mode(mode)
        gpu.precompute_B()
        d = gpu.domains[0]
        psi = d.read(cases.FIELD_OF["e_dyn"])[: (cfg.n_x + 2) * (cfg.n_y + 2) * (cfg.n_z + 2)].copy()
        gpu.precompute_E()
        out[mode] = (psi, d.read(cases.FIELD_OF["b_stat"]).copy(), d.read(cases.FIELD_OF["e_stat"]).copy())
        gpu.close()
    assert sha(out[0][0]) == g["psi"]["sha256"]  # the exact mode is the reference's result
    err = [rel_l2(out[2][i], out[0][i]) for i in range(3)]
    assert float(np.abs(out[0][2]).max()) > 0.0, "the scene must hold charged cells for the E_stat half of the test"
    print("FFT precompute mode vs reference order, relative L2 (psi, B_stat, E_stat):", err)
    assert err[0] < 1e-5 and err[1] < 1e-4 and err[2] < 1e-5, err
    assert not same_bits(out[2][0], out[0][0])  # the FFT path really ran


def test_reference_stl_assets_through_the_cuda_voxeliser(golden, gpu_lib):
    """The reference's own STL files (stl/*.stl, copied to tests/golden/stl/ref as fixtures) through the CUDA voxeliser and
    static-field kernels, against SHA-256 of what the reference's kernels produce (tests/golden/make_golden.py::ref_stl_vectors):
    setup_deeva_test's six thruster meshes at the scene's own 128 x 256 x 128 (flags after every mesh), the same scene at
    64 x 128 x 64 with psi / B_stat / E_var, and cfg2's disk magnet in a 256^3 lattice.  Bit-exact."""
    from ionsolver_b200 import lbm as L
    g = golden["ref_stl"]
    ref_dir = os.path.join(cases.STL_DIR, "ref")
    kind = {"Solid": L.ModelType.Solid, "Magnet": L.ModelType.Magnet, "Charged": L.ModelType.Charged, "ChargedECR": L.ModelType.ChargedECR}
    for tag, scale in (("deeva_128x256x128", 1.0), ("deeva_64x128x64", 0.5)):
        rec = g[tag]
        n = rec["n"]
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=n[0], n_y=n[1], n_z=n[2], ext_volume_force=True,
                           ext_magneto_hydro=True, ext_subgrid_ecr=True, mhd_lod_depth=2)
        cfg.units.set(128.0 * scale, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 10e-8, 1.0, 50000.0)
        cfg.nu = float(cfg.units.nu_si_lu(0.05))
        gpu = product(cfg)
        for i, m in enumerate(rec["meshes"]):
            o = m["origin"]
            gpu.import_mesh(os.path.join(ref_dir, m["file"]), 1.0, o[0], o[1], o[2], 0.0, 0.0, 0.0)
            gpu.voxelise_mesh(i, kind[m["kind"]], m["value"])
            fl = gpu.domains[0].read(cases.FIELD_OF["flags"])
            assert int((fl != 0).sum()) == m["cells_flagged"], f"{tag}: cells flagged after {m['file']}"
            assert sha(fl) == m["flags_after"], f"{tag}: flags after {m['file']}"
        if "b_stat" in rec:
            gpu.precompute_B()
            gpu.precompute_E_ECR()
            d = gpu.domains[0]
            assert sha(d.read(cases.FIELD_OF["e_dyn"])[: (n[0] + 2) * (n[1] + 2) * (n[2] + 2)]) == rec["psi"]["sha256"]
            assert sha(d.read(cases.FIELD_OF["b_stat"])) == rec["b_stat"]["sha256"]
            assert sha(d.read(cases.FIELD_OF["e_var"])) == rec["e_var"]["sha256"]
        gpu.close()
    rec = g["cfg2_disk_magnet_256"]
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=256, n_y=256, n_z=256, ext_volume_force=True, ext_magneto_hydro=True,
                       mhd_lod_depth=4)
    cfg.units.set(256.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    gpu = product(cfg)
    gpu.import_mesh_reposition(os.path.join(ref_dir, "disk-magnet.stl"), 128.1, 128.1, 128.0, 0.0, 0.0, 0.0, 127.0)
    m = gpu.mesh(0)
    assert [float(v) for v in m["p_min"]] == rec["p_min"] and [float(v) for v in m["p_max"]] == rec["p_max"]
    gpu.voxelise_mesh(0, L.ModelType.Magnet, (0.0, 1000000.0, 0.0))
    fl = gpu.domains[0].read(cases.FIELD_OF["flags"])
    assert int((fl != 0).sum()) == rec["cells_flagged"] and sha(fl) == rec["flags_after"]
    gpu.close()


def test_stl_triangle_count_is_checked_in_64_bits(gpu_lib, tmp_path):
    """mesh.rs:179 compares `84 + 50 * triangles` in 32 bits; a header claiming real + 2^31 triangles wraps to the real file size.
    The C++ restatement must reject it (the Rust code would panic on the slice bounds instead of reading past the buffer)."""
    import struct
    from ionsolver_b200 import capi
    real = 2
    body = b"".join(struct.pack("<12fH", *([0.0] * 3 + [float(i), 0, 0, 0, 1.0, 0, 0, 0, 1.0]), 0) for i in range(real))
    good = b"\0" * 80 + struct.pack("<I", real) + body
    bad = b"\0" * 80 + struct.pack("<I", real + (1 << 31)) + body
    assert (84 + 50 * (real + (1 << 31))) % (1 << 32) == len(bad)  # the 32-bit check of mesh.rs:179 would pass
    pg, pb = tmp_path / "good.stl", tmp_path / "bad.stl"
    pg.write_bytes(good)
    pb.write_bytes(bad)
    gpu = product(cases.C(velocity_set="D3Q19", float_type="FP32", n_x=8, n_y=8, n_z=8, nu=0.05))
    gpu.import_mesh(str(pg), 1.0, 4.0, 4.0, 4.0, 0.0, 0.0, 0.0)
    assert gpu.mesh(0)["triangle_number"] == real if "triangle_number" in gpu.mesh(0) else True
    with pytest.raises(capi.IonError):
        gpu.import_mesh(str(pb), 1.0, 4.0, 4.0, 4.0, 0.0, 0.0, 0.0)
    gpu.close()


taylor_green_numpy = cases.taylor_green_numpy


def test_scene_helpers(gpu_lib):
    from ionsolver_b200 import lbm as L
    n = 32
    lbm = L.Lbm.setup_taylor_green(n, devices=[0])
    assert np.float32(lbm.config.nu) == np.float32(0.1) and lbm.config.velocity_set == L.VelocitySet.D3Q19
    u, rho = taylor_green_numpy(n)
    d = lbm.domains[0]
    assert np.abs(d.read(cases.FIELD_OF["u"]) - u).max() < 2e-7      # libm vs numpy sinf/cosf: last-bit differences only
    assert np.abs(d.read(cases.FIELD_OF["rho"]) - rho).max() < 5e-7
    got_rho = d.read(cases.FIELD_OF["rho"])
    assert got_rho.min() < 0.0 and got_rho.max() > 2.0              # quirk Q11: the missing parentheses are kept
    lbm.close()
    # split lattice: halo cells stay untouched (zero), interior equals the single-domain field
    split = L.Lbm.setup_taylor_green(n, d_z=2, devices=[0])
    for dom in split.domains:
        r = dom.read(cases.FIELD_OF["rho"]).reshape(dom.n_z, dom.n_y, dom.n_x)
        assert (r[0] == 0).all() and (r[-1] == 0).all()
        z0 = dom.o_z + 1
        assert np.array_equal(r[1:-1].ravel(), got_rho.reshape(n, n, n)[z0:z0 + dom.n_z - 2].ravel())
    split.close()
    cav = L.Lbm.setup_lid_driven_cavity(16, devices=[0])
    fl = cav.domains[0].read(cases.FIELD_OF["flags"]).reshape(16, 16, 16)
    assert (fl[15] == 0x02).all() and (fl[0] == 0x01).all() and fl[1:15, 1:15, 1:15].sum() == 0
    cav.run(5)
    cav.close()
    ch = L.Lbm.setup_charged_fluid(32, 32, 32, lod_depth=2, magnet_stl=os.path.join(cases.STL_DIR, "disk_magnet.stl"), devices=[0])
    dch = ch.domains[0]
    assert (dch.read(cases.FIELD_OF["qc"]) == np.float32(0.002)).all()
    assert (dch.read(cases.FIELD_OF["flags"]) == 0x11).sum() > 50
    assert np.abs(dch.read(cases.FIELD_OF["b_stat"])).max() > 0
    ch.run(3)
    text = ch.dump_cell(0, 5 + 32 * (6 + 32 * 7))
    assert "x: 5, y: 6, z: 7" in text and "b_stat" in text
    ch.close()


def test_error_behaviour(gpu_lib):
    """Option::None buffers, out-of-range I/O and configurations the reference cannot build are errors, not fallbacks."""
    import ctypes
    from ionsolver_b200 import capi, lbm as L
    plain = L.Lbm(L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, float_type=L.FloatType.FP32, n_x=8, n_y=8, n_z=8), devices=[0])
    d = plain.domains[0]
    with pytest.raises(capi.IonError) as e:
        d.read(capi.FIELD_Q)
    assert e.value.code == capi.ION_ERR_ABSENT
    with pytest.raises(capi.IonError) as e:
        d.enqueue_update_e_b_dyn()
    assert e.value.code == capi.ION_ERR_ABSENT
    with pytest.raises(capi.IonError) as e:
        d.enqueue_transfer_extract(capi.TRANSFER_FI, 2, 0)
    assert e.value.code == capi.ION_ERR_ABSENT  # axis not split -> no transfer buffers (domain.rs:311-318)
    with pytest.raises(capi.IonError) as e:
        d.write(capi.FIELD_RHO, np.zeros(8 * 8 * 8 + 1, np.float32))
    assert e.value.code == capi.ION_ERR_RANGE
    for mode in (-1, 3):  # 0 reference order, 1 fast direct sum, 2 FFT convolution
        with pytest.raises(capi.IonError) as e:
            d.set_precompute_mode(mode)
        assert e.value.code == capi.ION_ERR_INVALID
    plain.close()
    bad = [dict(ext_magneto_hydro=True),                                              # MHD without VOLUME_FORCE does not compile in the reference
           dict(ext_magneto_hydro=True, ext_volume_force=True, mhd_lod_depth=5),      # 1<<(1<<5) shifts out of range
           dict(ext_volume_force=True, ext_subgrid_ecr=True)]                          # SUBGRID_ECR lives inside the MHD block
    for kw in bad:
        with pytest.raises(capi.IonError) as e:
            L.Lbm(L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, float_type=L.FloatType.FP32, n_x=32, n_y=32, n_z=32, **kw), devices=[0])
        assert e.value.code == capi.ION_ERR_UNSUPPORTED
    with pytest.raises(capi.IonError):
        L.Lbm(L.LbmConfig(velocity_set=L.VelocitySet.D2Q9, n_x=32, n_y=32, n_z=1, ext_magneto_hydro=True, ext_volume_force=True), devices=[0])
    p = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, n_x=8, n_y=8, n_z=8).make_params(0)
    h = ctypes.c_void_p()
    assert capi.load().ion_domain_create(ctypes.byref(p), 9999, ctypes.byref(h)) == capi.ION_ERR_INVALID


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_decomposition_invariance_and_mass(gpu_lib):
    """256^3 D3Q19 FP32 (the lattice of BASELINE config 2): a periodic Taylor-Green box gives bit-identical rho/u whether
    it runs as one domain or as two z-slabs with halo exchange, and conserves total mass."""
    from ionsolver_b200 import lbm as L
    n, steps = 256, 6
    one = L.Lbm.setup_taylor_green(n, d_z=1, graphics_active=True, devices=[0])
    two = L.Lbm.setup_taylor_green(n, d_z=2, graphics_active=True, devices=[0])
    rho0 = one.domains[0].read(cases.FIELD_OF["rho"]).astype(np.float64).sum()
    one.run(steps)
    two.run(steps)
    one.finish_queues()
    two.finish_queues()
    rho = one.domains[0].read(cases.FIELD_OF["rho"]).reshape(n, n, n)
    u = one.domains[0].read(cases.FIELD_OF["u"]).reshape(3, n, n, n)
    assert abs(rho.astype(np.float64).sum() - rho0) / rho0 < 1e-6
    for dom in two.domains:
        z0 = dom.o_z + 1
        r = dom.read(cases.FIELD_OF["rho"]).reshape(dom.n_z, n, n)[1:-1]
        uu = dom.read(cases.FIELD_OF["u"]).reshape(3, dom.n_z, n, n)[:, 1:-1]
        assert np.array_equal(r, rho[z0:z0 + n // 2]), f"rho of slab {dom.d_i}"
        assert np.array_equal(uu, u[:, z0:z0 + n // 2]), f"u of slab {dom.d_i}"
    one.close()
    two.close()


def test_full_size_mhd_charge_conservation(gpu_lib):
    """256^3 D3Q19 FP32 MHD (BASELINE config 2 lattice, LOD depth 3 to keep the test short): the D3Q7 charge lattice
    conserves total gas charge, the LOD pyramid's finest level sums to total Q, and coarser levels (gather kernel) keep it."""
    from ionsolver_b200 import lbm as L
    n = 256
    lbm = L.Lbm.setup_charged_fluid(n, n, n, lod_depth=3, magnet_stl=None, devices=[0])
    d = lbm.domains[0]
    q0 = d.read(cases.FIELD_OF["qc"]).astype(np.float64).sum()
    lbm.run(3)
    lbm.finish_queues()
    q = d.read(cases.FIELD_OF["qc"]).astype(np.float64)
    assert not np.isnan(q).any()
    lod = d.read(cases.FIELD_OF["qu_lod"]).reshape(-1, 4).astype(np.float64)
    fine = lod[:512]  # single domain: finest level at offset 0 (sim_kernels.cl:667-671)
    assert abs(fine[:, 0].sum() - q.sum()) / abs(q.sum()) < 1e-5
    d.enqueue_lod_part_2_gather()
    lod2 = d.read(cases.FIELD_OF["qu_lod"]).reshape(-1, 4).astype(np.float64)
    assert np.isfinite(lod2).all()
    lbm.close()
    assert q0 > 0


def _global_field(gpu, cfg, field, planes):
    """Whole-lattice array [planes, Nz, Ny, Nx] assembled from the domains' buffers without halo layers (host-side reference
    for the slice read-back)."""
    out = np.zeros((planes, cfg.n_z, cfg.n_y, cfg.n_x), np.float32)
    hx, hy, hz = int(cfg.d_x > 1), int(cfg.d_y > 1), int(cfg.d_z > 1)
    for d in gpu.domains:
        p = d.params
        raw = d.read(field)
        a = (raw.astype(np.float32) if raw.dtype == np.uint8 else raw).reshape(planes, p.nz, p.ny, p.nx)
        sx, sy, sz = slice(hx, p.nx - hx), slice(hy, p.ny - hy), slice(hz, p.nz - hz)
        out[:, p.oz + hz:p.oz + p.nz - hz, p.oy + hy:p.oy + p.ny - hy, p.ox + hx:p.ox + p.nx - hx] = a[:, sz, sy, sx]
    return out


def _decode_png_rgb(path):
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], "chunk CRC"
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 3)


@pytest.mark.parametrize("name", ["z3_d3q19_fp16s_trt", "x2y2_d3q27_fp32", "d3q19_fp32_srt"])
def test_slice_read_back_and_png(name, gpu_lib, tmp_path):
    """one plane of rho / u / flags gathered on the device equals the same plane of the full buffers (bit-exact, halo
    layers removed, every slice mode), and the PNG writer produces a decodable RGB image coloured with the reference's iron map."""
    from ionsolver_b200 import capi
    cfg = dict(cases.single_domain_cases() + cases.multi_domain_cases())[name]
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg)
    cases.upload_inputs(ref, gpu)
    gpu.initialize()
    gpu.do_time_step()
    gpu.finish_queues()
    u = _global_field(gpu, cfg, 2, 3)
    rho = _global_field(gpu, cfg, 1, 1)[0]
    flags = _global_field(gpu, cfg, 3, 1)[0]
    for mode, n_axis, take in ((1, cfg.n_x, lambda a, i: a[..., :, :, i]), (2, cfg.n_y, lambda a, i: a[..., :, i, :]),
                               (3, cfg.n_z, lambda a, i: a[..., i, :, :])):
        for index in sorted({0, n_axis // 2, n_axis - 1}):
            assert same_bits(gpu.read_slice(1, mode, index), take(rho, index)), (mode, index, "rho")
            assert same_bits(gpu.read_slice(3, mode, index), take(flags, index)), (mode, index, "flags")
            for comp in range(3):
                assert same_bits(gpu.read_slice(2, mode, index, component=comp), take(u, index)[comp]), (mode, index, "u", comp)
            mag = np.sqrt(take(u, index)[0] ** 2 + take(u, index)[1] ** 2 + take(u, index)[2] ** 2, dtype=np.float32)
            assert np.allclose(gpu.read_slice(2, mode, index), mag, rtol=2e-7, atol=0)
    # PNG: |u| on the middle z plane
    v_max = float(np.sqrt((u ** 2).sum(0)).max()) or 1.0
    path = tmp_path / "u.png"
    gpu.write_slice_png(path, 2, 3, cfg.n_z // 2, 0.0, v_max)
    img = _decode_png_rgb(path)
    plane = gpu.read_slice(2, 3, cfg.n_z // 2)
    assert img.shape == (cfg.n_y, cfg.n_x, 3)
    lib = capi.load()
    want = np.array([[lib.ion_iron_colormap(float(np.float32(v - np.float32(0.0)) * np.float32(1.0 / np.float32(v_max)))) for v in row]
                     for row in plane[::-1]], np.uint32)
    got = (img[..., 0].astype(np.uint32) << 16) | (img[..., 1].astype(np.uint32) << 8) | img[..., 2]
    assert (got == want).mean() > 0.999  # the host multiplies by 1/(v_max - v_min) in float: allow a last-bit colour step
    with pytest.raises(capi.IonError):
        gpu.read_slice(2, 3, cfg.n_z)          # index outside the lattice
    with pytest.raises(capi.IonError):
        gpu.read_slice(0, 3, 0)                # fi is not a per-cell field
    gpu.close()


def test_vector_kernel_parity_on_every_configuration(gpu_lib):
    """The four-cells-per-thread stream_collide (stream_collide_v4.cuh) is the default only for FP32, Q <= 19, nx >= 256; with
    ION_SC_VEC=1 it runs wherever nx % 4 == 0.  The library reads the switch once, so the bit-exact single- and multi-domain parity
    tests (all velocity sets, storage codecs, SRT/TRT, equilibrium boundaries, force field, split x/y/z) are re-run in a child
    process with the switch set."""
    import os
    import subprocess
    import sys
    if os.environ.get("ION_SC_VEC"):
        pytest.skip("already inside the forced run")
    env = dict(os.environ, ION_SC_VEC="1")
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k",
                          "test_single_domain_bit_exact or test_multi_domain_bit_exact or test_ion_file_round_trip or test_extra_combinations_bit_exact"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]
    assert " passed" in out.stdout


EXTRA = cases.extra_oracle_cases()


@pytest.mark.parametrize("name,cfg", EXTRA, ids=[c[0] for c in EXTRA])
def test_extra_combinations_bit_exact(name, cfg, gpu_lib):
    """Combinations outside the golden matrix (MHD with TRT, MHD on D3Q15 with FP16C, D3Q15 / D2Q9 / D3Q27 with other codec,
    collision and extension mixes, a y-split): every buffer bit-identical to the oracle after initialize and after 4 steps.  The
    oracle is pinned on exactly these cases against the reference build (tests/test_oracle.py); MHD cases run the deterministic
    path, whose E/B and LOD sums are reproducible bit for bit."""
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg, seed=9)
    gpu = product(cfg, deterministic=cfg.ext_magneto_hydro)
    cases.upload_inputs(ref, gpu)
    names = buffer_names(cfg)
    ref.initialize()
    gpu.initialize()
    assert_bit_exact(ref, gpu, cfg, names, "after initialize")
    if cfg.ext_magneto_hydro:
        cases.seed_electron_gas(ref, gpu)
    for _ in range(4):
        ref.do_time_step()
        gpu.do_time_step()
    gpu.finish_queues()
    # a scene that blows up compares NaN payloads, which differ between the host FPU and the GPU: the cases are chosen to stay finite
    assert not any(np.isnan(getattr(d, n)).any() for d in ref.domains for n in names if getattr(d, n).dtype.kind == "f"), "oracle went NaN"
    assert_bit_exact(ref, gpu, cfg, names, "after 4 steps")
    gpu.close()
