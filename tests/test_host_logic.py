"""Host layer without a GPU: Units, LbmConfig defaults, get_device_defines-equivalent parameters, JSON config I/O and
.ion parsing errors, checked against the oracle's restatement of the Rust host (oracle/ref_host.py) and against the
literal values in the reference sources."""
import ctypes
import json
import re
import struct

import numpy as np
import pytest

import cases
from ionsolver_b200 import capi, lbm as L
from oracle import ref_host as rh

UNIT_SETS = [
    (128.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0),     # setup.rs:144
    (1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 1.0, 1.0, 10.0, 1.0),                   # setup.rs:205
    (1.0, 1.0, 1.0, 1.0, 1.0, 0.01, 10000.0, 10E-8, 0.0000000000001, 50000.0),  # setup.rs:282
    (128.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 10e-8, 1.0, 50000.0),            # setup.rs:397
]


@pytest.mark.parametrize("args", UNIT_SETS)
def test_units_match_oracle_bit_for_bit(args):
    u = L.Units()
    u.set(*args)
    r = rh.Units()
    r.set(*args)
    for name in ("m", "kg", "s", "a", "k"):
        assert np.float32(getattr(u, name)) == getattr(r, name), name
    pairs = [("epsilon_0_lu", r.epsilon_0_lu()), ("ke_lu", r.ke_lu()), ("mu_0_lu", r.mu_0_lu()), ("kkge_lu", r.kkge_lu()),
             ("kimg_lu", r.kimg_lu()), ("kveV_lu", r.kveV_lu()), ("kkBme_lu", r.kkBme_lu()), ("keabs_lu", r.keabs_lu()),
             ("kme_lu", r.kme_lu())]
    for name, want in pairs:
        got = np.float32(getattr(u, name)())
        assert got == want or (np.isnan(got) and np.isnan(want)) or (np.isinf(got) and np.isinf(want)), (name, got, want)
    for name, v in (("len_si_lu", 0.37), ("nu_si_lu", 1.48e-5), ("charge_si_lu", 2.1e-13), ("mag_flux_si_lu", 0.0875),
                    ("e_field_si_lu", 1000.0), ("magnetization_si_lu", 1.0e6), ("time_lu_si", 2.45e9)):
        assert np.float32(getattr(u, name)(v)) == getattr(r, name)(v), name


def test_lbm_config_defaults_are_the_reference_defaults():
    c = L.LbmConfig()  # mod.rs:103-135
    assert (c.velocity_set, c.relaxation_time, c.float_type) == (L.VelocitySet.D2Q9, L.RelaxationTime.Srt, L.FloatType.FP16S)
    assert (c.n_x, c.n_y, c.n_z, c.d_x, c.d_y, c.d_z) == (1, 1, 1, 1, 1, 1)
    assert np.float32(c.nu) == np.float32(1.0) / np.float32(6.0)
    assert c.mhd_lod_depth == 4 and c.graphics_config.graphics_active is True
    d = capi.IonLbmConfig()
    capi.load().ion_lbm_config_default(d)
    assert L.LbmConfig.from_c(d) == c


ALL = cases.all_cases()


@pytest.mark.parametrize("name,cfg", ALL, ids=[c[0] for c in ALL])
def test_make_params_matches_get_device_defines(name, cfg):
    """IonParams (the struct form of get_device_defines, domain.rs:736-858) against the oracle's #define block."""
    pc = cases.to_lbm_config(cfg)
    rcfg = rh.RefLbm.__new__(rh.RefLbm)  # only for the resolution rounding rule
    n_d = cfg.d_x * cfg.d_y * cfg.d_z
    for d in range(n_d):
        x, y, z = rh.domain_coords(d, cfg.d_x, cfg.d_y)
        g = rh.domain_geometry(cfg, x, y, z, d)
        p = pc.make_params(d)
        assert (p.nx, p.ny, p.nz) == (g.n_x, g.n_y, g.n_z)
        assert (p.ox, p.oy, p.oz) == (g.o_x, g.o_y, g.o_z)
        assert (p.dx, p.dy, p.dz, p.di) == (cfg.d_x, cfg.d_y, cfg.d_z, d)
        assert (p.n_lod, p.n_lod_own, p.lod_depth) == (g.n_lod, g.n_lod_own, cfg.mhd_lod_depth)
        defines = dict(re.findall(r"#define (\w+) (\S+)", rh.device_defines(cfg, g)))

        def lit(key):
            return np.float32(float(defines[key].rstrip("f")))
        assert np.float32(p.w) == lit("DEF_W")
        if cfg.ext_magneto_hydro:
            for key, val in (("DEF_KE", p.ke), ("DEF_KMU", p.kmu), ("DEF_KMU0", p.kmu0), ("DEF_KKGE", p.kkge), ("DEF_KIMG", p.kimg),
                             ("DEF_KVEV", p.kvev), ("DEF_KME", p.kme), ("DEF_WQ", p.wq)):
                assert np.float32(val) == lit(key), key
        ext = (1 * cfg.ext_equilibrium_boudaries | 2 * cfg.ext_volume_force | 4 * cfg.ext_force_field | 8 * cfg.ext_magneto_hydro
               | 16 * cfg.ext_subgrid_ecr | 32 * cfg.graphics_active)
        assert p.ext == ext
        if cfg.ext_subgrid_ecr:
            for key, val in (("DEF_KKBME", p.kkbme), ("DEF_KEABS", p.keabs)):
                assert np.float32(val) == lit(key), key


def test_lod_counts_of_the_reference_scenes():
    # setup_bfield_spin (setup.rs:142-156): 128x128x256, d_z = 2, depth 4 -> own 4681, + one foreign level 3 (512)
    c = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, n_x=128, n_y=128, n_z=256, d_z=2, ext_volume_force=True, ext_magneto_hydro=True)
    p = c.make_params(1)
    assert (p.n_lod_own, p.n_lod) == (4681, 4681 + 512)
    assert (p.nx, p.ny, p.nz, p.oz) == (128, 128, 130, 127)
    # resolution is rounded down to a multiple of the domain count (mod.rs:167-179)
    c = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, n_x=10, n_y=10, n_z=11, d_z=2)
    assert c.make_params(0).nz == 5 + 2


def test_velocity_set_tables():
    assert [rh.SET_VALUES[k] for k in ("D2Q9", "D3Q15", "D3Q19", "D3Q27")] == [(2, 9, 3), (3, 15, 5), (3, 19, 5), (3, 27, 9)]
    assert capi.SET_VALUES == {0: (2, 9, 3), 1: (3, 15, 5), 2: (3, 19, 5), 3: (3, 27, 9)}  # types.rs:37-46
    assert (L.FloatType.size_of(L.FloatType.FP16S), L.FloatType.size_of(L.FloatType.FP16C), L.FloatType.size_of(L.FloatType.FP32)) == (2, 2, 4)


def test_json_config_round_trip_and_serde_shape():
    c = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, relaxation_time=L.RelaxationTime.Trt, float_type=L.FloatType.FP16C, n_x=128,
                    n_y=256, n_z=128, d_z=2, nu=0.05, f_x=1e-5, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3,
                    ecr_freq=0.5, run_steps=77, graphics_config=L.GraphicsConfig(False))
    c.units.set(128.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    text = c.to_json()
    j = json.loads(text)  # what serde_json::to_vec(&LbmConfig) produces: unit-variant enums as strings, nested structs as objects
    assert j["velocity_set"] == "D3Q19" and j["relaxation_time"] == "Trt" and j["float_type"] == "FP16C"
    assert j["units"]["prop"] == "H" and j["graphics_config"]["graphics_active"] is False
    assert set(j) == {"velocity_set", "relaxation_time", "float_type", "units", "n_x", "n_y", "n_z", "d_x", "d_y", "d_z", "nu", "f_x", "f_y",
                      "f_z", "ext_equilibrium_boudaries", "ext_volume_force", "ext_force_field", "ext_magneto_hydro", "ext_subgrid_ecr",
                      "mhd_lod_depth", "ecr_freq", "ecr_field_strength", "graphics_config", "run_steps"}
    assert "keyframes" in j["graphics_config"] and j["graphics_config"]["camera_width"] == 1920  # GraphicsConfig::new, graphics.rs:158-193
    back = L.LbmConfig.from_json(text)
    assert bytes(back.to_c()) == bytes(c.to_c())  # every f32 survives the text round trip bit for bit
    # a config written by a real IonSolver build (extra graphics fields, any key order, whitespace) parses
    j["graphics_config"]["v_max"] = 500.0
    j["graphics_config"]["keyframes"] = [{"time": 3, "repeat": True, "cam_zoom": 1.5, "note": "a } brace in a string"}]
    again = L.LbmConfig.from_json(json.dumps(dict(reversed(list(j.items()))), indent=2))
    assert bytes(again.to_c()) == bytes(c.to_c())
    with pytest.raises(capi.IonError):
        L.LbmConfig.from_json('{"velocity_set":"D4Q99"}')
    with pytest.raises(capi.IonError):
        L.LbmConfig.from_json('{"n_x": }')


def ion_header(ft=2, n=(4, 4, 4), d=(1, 1, 1), ext=0, vs=2):
    b = b"IonSolver setup\n" + bytes([vs, 0, ft]) + struct.pack("<4f", 1, 1, 1, 1) + struct.pack("<3I", *n) + b"\0" + \
        struct.pack("<3I", *d) + struct.pack("<4f", 1 / 6, 0, 0, 0) + bytes([ext, 4])
    assert len(b) == 16 + 62  # FILE_LAYOUT.txt
    return b


def test_ion_decode_rejects_bad_files_before_touching_the_gpu():
    lib = capi.load()
    h = ctypes.c_void_p()
    c = capi.IonLbmConfig()
    lib.ion_lbm_config_default(c)
    bad = b"NotIonSolver....\n" + b"\0" * 100
    assert lib.ion_lbm_decode(bad, len(bad), c, 1, None, 0, ctypes.byref(h)) == capi.ION_ERR_INVALID
    assert b"Invalid Format!" in lib.ion_last_error_string()  # file.rs:50-52
    trunc = ion_header()[:40]
    assert lib.ion_lbm_decode(trunc, len(trunc), c, 1, None, 0, ctypes.byref(h)) == capi.ION_ERR_RANGE
    short = ion_header() + b"\0" * 10  # sections missing
    assert lib.ion_lbm_decode(short, len(short), c, 1, None, 0, ctypes.byref(h)) == capi.ION_ERR_RANGE
    assert not h.value
    # the header fields land in the config before the payload check (decode fills `config` in place, file.rs:54-104)
    hdr = ion_header(ft=1, n=(8, 6, 4), d=(1, 1, 2), ext=0b1010)
    lib.ion_lbm_decode(hdr, len(hdr), c, 0, None, 0, ctypes.byref(h))
    assert (c.n_x, c.n_y, c.n_z, c.d_z, c.ext_volume_force, c.ext_magneto_hydro, c.ext_force_field) == (8, 6, 4, 2, 1, 1, 0)
    assert c.float_type == L.FloatType.FP16C      # FILE_LAYOUT / types.rs discriminant
    lib.ion_lbm_decode(hdr, len(hdr), c, 1, None, 0, ctypes.byref(h))
    assert c.float_type == L.FloatType.FP16S      # the reference decoder's swapped table, file.rs:68-73


def test_png_writer_and_iron_colormap(tmp_path):
    """Host-only parts of the slice writer: the PNG encoder round-trips through zlib, and the colour map is the
    reference's iron_colormap (graphics_kernels.cl:412-428 with color_from_floats :93-95)."""
    import zlib
    lib = capi.load()

    def iron(x):  # graphics_kernels.cl:412-428 in float32
        f = np.float32
        x = f(min(max(f(4.0) * (f(1.0) - f(x)), f(0.0)), f(4.0)))
        r, g, b = f(1.0), f(0.0), f(0.0)
        if x < f(0.66666667):
            g, b = f(1.0), f(1.0) - x * f(1.5)
        elif x < f(2.0):
            g = f(1.5) - x * f(0.75)
        elif x < f(3.0):
            r, b = f(2.0) - x * f(0.5), x - f(2.0)
        else:
            r, b = f(2.0) - x * f(0.5), f(4.0) - x
        ch = lambda v: int(min(max(int(np.float32(255.0) * v + np.float32(0.5)), 0), 255))
        return ch(r) << 16 | ch(g) << 8 | ch(b)

    for x in np.linspace(-0.2, 1.2, 141):
        got, want = lib.ion_iron_colormap(float(np.float32(x))), iron(np.float32(x))
        assert all(abs(((got >> s) & 255) - ((want >> s) & 255)) <= 1 for s in (16, 8, 0)), (x, hex(got), hex(want))
    assert lib.ion_iron_colormap(1.0) == 0xFFFFFF and lib.ion_iron_colormap(0.0) == 0x000000

    rng = np.random.default_rng(5)
    for w, h in ((1, 1), (7, 3), (300, 91)):  # the last one needs more than one stored deflate block (65535 bytes)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        path = tmp_path / f"t{w}x{h}.png"
        capi.check(lib.ion_write_png_rgb(str(path).encode(), rgb.ctypes.data, w, h))
        data = path.read_bytes()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        pos, idat = 8, b""
        while pos < len(data):
            n, typ = struct.unpack(">I4s", data[pos:pos + 8])
            body = data[pos + 8:pos + 8 + n]
            assert zlib.crc32(typ + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
            if typ == b"IHDR":
                assert struct.unpack(">IIBBBBB", body) == (w, h, 8, 2, 0, 0, 0)
            if typ == b"IDAT":
                idat += body
            pos += 12 + n
        raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
        assert (raw[:, 0] == 0).all() and (raw[:, 1:].reshape(h, w, 3) == rgb).all()


def test_fp16c_encoder_identity_on_cpu(tmp_path):
    """The device FP16C encoder is one round-toward-zero multiplication (lattice.cuh); tests/tools/fp16c_encode_check.c proves it
    equal to the reference's bit assembly (sim_kernels.cl:79-84) with the host FPU in FE_TOWARDZERO mode.  Here: every 97th bit
    pattern plus dense sweeps of the binade edges (the full 2^32 run takes a minute: `./fp16c_encode_check`)."""
    import os
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "fp16c_encode_check.c")
    exe = tmp_path / "fp16c_encode_check"
    subprocess.run(["gcc", "-O2", "-frounding-math", "-o", str(exe), src, "-lm"], check=True)
    out = subprocess.run([str(exe), "97"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-500:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["mismatches_non_nan"] == 0 and res["checked"] > 44_000_000


@pytest.mark.parametrize("args", [("4", "32", "32", "34", "2"), ("3", "32", "16", "26", "2"), ("4", "16", "32", "16", "1"),
                                  ("3", "32", "16", "26", "2", "1"), ("3", "24", "16", "16", "1", "1")],
                         ids=["depth4_slab_with_halo_window", "depth3_slab_with_halo_window", "depth4_single",
                              "depth3_slab_mirrored_spectra_two_sets", "depth3_single_mirrored_spectra_odd_block"])
def test_polyphase_fft_field_update_on_cpu(tmp_path, args):
    """update_e_b_dynamic as a polyphase FFT convolution (ionsolver_b200/csrc/eb_fft_core.cuh): the phase functions the CUDA
    kernels are made of are run thread by thread on the CPU (tests/tools/eb_fft_emul.cpp) and compared with a direct
    double-precision evaluation of the reference's own-LOD loop (sim_kernels.cl:940-955): window quirk Q5, self-skip, halo
    layers, solid cells left untouched, the extra z window of halo-inclusive slabs.  Relative L2 below 1e-5 (measured 2e-7).
    The `mirrored` cases run the tasks with 2 ox > dsx on their partner's kernel spectra (main_phase_product_mirror /
    main_phase_product2_mirror: reflection in frequency space, plus a one-block phase for the neighbour's level)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "eb_fft_emul"
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", cuda_inc, os.path.join(root, "tests", "tools", "eb_fft_emul.cpp"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=600)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0, res
    assert res["rel_l2_E"] < 1e-6 and res["rel_l2_B"] < 1e-6 and res["untouched_bad"] == 0


@pytest.mark.parametrize("farb,r_min,bound", [(8, 242.0, 1.6e-5), (4, 65.0, 6.4e-5)], ids=["8cubed_blocks", "4cubed_blocks"])
def test_far_slab_taylor_path_on_cpu(tmp_path, farb, r_min, bound):
    """Far slabs of update_e_b_dynamic (sim_kernels.cl:957-983, slabs two and more below): second-order Taylor tensors per FARB^3 block
    of cells (eb_fft_core.cuh::far_accumulate), evaluated per cell in the row-reduced form a + dx (b + c dx) that k_eb_combine uses
    (far_reduce_x / far_eval_x).  tests/tools/eb_far_emul.cpp runs those functions on the CPU at the smallest source distance the host
    code admits for each block size (eb_fft.cu::far_set: R >= 40 x 6.06 for 8^3, R >= 25 x 2.6 for 4^3) against a double-precision
    direct sum: error below the documented bound relative to the largest far-field value, and the reduced form equals the full one."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    exe = tmp_path / "eb_far_emul"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", cuda_inc, os.path.join(root, "tests", "tools", "eb_far_emul.cpp"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(farb), str(r_min)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["taylor_err_rel"] < bound and res["x_reduced_err_rel"] < bound and res["forms_differ_rel"] < 1e-6, res


def test_mirrored_task_layout_on_cpu(tmp_path):
    """Host side of the mirrored / streamed kernel spectra (ionsolver_b200/csrc/eb_fft_layout.hpp, used by eb_fft.cu::eb_fft_create):
    tests/tools/eb_layout_check.cpp lays out 768 combinations of block shape, extra z window and slot capacity and checks that every
    task appears once, that a mirrored task's slot is the canonical task with x offset dsx - ox and equal (oy, oz, wz) in its own
    batch, and that batches respect the capacity."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    exe = tmp_path / "eb_layout_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", cuda_inc, os.path.join(root, "tests", "tools", "eb_layout_check.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and res["violations"] == 0 and res["layouts_checked"] >= 700, res


def test_fft_precompute_geometry_in_numpy():
    """Index logic of the FFT precompute mode (mesh_kernels.cu::launch_precompute_fft), restated with numpy transforms: sources inside
    their bounding box [mn, mn + S), outputs 0 .. L-1, transform lengths P = smooth7(L + S - 1), kernel array K[j mod P] = G(j - mn) for
    j in [-(S-1), L-1] and zero elsewhere / at d = 0.  The circular convolution then equals the direct sum of M . d / |d|^3 on every
    output (no wrap-around), for a box that touches the lattice edge as well."""
    rng = np.random.default_rng(1)

    def smooth7(n):
        while True:
            m = n
            for p in (2, 3, 5, 7):
                while m % p == 0:
                    m //= p
            if m == 1:
                return n
            n += 1

    assert [smooth7(n) for n in (1, 11, 13, 121, 515, 1031)] == [1, 12, 14, 125, 525, 1050]
    for L, mn, S in ((np.array([14, 11, 9]), np.array([3, 2, 5]), np.array([4, 6, 3])), (np.array([8, 9, 10]), np.array([0, 0, 7]), np.array([8, 2, 3]))):
        P = np.array([smooth7(int(L[i] + S[i] - 1)) for i in range(3)])
        cells = {tuple(mn), tuple(mn + S - 1)}
        while len(cells) < 12:
            cells.add(tuple(int(v) for v in mn + rng.integers(0, S)))
        src = [(np.array(c), rng.normal(size=3)) for c in sorted(cells)]
        direct = np.zeros(L[::-1])
        for z in range(L[2]):
            for y in range(L[1]):
                for x in range(L[0]):
                    for c, m in src:
                        d = np.array([x, y, z]) - c
                        r2 = float((d * d).sum())
                        if r2 > 0:
                            direct[z, y, x] += float(d @ m) / r2 ** 1.5
        acc = np.zeros((P[2], P[1], P[0] // 2 + 1), complex)
        idx = [np.arange(P[a]) for a in range(3)]
        j = [np.where(idx[a] < L[a], idx[a], idx[a] - P[a]) for a in range(3)]
        jz, jy, jx = np.meshgrid(j[2], j[1], j[0], indexing="ij")
        valid = (jx > -S[0]) & (jy > -S[1]) & (jz > -S[2])
        d = [jx - mn[0], jy - mn[1], jz - mn[2]]
        r2 = (d[0] ** 2 + d[1] ** 2 + d[2] ** 2).astype(float)
        inv = np.where(r2 > 0, 1.0 / np.maximum(r2, 1.0) ** 1.5, 0.0) * valid
        for comp in range(3):
            R = np.zeros(P[::-1])
            for c, m in src:
                R[c[2] - mn[2], c[1] - mn[1], c[0] - mn[0]] = m[comp]
            acc += np.fft.rfftn(R) * np.fft.rfftn(d[comp] * inv)
        out = np.fft.irfftn(acc, s=tuple(int(v) for v in P[::-1]), axes=(0, 1, 2))[: L[2], : L[1], : L[0]]
        assert np.abs(out - direct).max() < 1e-12 * max(1.0, np.abs(direct).max())


def test_voxeliser_fill_rule_equals_the_reference_state_machine():
    """The CUDA voxeliser (mesh_kernels.cu) keeps the ray hits of a column as an XOR bitmap and fills by a closed form instead of
    the reference's sorted list + state machine (sim_kernels.cl:1194-1230).  Both rules, restated in Python, on 100 000 random hit
    lists -- duplicates, more than 64 hits (only the first 64 are stored), hits beyond the column, every parity of the backward
    count: identical sets of filled cells."""
    import random

    def reference_fill(hits, behind, h0, hmax):
        count = len(hits)
        if count == 0:
            return set()
        dist = sorted(hits[:64])
        dist += [0] * (64 - len(dist))
        inside = (count % 2 == 1) and (behind % 2 == 1)
        k = 1 if (count % 2) != (behind % 2) else 0
        hmesh = h0 + dist[min(count - 1, 63)]
        out = set()
        for h in range(h0, hmax):
            while k < count and h > h0 + dist[min(k, 63)]:
                inside = not inside
                k += 1
            inside = inside and (k < count and h < hmesh)
            if inside:
                out.add(h)
        return out

    def bitmap_fill(hits, behind, h0, hmax, words):
        count = len(hits)
        if count == 0:
            return set()
        bits, dmin, dmax = [0] * (32 * words), 1 << 32, 0
        for i, d in enumerate(hits):
            if i < 64:
                if d < 32 * words:
                    bits[d] ^= 1
                dmin, dmax = min(dmin, d), max(dmax, d)
        inside0 = (count % 2 == 1) and (behind % 2 == 1)
        skip_nearest = (count % 2) != (behind % 2)
        out, below_odd = set(), False
        for h in range(h0, min(hmax, h0 + dmax)):
            m = h - h0
            if m > 0:
                below_odd ^= bool(bits[m - 1])
            toggled = ((not below_odd) if m > dmin else False) if skip_nearest else below_odd
            if inside0 != toggled:
                out.add(h)
        return out

    rng = random.Random(1)
    for _ in range(100000):
        ext = rng.choice([5, 17, 40, 100])
        h0 = rng.randint(0, 3)
        hmax = h0 + rng.randint(0, ext)
        n = rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 8, 63, 64, 65, 70, 130]) if rng.random() < 0.3 else rng.randint(0, 9)
        hits = [rng.randint(0, ext + 5) for _ in range(n)]
        behind = rng.randint(0, 5)
        words = (hmax - h0 + 31) // 32 + 1
        assert reference_fill(hits, behind, h0, hmax) == bitmap_fill(hits, behind, h0, hmax, words), (hits, behind, h0, hmax)
