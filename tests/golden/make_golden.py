"""Generates tests/golden/reference_vectors.json (+ small .npy arrays) from the REFERENCE ITSELF.

The reference ships no tests, golden vectors or fixtures, so the pins are produced here: the
reference's own kernel source (/root/reference/src/kernels/sim_kernels.cl) is compiled for the host by
oracle/build_ref.py and driven through the host sequence of src/lbm/mod.rs on seeded synthetic inputs
(tests/cases.py).  What is recorded per case: SHA-256 of every output buffer (bit-exact pins for flags, DDFs, halo
buffers, neighbour tables, codecs) plus float64 sums for a readable cross-check.

Only runs where /root/reference exists (this container); the outputs are committed so that the GPU box, which has
no /root/reference, can check both the C restatement (oracle/lbm_oracle.c) and the CUDA kernels against them.

Run: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import build_ref, ref_host as rh  # noqa: E402

STEPS = 8


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(a):
    a = np.ascontiguousarray(a)
    out = {"sha256": sha(a), "dtype": str(a.dtype), "size": int(a.size)}
    if a.dtype.kind == "f":
        out["sum"] = float(np.nansum(a.astype(np.float64)))
        out["nan"] = int(np.isnan(a).sum())
    return out


def buffers_of(d, cfg):
    names = ["fi", "rho", "u", "flags"]
    if cfg.d_x * cfg.d_y * cfg.d_z > 1:
        names += ["transfer_p", "transfer_m"]
    if cfg.ext_magneto_hydro:
        names += ["ei", "fqi", "qc", "e_dyn", "b_dyn", "qu_lod"]
    if cfg.ext_subgrid_ecr:
        names += ["eti", "et"]
    return {n: digest(getattr(d, n)) for n in names}


def run_case(cfg):
    lbm = rh.RefLbm(cfg, threads=1)  # one thread: the LOD float atomics (quirk Q6) then sum in cell order
    cases.fill_inputs(lbm, cfg)
    rec = {"inputs": [{n: sha(getattr(d, n)) for n in ("rho", "u", "flags")} for d in lbm.domains]}
    lbm.initialize()
    rec["after_initialize"] = [buffers_of(d, cfg) for d in lbm.domains]
    if cfg.ext_magneto_hydro:
        cases.seed_electron_gas(lbm)  # see cases.RHO_E0: without it the reference NaNs out within three steps
    for _ in range(STEPS):
        lbm.do_time_step()
    rec["after_steps"] = [buffers_of(d, cfg) for d in lbm.domains]
    nans = sum(v.get("nan", 0) for d in rec["after_steps"] for v in d.values())
    assert nans == 0, f"NaN in the reference run: {nans}"
    rec["steps"] = STEPS
    return rec


def codec_vectors():
    """FP16S / FP16C storage codecs (domain.rs:773-780, sim_kernels.cl:79-90): all 65 536 codes decoded, and a fixed
    set of floats encoded."""
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.uniform(-2.0, 2.0, 200000), rng.standard_normal(100000) * 1e-3, rng.standard_normal(100000) * 1e-6,
        np.array([0.0, -0.0, 1.0, -1.0, 1.99951168, 2.0, 6.10351562e-5, 2.98023224e-8, 1e-9, 65504.0 / 32768.0, 3.0])]).astype(np.float32)
    out = {"floats_sha256": sha(x)}
    for ft in ("FP16S", "FP16C"):
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type=ft, n_x=4, n_y=4, n_z=4)
        d = rh.RefDomain(cfg, 0, 0, 0, 0)
        codes = np.arange(65536, dtype=np.uint16)
        if ft == "FP16S":  # skip NaN/Inf halves: their float image is not a single bit pattern across compilers
            h = codes.view(np.float16)
            codes = codes[np.isfinite(h)]
        out[ft] = {"decode_all_codes": digest(d.codec(codes, 1)), "encode_floats": digest(d.codec(x, 0))}
    return out


def neighbor_vectors():
    """neighbors() of sim_kernels.cl:260-302 for every velocity set on odd sizes, all cells."""
    out = {}
    for vs, dims in (("D2Q9", (7, 5, 1)), ("D3Q15", (5, 3, 7)), ("D3Q19", (7, 3, 5)), ("D3Q27", (3, 5, 7))):
        cfg = rh.RefConfig(velocity_set=vs, float_type="FP32", n_x=dims[0], n_y=dims[1], n_z=dims[2])
        d = rh.RefDomain(cfg, 0, 0, 0, 0)
        tab = np.stack([d.neighbors(n) for n in range(d.g.n)])
        np.save(os.path.join(HERE, f"neighbors_{vs}.npy"), tab.astype(np.uint32))
        out[vs] = {"dims": dims, "sha256": sha(tab.astype(np.uint32))}
    return out


def voxel_vectors():
    """voxelize_mesh + psi/static_b/static_e of the reference on the synthetic STLs (tests/golden/stl)."""
    out = {}
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=40, n_z=44, nu=0.05, ext_volume_force=True,
                       ext_magneto_hydro=True, mhd_lod_depth=2)
    cfg.units.set(48.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    lbm = rh.RefLbm(cfg, threads=0)
    meshes = [("disk_magnet.stl", "Magnet", (0.0, 1000000.0, 0.0), (24.1, 8.1, 22.0)),
              ("ring_magnet.stl", "Magnet", (0.0, 500000.0, 0.0), (24.1, 30.3, 22.0)),
              ("tube.stl", "Solid", None, (24.001, -6.0, 22.0)),
              ("plate1.stl", "Charged", 1.0e-13, (24.0, -8.0, 22.0)),
              ("plate2.stl", "ChargedECR", -1.0e-13, (24.0, -8.0, 22.0))]
    for i, (f, kind, val, origin) in enumerate(meshes):
        data = open(os.path.join(cases.STL_DIR, f), "rb").read()
        lbm.import_mesh(data, 1.0, origin[0], origin[1], origin[2], 0.0, 0.0, 0.0)
        lbm.voxelise_mesh(i, kind, val)
        out[f] = {"flags_after": sha(lbm.domains[0].flags), "cells": int((lbm.domains[0].flags != 0).sum()),
                  "p_min": [float(v) for v in lbm.meshes[i].p_min], "p_max": [float(v) for v in lbm.meshes[i].p_max]}
    lbm.precompute_B()
    lbm.precompute_E()
    d = lbm.domains[0]
    out["b_stat"] = digest(d.b_stat)
    out["e_stat"] = digest(d.e_stat)
    out["psi"] = digest(d.e_dyn[: (cfg.n_x + 2) * (cfg.n_y + 2) * (cfg.n_z + 2)])
    out["config"] = {"n": [cfg.n_x, cfg.n_y, cfg.n_z], "meshes": [[m[0], m[1], m[2], list(m[3])] for m in meshes]}
    return out


REF_STL = os.path.join(HERE, "stl", "ref")  # the reference's own stl/*.stl, copied as test fixtures


def deeva_parts(scale):
    """Meshes, origins (for a lattice `scale` x 128 cells wide) and model types of setup_deeva_test, setup.rs:423-441."""
    s = float(scale)
    return [("deeva_disk_magnet.stl", (64.001 * s, 0.0, 64.0 * s), "Magnet", (0.0, 1000000.0, 0.0)),
            ("deeva_inlet.stl", (64.0 * s, 0.0, 64.0 * s), "Solid", None),
            ("deeva_quartz_tube.stl", (64.001 * s, 0.0, 64.0 * s), "Solid", None),
            ("deeva_ring_magnet.stl", (64.001 * s, -0.5 * s, 64.0 * s), "Magnet", (0.0, 500000.0, 0.0)),
            ("deeva_e_plate1.stl", (64.0 * s, 0.0, 64.0 * s), "ChargedECR", 0.00000000000021844213 / 2.0),
            ("deeva_e_plate2.stl", (64.0 * s, 0.0, 64.0 * s), "ChargedECR", -0.00000000000021844213 / 2.0)]


def ref_stl_vectors():
    """The reference's own STL assets through the reference's own voxeliser / static-field kernels:
    (1) setup_deeva_test (setup.rs:395-446) at its own size 128 x 256 x 128: flags after each of the six meshes;
    (2) the same scene at half the size (64 x 128 x 64, so that the O(N^2) psi / static_e kernels finish on a CPU):
        flags, psi, B_stat, E_var;
    (3) BASELINE cfg2's magnet: disk-magnet.stl repositioned into a 256^3 lattice (pattern of setup.rs:383-384): flags."""
    out = {}
    for tag, scale, fields in (("deeva_128x256x128", 1.0, False), ("deeva_64x128x64", 0.5, True)):
        n = (int(128 * scale), int(256 * scale), int(128 * scale))
        cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=n[0], n_y=n[1], n_z=n[2], ext_volume_force=True,
                           ext_magneto_hydro=True, ext_subgrid_ecr=True, mhd_lod_depth=2)
        cfg.units.set(128.0 * scale, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 10e-8, 1.0, 50000.0)  # setup.rs:397
        cfg.nu = float(cfg.units.nu_si_lu(0.05))
        lbm = rh.RefLbm(cfg, threads=0)
        rec = {"n": list(n), "meshes": []}
        for i, (f, o, kind, val) in enumerate(deeva_parts(scale)):
            lbm.import_mesh(open(os.path.join(REF_STL, f), "rb").read(), 1.0, o[0], o[1], o[2], 0.0, 0.0, 0.0)
            lbm.voxelise_mesh(i, kind, val)
            fl = lbm.domains[0].flags
            rec["meshes"].append({"file": f, "origin": list(o), "kind": kind, "value": val, "flags_after": sha(fl),
                                  "cells_flagged": int((fl != 0).sum())})
        if fields:
            lbm.precompute_B()
            lbm.precompute_E_ECR()
            d = lbm.domains[0]
            rec["psi"] = digest(d.e_dyn[: (n[0] + 2) * (n[1] + 2) * (n[2] + 2)])
            rec["b_stat"] = digest(d.b_stat)
            rec["e_var"] = digest(d.e_var)
        out[tag] = rec
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=256, n_y=256, n_z=256, ext_volume_force=True, ext_magneto_hydro=True,
                       mhd_lod_depth=4)
    cfg.units.set(256.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    lbm = rh.RefLbm(cfg, threads=0)
    lbm.import_mesh_reposition(open(os.path.join(REF_STL, "disk-magnet.stl"), "rb").read(), 128.1, 128.1, 128.0, 0.0, 0.0, 0.0, 127.0)
    lbm.voxelise_mesh(0, "Magnet", (0.0, 1000000.0, 0.0))
    fl = lbm.domains[0].flags
    out["cfg2_disk_magnet_256"] = {"n": [256, 256, 256], "flags_after": sha(fl), "cells_flagged": int((fl != 0).sum()),
                                   "p_min": [float(v) for v in lbm.meshes[0].p_min], "p_max": [float(v) for v in lbm.meshes[0].p_max]}
    return out


def main():
    if build_ref.reference_root() is None:
        raise SystemExit("needs /root/reference: golden vectors come from the reference's own kernels")
    gold = {"generator": "tests/golden/make_golden.py", "reference_kernel_sha256": hashlib.sha256(
        open(os.path.join(build_ref.reference_root(), "src", "kernels", "sim_kernels.cl"), "rb").read()).hexdigest(),
        "cases": {}}
    for name, cfg in cases.all_cases():
        print("case", name, flush=True)
        gold["cases"][name] = run_case(cfg)
    gold["codecs"] = codec_vectors()
    gold["neighbors"] = neighbor_vectors()
    gold["voxelize"] = voxel_vectors()
    gold["ref_stl"] = ref_stl_vectors()
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "reference_vectors.json"))


if __name__ == "__main__":
    main()
