#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations other than the headline one (cfg2 is bench.py), one GPU:
   cfg1  64^3 D3Q19 FP32, Taylor-Green (setup.rs:92) and a lid-driven cavity, no MHD
   cfg3  512x256x256 D3Q19 FP32 MHD thruster-like scene (ring magnet + disk magnet + quartz tube from synthetic STLs,
         voxelize_mesh + precompute_B), full E/B update
   cfg4  512^3 D3Q27 FP16S MHD (the 1-GPU point of the strong-scaling series)
   cfg5  D3Q19 FP16C MHD at ~0.54 G cells (2048 x 2048 x 128; --full tries the 1.07 G-cell slab that fills 180 GB)
One JSON line per configuration: MLUPs/s over whole time steps (CUDA events on the domain stream) and the per-kernel times.
usage: python tests/tools/bench_configs.py [--only cfg1,cfg3] [--full]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
STL = os.path.join(ROOT, "tests", "golden", "stl")


def fill(d, field, value, dtype=np.float32, chunk=1 << 26):
    """constant fill without a domain-sized host array"""
    total = d.size(field) // np.dtype(dtype).itemsize
    buf = np.full(min(chunk, total), value, dtype)
    off = 0
    while off < total:
        k = min(chunk, total - off)
        d.write(field, buf[:k], off * buf.itemsize)
        off += k


def timed_steps(lbm, steps, warmup):
    import torch
    d = lbm.domains[0]
    stream = torch.cuda.ExternalStream(d.stream(), device=0)
    for _ in range(warmup):
        lbm.do_time_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lbm.finish_queues()
    e0.record(stream)
    for _ in range(steps):
        lbm.do_time_step()
    e1.record(stream)
    lbm.finish_queues()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # per-kernel
    mhd = lbm.config.ext_magneto_hydro
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = lbm.get_time_step()
    ev[0].record(stream)
    if mhd:
        d.enqueue_clear_qu_lod()
    ev[1].record(stream)
    d.enqueue_stream_collide(t)
    ev[2].record(stream)
    if mhd:
        d.enqueue_update_e_b_dyn()
    ev[3].record(stream)
    lbm.set_time_step(t + 1)
    lbm.finish_queues()
    torch.cuda.synchronize()
    return ms, {"stream_collide": ev[1].elapsed_time(ev[2]), "update_e_b_dynamic": ev[2].elapsed_time(ev[3]) if mhd else 0.0}


def report(name, lbm, steps, warmup, extra=None):
    c = lbm.config
    cells = c.n_x * c.n_y * c.n_z
    ms, k = timed_steps(lbm, steps, warmup)
    d = lbm.domains[0]
    rho = d.read(1, 0, 4 * min(d.n, 1 << 20))
    line = {"config": name, "cells": cells, "steps": steps, "ms_per_step": round(ms, 4), "mlups": round(cells / ms / 1e3, 1),
            "kernel_ms": {kk: round(v, 4) for kk, v in k.items()}, "finite": bool(np.isfinite(rho).all())}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg1,cfg3,cfg4,cfg5")
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    only = set(args.only.split(","))
    from ionsolver_b200 import lbm as L
    V, F = L.VelocitySet, L.FloatType

    if "cfg1" in only:
        lbm = L.Lbm.setup_taylor_green(64, devices=[0])
        lbm.initialize()
        report("cfg1 64^3 D3Q19 FP32 Taylor-Green (setup.rs:92), no MHD", lbm, 1000, 100)
        lbm.close()
        lbm = L.Lbm.setup_lid_driven_cavity(64, devices=[0])
        lbm.initialize()
        report("cfg1 64^3 D3Q19 FP32 lid-driven cavity (equilibrium boundaries), no MHD", lbm, 1000, 100)
        lbm.close()

    if "cfg3" in only:
        cfg = L.LbmConfig(velocity_set=V.D3Q19, float_type=F.FP32, n_x=512, n_y=256, n_z=256, ext_volume_force=True, ext_magneto_hydro=True,
                          mhd_lod_depth=4, graphics_config=L.GraphicsConfig(False))
        cfg.units.set(256.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
        cfg.nu = cfg.units.nu_si_lu(1.48E-5)
        lbm = L.Lbm(cfg, devices=[0])
        t0 = time.perf_counter()
        lbm.import_mesh_reposition(os.path.join(STL, "ring_magnet.stl"), 128.1, 128.1, 128.0, 0.0, 0.0, 90.0, 100.0)
        lbm.import_mesh_reposition(os.path.join(STL, "disk_magnet.stl"), 400.1, 128.1, 128.0, 0.0, 0.0, 90.0, 80.0)
        lbm.import_mesh_reposition(os.path.join(STL, "tube.stl"), 256.1, 128.1, 128.0, 0.0, 0.0, 90.0, 200.0)
        lbm.voxelise_mesh(0, L.ModelType.Magnet, (1000000.0, 0.0, 0.0))
        lbm.voxelise_mesh(1, L.ModelType.Magnet, (500000.0, 0.0, 0.0))
        lbm.voxelise_mesh(2, L.ModelType.Solid)
        lbm.precompute_B()
        lbm.finish_queues()
        t_scene = time.perf_counter() - t0
        fill(lbm.domains[0], 11, 0.002)
        lbm.setup_velocity_field((0.05, 0.0, 0.0), 1.0)
        lbm.initialize()
        flags = lbm.domains[0].read(3)
        report("cfg3 512x256x256 D3Q19 FP32 MHD, ring + disk magnet + tube (synthetic STLs), LOD depth 4", lbm, 5, 2,
               {"scene_build_s": round(t_scene, 2), "solid_cells": int((flags & 1).sum()), "magnet_cells": int(((flags & 0x10) != 0).sum())})
        lbm.close()

    if "cfg4" in only:
        cfg = L.LbmConfig(velocity_set=V.D3Q27, float_type=F.FP16S, n_x=512, n_y=512, n_z=512, ext_volume_force=True, ext_magneto_hydro=True,
                          mhd_lod_depth=4, graphics_config=L.GraphicsConfig(False))
        cfg.units.set(512.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
        cfg.nu = 0.1
        lbm = L.Lbm(cfg, devices=[0])
        d = lbm.domains[0]
        fill(d, 11, 0.002)
        n = d.n
        fill(d, 6, 0.0)
        plane = np.full(1 << 24, 0.01, np.float32)  # uniform B_stat = (0, 0, 0.01): z plane of the 3-plane field
        for off in range(0, n, 1 << 24):
            k = min(1 << 24, n - off)
            d.write(6, plane[:k], (2 * n + off) * 4)
        lbm.setup_velocity_field((0.05, 0.01, 0.0), 1.0)
        lbm.initialize()
        report("cfg4 512^3 D3Q27 FP16S MHD, uniform B_stat, charged fluid, LOD depth 4 (1 GPU)", lbm, 5, 2)
        lbm.close()

    if "cfg5" in only:
        for nz, label in ((128, "0.54 G cells"),) + (((256, "1.07 G cells (fills 180 GB)"),) if args.full else ()):
            cfg = L.LbmConfig(velocity_set=V.D3Q19, float_type=F.FP16C, n_x=2048, n_y=2048, n_z=nz, ext_volume_force=True,
                              ext_magneto_hydro=True, mhd_lod_depth=4, graphics_config=L.GraphicsConfig(False))
            cfg.units.set(2048.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
            cfg.nu = 0.1
            try:
                lbm = L.Lbm(cfg, devices=[0])
            except Exception as e:  # out of memory is reported, not fatal
                print(json.dumps({"config": f"cfg5 2048x2048x{nz} D3Q19 FP16C MHD ({label})", "error": str(e)[:200]}), flush=True)
                continue
            d = lbm.domains[0]
            fill(d, 11, 0.002)
            fill(d, 2, 0.02)
            lbm.initialize()
            report(f"cfg5 2048x2048x{nz} D3Q19 FP16C MHD ({label}), LOD depth 4", lbm, 3, 1)
            lbm.close()


if __name__ == "__main__":
    main()
