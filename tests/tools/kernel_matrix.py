#!/usr/bin/env python
"""Per-kernel timings of the time-step kernels over the storage / velocity-set / extension matrix of BASELINE.json's
configs, on one GPU, with CUDA events on the domain's stream: one JSON line per configuration (MLUPs/s, achieved GB/s
against the algorithmic bytes of SURVEY 8d, fraction of the measured HBM copy peak).  Used for profiles/*_kernel_matrix.md.

usage: python tests/tools/kernel_matrix.py [--side 256] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--shapes", default="", help="extra D3Q19 FP32 MHD rows, e.g. 512x256x256,512x256x254")
    ap.add_argument("--lod-depth", type=int, default=3, help="mhd_lod_depth of the MHD rows (LOD deposit atomics)")
    ap.add_argument("--only", default="", help="run only the rows whose label contains this text (for ncu captures)")
    args = ap.parse_args()
    import torch
    from ionsolver_b200 import lbm as L
    peak = 6527.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    n = args.side
    V, F, R = L.VelocitySet, L.FloatType, L.RelaxationTime
    rows = [
        ("cfg1-like 64^3 D3Q19 FP32 SRT (fits L2)", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=64, n_y=64, n_z=64)),
        ("D3Q19 FP32 SRT", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=n, n_y=n, n_z=n)),
        ("D3Q19 FP32 TRT", dict(velocity_set=V.D3Q19, float_type=F.FP32, relaxation_time=R.Trt, n_x=n, n_y=n, n_z=n)),
        ("D3Q19 FP16S SRT", dict(velocity_set=V.D3Q19, float_type=F.FP16S, n_x=n, n_y=n, n_z=n)),
        ("D3Q19 FP16C SRT", dict(velocity_set=V.D3Q19, float_type=F.FP16C, n_x=n, n_y=n, n_z=n)),
        ("D3Q27 FP32 SRT", dict(velocity_set=V.D3Q27, float_type=F.FP32, n_x=n, n_y=n, n_z=n)),
        ("cfg2 D3Q19 FP32 MHD", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=n, n_y=n, n_z=n, mhd=True)),
        ("cfg3 shape 512x256x256 D3Q19 FP32 MHD", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=2 * n, n_y=n, n_z=n, mhd=True)),
        ("cfg4 kernel D3Q27 FP16S MHD", dict(velocity_set=V.D3Q27, float_type=F.FP16S, n_x=n, n_y=n, n_z=n, mhd=True)),
        ("cfg5 kernel D3Q19 FP16C MHD", dict(velocity_set=V.D3Q19, float_type=F.FP16C, n_x=n, n_y=n, n_z=n, mhd=True)),
    ]
    for shp in [x for x in args.shapes.split(",") if x]:
        sx, sy, sz = (int(v) for v in shp.split("x"))
        rows.append((f"shape {shp} D3Q19 FP32 MHD", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=sx, n_y=sy, n_z=sz, mhd=True)))
        rows.append((f"shape {shp} D3Q19 FP32 SRT", dict(velocity_set=V.D3Q19, float_type=F.FP32, n_x=sx, n_y=sy, n_z=sz)))
    for label, kw in rows:
        if args.only and args.only not in label:
            continue
        mhd = kw.pop("mhd", False)
        cfg = L.LbmConfig(nu=0.1, graphics_config=L.GraphicsConfig(False), ext_volume_force=mhd, ext_magneto_hydro=mhd, mhd_lod_depth=args.lod_depth, **kw)
        if mhd:
            cfg.units.set(float(n), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
        lbm = L.Lbm(cfg, devices=[0])
        d = lbm.domains[0]
        cells = d.n
        rng = np.random.default_rng(0)
        d.write(2, (0.05 * rng.standard_normal(3 * cells)).astype(np.float32))
        if mhd:
            d.write(11, np.full(cells, 0.002, np.float32))
        lbm.initialize()
        stream = torch.cuda.ExternalStream(d.stream(), device=0)
        q = {V.D2Q9: 9, V.D3Q15: 15, V.D3Q19: 19, V.D3Q27: 27}[cfg.velocity_set]
        s = 4 if cfg.float_type == F.FP32 else 2
        bpc = (1 + 4 * q * s + 14 * s + 24 + 4) if mhd else (1 + 2 * q * s)
        for t in range(3):
            d.enqueue_stream_collide(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lbm.finish_queues()
        e0.record(stream)
        for t in range(3, 3 + args.steps):
            if mhd:
                d.enqueue_clear_qu_lod()
            d.enqueue_stream_collide(t)
        e1.record(stream)
        lbm.finish_queues()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        gbs = cells * bpc / (ms * 1e-3) / 1e9
        print(json.dumps({"config": label, "cells": cells, "kernel": "stream_collide", "ms": round(ms, 4), "mlups": round(cells / ms / 1e3, 1),
                          "bytes_per_cell": bpc, "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3)}), flush=True)
        lbm.close()


if __name__ == "__main__":
    main()
