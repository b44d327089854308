#!/usr/bin/env python
"""Multi-GPU runs of the two scaling configurations of BASELINE.json (bench.py is cfg2), one process per GPU under torchrun:
   cfg4  512^3 D3Q27 FP16S MHD, z-slab split STRONG-scaled (the 512 z layers are divided over the ranks)
   cfg5  D3Q19 FP16C MHD WEAK-scaled, 2048 x 2048 x NZ cells per GPU (--nz, default 128 = 0.54 G cells; 256 fills 180 GB)
Charged fluid (Q = 0.002, uniform velocity) in a uniform static B field, LOD depth --lod-depth (default 4).  Timing as in bench.py:
barrier + synchronize around K steps, CUDA events on the domain stream, max over ranks; per-kernel times of the slowest rank.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/tools/bench_scaling.py --config cfg4"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def fill(d, field, value, plane=None, planes=1, chunk=1 << 26):
    """constant fill of one plane (or all) of a field without a domain-sized host array"""
    n = d.n
    buf = np.full(min(chunk, n), value, np.float32)
    for pl in (range(planes) if plane is None else [plane]):
        off = 0
        while off < n:
            k = min(chunk, n - off)
            d.write(field, buf[:k], (pl * n + off) * 4)
            off += k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nz", type=int, default=128, help="cfg5: z layers per GPU")
    ap.add_argument("--lod-depth", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    from ionsolver_b200 import capi, lbm as L
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL banner etc. to stderr: stdout carries one JSON line
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, F = L.VelocitySet, L.FloatType
    if args.config == "cfg4":
        vs, ft, nx, ny, nz_total, q, s, scaling = V.D3Q27, F.FP16S, 512, 512, 512, 27, 2, "strong"
        label = "cfg4: 512^3 D3Q27 FP16S MHD, z-slab split strong-scaled"
    else:
        vs, ft, nx, ny, nz_total, q, s, scaling = V.D3Q19, F.FP16C, 2048, 2048, args.nz * world, 19, 2, "weak"
        label = f"cfg5: D3Q19 FP16C MHD, 2048x2048x{args.nz} cells per GPU weak-scaled"
    cfg = L.LbmConfig(velocity_set=vs, float_type=ft, n_x=nx, n_y=ny, n_z=nz_total, d_z=world, ext_volume_force=True, ext_magneto_hydro=True,
                      mhd_lod_depth=args.lod_depth, graphics_config=L.GraphicsConfig(False))
    cfg.units.set(float(nx), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    cfg.nu = 0.1
    if world == 1:
        lbm = L.Lbm(cfg, devices=[local])
    else:
        ident = [L.Lbm.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        lbm = L.Lbm.new_distributed(cfg, rank, world, local, ident[0])
    d = lbm.domains[0]
    fill(d, 11, 0.002)                 # Q
    fill(d, 6, 0.01, plane=2)          # B_stat = (0, 0, 0.01)
    fill(d, 2, 0.05, plane=0)          # u = (0.05, 0.01, 0)
    fill(d, 2, 0.01, plane=1)
    lbm.initialize()
    stream = torch.cuda.ExternalStream(d.stream(), device=local)

    def barrier():
        lbm.finish_queues()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(vals):
        if dist is None:
            return vals
        t = torch.tensor(vals, device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    for _ in range(max(args.warmup, 3)):
        lbm.do_time_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        lbm.do_time_step()
    e1.record(stream)
    barrier()
    ms_step = maxr([e0.elapsed_time(e1) / args.steps])[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t = lbm.get_time_step()
    barrier()
    ev[0].record(stream)
    d.enqueue_clear_qu_lod()
    d.enqueue_stream_collide(t)
    ev[1].record(stream)
    d.enqueue_update_e_b_dyn()
    ev[2].record(stream)
    lbm.set_time_step(t + 1)
    barrier()
    sc_ms, eb_ms = maxr([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    cells_local = nx * ny * (nz_total // world)
    cells_global = nx * ny * nz_total
    bpc = 1 + 4 * q * s + 14 * s + 24 + 4
    peak = 6534.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    finite = bool(np.isfinite(d.read(1)[: 1 << 20]).all())
    os.dup2(saved, 1)
    if rank == 0:
        print(json.dumps({"config": label, "n_gpus": world, "scaling": scaling, "lattice": [nx, ny, nz_total], "cells_per_gpu": cells_local,
                          "lod_depth": args.lod_depth, "steps": args.steps, "ms_per_step": round(ms_step, 3), "mlups": round(cells_global / ms_step / 1e3, 1),
                          "kernel_ms_slowest_rank": {"stream_collide": round(sc_ms, 3), "update_e_b_dynamic": round(eb_ms, 3)},
                          "stream_collide_frac_of_hbm_peak": round(cells_local * bpc / (sc_ms * 1e-3) / 1e9 / peak, 3), "finite": finite}), flush=True)
    lbm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
