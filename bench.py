#!/usr/bin/env python
"""bench.py -- IonSolver extended-LBM MHD time step on B200: MLUPs/s, roofline and CPU baseline in one JSON line.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it is launched under
torchrun, one rank per GPU.  A "step" is one `Lbm::do_time_step` (clear_qu_lod + stream_collide + update_e_b_dynamic,
plus halo / LOD exchange when N > 1) over the whole lattice.

Workload (BASELINE.json configs[1]): 256^3 D3Q19 FP32 MHD, charged fluid (Q = 0.002/cell, u = (0.1, 0.01, 0), scene of
setup_bfield_spin, setup.rs:142-201) in the static field of a voxelised disk magnet (synthetic STL with the dimensions
of stl/disk-magnet.stl; voxelize_mesh + precompute_B, setup.rs:346-393), default mhd_lod_depth = 4.  N GPUs: weak
scaling, one 256x256x256 z-slab per GPU (d_z = N), halos and LOD pyramids exchanged with NCCL over NVLink.

MLUPs/s = lattice cells (halos excluded) x steps / seconds / 1e6 (src/info.rs:73-81).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_SIDE = 256
STL = os.path.join(ROOT, "tests", "golden", "stl", "disk_magnet.stl")
METRIC = "MLUPs/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own kernels on host cores
# ----------------------------------------------------------------------------------------------------------------
SAMPLE_SIDE = 128


def reference_sample_config(lod_depth=4):
    """Bounded sample of the workload for the CPU arms: same kernels, extensions, LOD depth (so the same 8^depth source
    terms per cell in update_e_b_dynamic) and units, on a 128^3 lattice instead of 256^3."""
    from oracle import ref_host as rh
    cfg = rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=SAMPLE_SIDE, n_y=SAMPLE_SIDE, n_z=SAMPLE_SIDE,
                       ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=lod_depth, graphics_active=False)
    cfg.units.set(float(N_SIDE), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    cfg.nu = float(cfg.units.nu_si_lu(1.48E-5))
    return cfg


def cpu_reference_run(steps, warmup, lod_depth, threads=0):
    """Times `steps` Lbm::do_time_step of the reference kernels (oracle/_ref: sim_kernels.cl compiled for the host; the
    C restatement if that library was not prebuilt) on the sample lattice.  Returns (MLUPs/s, ms/step, kind, cores)."""
    from oracle import build_ref, port, ref_host as rh
    cfg = reference_sample_config(lod_depth)
    geo = rh.domain_geometry(cfg, 0, 0, 0, 0)
    kind = "reference" if (os.path.isfile(build_ref.lib_path_for(rh.device_defines(cfg, geo))) or build_ref.reference_root()) else "port"
    cores = threads or port.max_threads()
    lbm = rh.RefLbm(cfg, threads=cores, backend="ref" if kind == "reference" else "port")
    d = lbm.domains[0]
    d.qc[:] = 0.002
    n = d.g.n
    d.u[:n] = 0.1
    d.u[n:2 * n] = 0.01
    d.b_stat[n:2 * n] = 1e-4  # stands in for the magnet field; static-field values do not change the per-cell work
    lbm.initialize()
    for _ in range(warmup):
        lbm.do_time_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        lbm.do_time_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n / dt / 1e6, dt * 1e3, kind, cores


def run_reference_arm(args, rank):
    if rank != 0:
        return
    mlups, ms, kind, cores = cpu_reference_run(args.steps, args.warmup, args.lod_depth)
    sample = f"{SAMPLE_SIDE}^3 lattice of the same scene (same kernels, LOD depth {args.lod_depth}), {args.steps} full time steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": mlups, "unit": "MLUPs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": mlups, "unit": "MLUPs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def workload_config(args):
    return {"workload": f"cfg2: {N_SIDE}^3 D3Q19 FP32 SRT MHD (volume_force + magneto_hydro), charged fluid Q=0.002 u=(0.1,0.01,0), "
                        f"static B of a voxelised disk magnet, mhd_lod_depth={args.lod_depth}",
            "lattice_per_gpu": [N_SIDE, N_SIDE, N_SIDE], "decomposition": f"d_z={args.gpus} z-slabs, one per GPU",
            "cache": "working set 7.3 GB per GPU >> 126 MB L2: every step streams from HBM, no L2 flush needed"}


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.stop, self.source = [], set(), None, threading.Event(), "nvidia-smi"
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop.wait(0.1)

    def _run_nvml(self):
        """Same quantities through NVML (nvidia_ml_py): a query takes microseconds, so a 0.5 s timed region gets ~50 samples instead
        of the one or two an nvidia-smi process manages.  Returns False (-> nvidia-smi loop) when NVML is not usable."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].strip().isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        self.source = "nvml"
        while not self.stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for b, n in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop.wait(0.01)
        return True

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "sm_mhz_min": float(np.min(self.samples)) if self.samples else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def build_scene(args, rank, world, device):
    """cfg2 through the public host API (LbmConfig/Lbm of mod.rs): returns an initialised Lbm."""
    from ionsolver_b200 import lbm as L
    cfg = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, relaxation_time=L.RelaxationTime.Srt, float_type=L.FloatType.FP32,
                      n_x=N_SIDE, n_y=N_SIDE, n_z=N_SIDE * world, d_z=world, ext_volume_force=True, ext_magneto_hydro=True,
                      mhd_lod_depth=args.lod_depth, graphics_config=L.GraphicsConfig(False))
    cfg.units.set(float(N_SIDE), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    cfg.nu = cfg.units.nu_si_lu(1.48E-5)
    if world == 1:
        lbm = L.Lbm(cfg, devices=[device])
    else:
        import torch
        import torch.distributed as dist
        ident = [L.Lbm.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        lbm = L.Lbm.new_distributed(cfg, rank, world, device, ident[0])
        torch.cuda.synchronize()
    lbm.import_mesh_reposition(STL, 0.5 * N_SIDE + 0.1, 0.5 * N_SIDE + 0.1, 0.5 * N_SIDE * world, 0.0, 0.0, 0.0, 0.5 * N_SIDE - 1.0)
    lbm.voxelise_mesh(0, L.ModelType.Magnet, (0.0, 1000000.0, 0.0))
    lbm.precompute_B()
    for d in lbm.domains:
        d.write(11, np.full(d.n, 0.002, np.float32))  # ION_FIELD_Q
    lbm.setup_velocity_field((0.1, 0.01, 0.0), 1.0)
    lbm.initialize()
    return lbm


def run_ours(args, rank, world, local_rank):
    import torch
    from ionsolver_b200 import capi
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; ionsolver_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    # NCCL prints its version banner to stdout; the driver wants exactly one JSON line there, so everything up to the final
    # print goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    lbm = build_scene(args, rank, world, device)
    dom = lbm.domains[0]
    stream = torch.cuda.ExternalStream(dom.stream(), device=device)
    cells_local = N_SIDE ** 3
    cells_global = cells_local * world

    def barrier():
        lbm.finish_queues()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=f"cuda:{device}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- whole-step throughput, inputs resident in HBM ----
    for _ in range(max(args.warmup, 3)):
        lbm.do_time_step()
    with ClockSampler(device) as clocks:
        l0 = capi.kernel_launch_count()
        ms_total = timed(lbm.do_time_step, args.steps)
        launches = capi.kernel_launch_count() - l0
    ms_step = ms_total / args.steps
    value = cells_global / (ms_step * 1e-3) / 1e6

    # ---- per-kernel durations inside the step (CUDA events on the launching stream) ----
    kern_ms = {"clear_qu_lod": 0.0, "stream_collide": 0.0, "update_e_b_dynamic": 0.0}
    k_prof = min(args.steps, 10)
    if True:  # every rank times its own kernels (no collective inside these three calls); rank 0 reports
        evs = []
        barrier()
        t = lbm.get_time_step()
        for s in range(k_prof):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(stream)
            dom.enqueue_clear_qu_lod()
            e[1].record(stream)
            dom.enqueue_stream_collide(t + s)
            e[2].record(stream)
            dom.enqueue_update_e_b_dyn()
            e[3].record(stream)
            evs.append(e)
        lbm.set_time_step(t + k_prof)
        barrier()
        for e in evs:
            kern_ms["clear_qu_lod"] += e[0].elapsed_time(e[1]) / k_prof
            kern_ms["stream_collide"] += e[1].elapsed_time(e[2]) / k_prof
            kern_ms["update_e_b_dynamic"] += e[2].elapsed_time(e[3]) / k_prof
    if dist is not None:  # slabs differ in how many foreign LOD sources they sum over: report the slowest rank's kernels
        tk = torch.tensor([kern_ms[k] for k in sorted(kern_ms)], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        kern_ms = {k: float(v) for k, v in zip(sorted(kern_ms), tk.tolist())}
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    sc_bytes = 1 + 4 * 19 * 4 + 14 * 4 + 24 + 4   # 389 B/cell: MHD stream_collide, D3Q19 FP32
    eb_bytes = 49                                 # update_e_b_dynamic
    pairs = cells_local * (8 ** args.lod_depth)   # (cell, LOD source) terms of the own pyramid; 9 FMA = 18 flop each
    kernels = {}
    roofline = None
    if True:
        fma_peak = capi.measure_fma_peak(device, packed=False)   # FMA/s, scalar FFMA, measured now on this GPU
        fma_peak_packed = capi.measure_fma_peak(device, packed=True)
        sc_gbs = cells_local * sc_bytes / (kern_ms["stream_collide"] * 1e-3) / 1e9
        eb_gbs = cells_local * eb_bytes / (kern_ms["update_e_b_dynamic"] * 1e-3) / 1e9
        eb_tflops = pairs * 18 / (kern_ms["update_e_b_dynamic"] * 1e-3) / 1e12
        kernels = {
            "stream_collide": {"ms": kern_ms["stream_collide"], "bound": "hbm", "bytes_per_cell": sc_bytes, "achieved_gbs": sc_gbs,
                               "frac_of_hbm_peak": sc_gbs / hbm_peak, "share_of_step": kern_ms["stream_collide"] / ms_step},
            "update_e_b_dynamic": {"ms": kern_ms["update_e_b_dynamic"], "bound": "fp32 FMA issue (8^depth source terms per cell, 9 FMA each)",
                                   "bytes_per_cell": eb_bytes, "achieved_gbs": eb_gbs, "frac_of_hbm_peak": eb_gbs / hbm_peak,
                                   "pairs_per_s": pairs / (kern_ms["update_e_b_dynamic"] * 1e-3), "achieved_tflops": eb_tflops,
                                   "fp32_peak_tflops": 2 * fma_peak / 1e12, "fp32_peak_tflops_packed": 2 * fma_peak_packed / 1e12,
                                   "frac_of_fp32_peak": eb_tflops / (2 * fma_peak / 1e12),
                                   "share_of_step": kern_ms["update_e_b_dynamic"] / ms_step,
                                   "pairs_note": "pairs = cells x 8^depth, the reference's loop count (sim.cl:943); LOD rows whose entries "
                                                 "are all zero are skipped by the kernel (in a single-domain run 585 of the 4096 window slots "
                                                 "are never filled, quirk Q5), so the executed FMA rate is ~14 % below achieved_tflops there"},
            "clear_qu_lod": {"ms": kern_ms["clear_qu_lod"], "share_of_step": kern_ms["clear_qu_lod"] / ms_step},
        }
        # The roofline object is for the dominant kernel of the step.  When that is update_e_b_dynamic (LOD depth 4) the HBM
        # fraction is tiny by construction -- the kernel is bound by CUDA-core FP32 issue, neither by HBM nor by tensor cores
        #  -- so its FP32 roofline is attached under "compute", and stream_collide's HBM roofline under "hbm_kernel".
        dom_name = max(("stream_collide", "update_e_b_dynamic"), key=lambda k: kernels[k]["ms"])
        hbm_kernel = {"kernel": "stream_collide", "bound": "hbm", "achieved": sc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": sc_gbs / hbm_peak,
                      "algorithmic_bytes_per_launch": cells_local * sc_bytes, "ms_per_launch": kern_ms["stream_collide"],
                      "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "traffic": 6.477e9,
                      "traffic_source": "profiles/r1_ncu_stream_collide.md (ncu --set full): dram__bytes_read.sum 3.442 GB + dram__bytes_write.sum 3.035 GB per launch"}
        if dom_name == "update_e_b_dynamic":
            # The dominant kernel of the step at this LOD depth is bound by CUDA-core FP32 issue -- neither by HBM (49 B per cell
            # against 8^depth x 18 flop) nor by tensor cores -- so its roofline is stated in FP32 TFLOP/s against
            # the FFMA issue peak measured in this process; the HBM-bound kernel of the step follows under "hbm_kernel".
            roofline = {"kernel": dom_name, "bound": "fp32", "achieved": eb_tflops, "peak": 2 * fma_peak / 1e12, "unit": "TFLOP/s",
                        "frac": eb_tflops / (2 * fma_peak / 1e12), "flop_per_launch": pairs * 18, "ms_per_launch": kern_ms["update_e_b_dynamic"],
                        "peak_source": "scalar FFMA issue micro-benchmark run in this process on this GPU (ion_measure_fma_peak); "
                                       "MEASURED_PEAKS.json has no FP32 entry",
                        "traffic": 0.792e9, "traffic_source": "profiles/r1_ncu_update_e_b_pair.md: dram__bytes_read.sum + dram__bytes_write.sum",
                        "hbm_view": {"algorithmic_bytes_per_launch": cells_local * eb_bytes, "achieved_gbs": eb_gbs, "frac_of_hbm_peak": eb_gbs / hbm_peak},
                        "hbm_kernel": hbm_kernel,
                        "note": "bound is 'fp32' (CUDA cores), outside the hbm|tensor pair: see DESIGN.md 4.2 for why; 'hbm_kernel' is stream_collide"}
        else:
            roofline = dict(hbm_kernel)

    # ---- end to end through the public API with HOST buffers: load state -> initialize -> step -> save state ----
    e2e = None
    if True:  # at N GPUs every rank uploads / downloads the sections of its own slab; initialize and the step exchange halos
        n = dom.n
        host = {name: torch.empty(sz, dtype=dt, pin_memory=True) for name, sz, dt in
                (("flags", n, torch.uint8), ("rho", n, torch.float32), ("u", 3 * n, torch.float32), ("q", n, torch.float32))}
        ids = {"flags": 3, "rho": 1, "u": 2, "q": 11}
        for name, tns in host.items():
            tns.numpy()[:] = dom.read(ids[name])
        lib = capi.load()

        def io(fn):
            for name, tns in host.items():
                capi.check(fn(dom.handle, ids[name], ctypes.c_void_p(tns.data_ptr()), 0, tns.numel() * tns.element_size()))

        def e2e_step():
            io(lib.ion_buffer_write)          # the .ion sections a loader uploads (file.rs:118-152), from pinned memory
            lbm.initialize()                  # main.rs:274-278: a loaded state is re-initialised
            lbm.do_time_step()
            lbm.finish_queues()
            io(lib.ion_buffer_read)           # the sections a writer downloads (file.rs:221-268)

        e2e_step()
        k_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        barrier()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if dist is not None:
            tt = torch.tensor([dt], device=f"cuda:{device}", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        nbytes = sum(t.numel() * t.element_size() for t in host.values()) * world
        e2e = {"value": cells_global / dt / 1e6, "unit": "MLUPs/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "what": "per step: upload flags/rho/u/Q from pinned host memory, Lbm::initialize, "
               "Lbm::do_time_step, download flags/rho/u/Q (the load -> step -> save cycle of file.rs through the C ABI)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mlups, ms, kind, cores = cpu_reference_run(2, 1, args.lod_depth)
        cpu = {"value": mlups, "unit": "MLUPs/s", "cores": cores, "kind": kind, "ms_per_step": ms,
               "sample": f"{SAMPLE_SIDE}^3 lattice of the same scene (same kernels, LOD depth {args.lod_depth}), 2 timed full time steps after 1 warm-up"}

    # Source terms per cell that update_e_b_dynamic sums (sim.cl:940-983).  They GROW with the number of domains -- the own window
    # is fully populated only in multi-domain runs (quirk Q5 leaves 585 of its 4096 slots empty in a single domain, and all-zero
    # rows are skipped) and every other slab adds its pyramid level max(depth - distance, 0) -- so MLUPs/s per GPU falls with N at
    # constant work per SOURCE TERM; this is the reference's algorithm, not communication (halos and LODs overlap with the update).
    D = args.lod_depth
    own_fine = 8 ** D
    if world == 1:
        empty = sum(8 ** i for i in range(D))
        terms = {"own": own_fine - (empty // (2 ** D)) * (2 ** D), "foreign_max": 0}
    else:
        # only slabs with a LOWER index count: for the others the reference's `domain_diff * DEF_N` is a negative int times a uint,
        # which wraps to ~4.29e9 cells (quirk Q18) -- their terms are < 1e-16 of the sums and the fast path skips them
        terms = {"own": own_fine, "foreign_max": max(sum(8 ** max(D - (r - o), 0) for o in range(r)) for r in range(world))}
    terms["per_cell_slowest_rank"] = terms["own"] + terms["foreign_max"]
    terms["note"] = ("update_e_b_dynamic work per cell depends on the domain count (reference algorithm: LOD window quirk Q5 + foreign pyramids); "
                     "compare runs by source terms per second, not only by MLUPs/s")
    terms["source_terms_per_s_per_gpu"] = terms["per_cell_slowest_rank"] * cells_local / (kern_ms["update_e_b_dynamic"] * 1e-3)

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MLUPs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernels": kernels, "lod_source_terms": terms, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    lbm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lod-depth", type=int, default=4, help="mhd_lod_depth (reference default 4, mod.rs:126)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
