#!/usr/bin/env python
"""bench.py -- IonSolver extended-LBM MHD time step on B200: MLUPs/s, roofline, parity and CPU baseline in one JSON line.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference] [--config cfgX] [--scaling weak|strong]`;
for N > 1 it is launched under torchrun, one rank per GPU.  A "step" is one `Lbm::do_time_step` (mod.rs:250-272: clear_qu_lod +
stream_collide + update_e_b_dynamic, plus halo / LOD exchange when N > 1) over the whole lattice.

Workloads (BASELINE.json `configs`):
  cfg1  64^3 D3Q19 FP32 Taylor-Green (setup_taylor_green, setup.rs:92-113), no MHD; working set 20 MB, L2 resident
  cfg2  256^3 D3Q19 FP32 MHD (DEFAULT; the configuration `metric` is quoted on): charged fluid Q = 0.002/cell, u = (0.1, 0.01, 0)
        (setup_bfield_spin, setup.rs:142-201) in the static field of the reference's stl/disk-magnet.stl, voxelised and turned into
        B_stat by precompute_B (setup_mesh_field_test, setup.rs:346-393), default mhd_lod_depth = 4
  cfg3  setup_deeva_test (setup.rs:395-453) with every length doubled: 256 x 512 x 256 (the 512x256x256 of BASELINE.json in the
        scene's own axis order), six deeva_* STLs, precompute_B + static E of the plates, ext_subgrid_ecr off
  cfg4  512^3 D3Q27 FP16S MHD, z slabs, strong scaling
  cfg5  2048 x 2048 x 256 per GPU, D3Q19 FP16C MHD, z slabs, weak scaling (~1.07 G cells / GPU)
N GPUs: cfg1/2/3/5 weak (one lattice per GPU, d_z = N), cfg4 strong unless --scaling says otherwise; halos and LOD pyramids over
NCCL / NVLink.

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

REF_STL = os.path.join(ROOT, "tests", "golden", "stl", "ref")  # the reference's stl/*.stl, kept as test fixtures
METRIC = "MLUPs/s"

# per-GPU lattice (weak) or global lattice (strong), kernel family, default scaling
CONFIGS = {
    "cfg1": {"n": (64, 64, 64), "vs": "D3Q19", "ft": "FP32", "mhd": False, "scaling": "weak",
             "what": "64^3 D3Q19 FP32 SRT Taylor-Green (setup_taylor_green), no extensions"},
    "cfg2": {"n": (256, 256, 256), "vs": "D3Q19", "ft": "FP32", "mhd": True, "scaling": "weak",
             "what": "256^3 D3Q19 FP32 SRT MHD (volume_force + magneto_hydro), charged fluid Q=0.002 u=(0.1,0.01,0), static B of the "
                     "voxelised stl/disk-magnet.stl (precompute_B)"},
    "cfg3": {"n": (256, 512, 256), "vs": "D3Q19", "ft": "FP32", "mhd": True, "scaling": "weak",
             "what": "setup_deeva_test x2: 256x512x256 D3Q19 FP32 MHD, six deeva_* STLs (ring + disk magnet, quartz tube, inlet, e-plates), "
                     "precompute_B + static E of the plates, ext_subgrid_ecr off"},
    "cfg4": {"n": (512, 512, 512), "vs": "D3Q27", "ft": "FP16S", "mhd": True, "scaling": "strong",
             "what": "512^3 D3Q27 (canonical weights) FP16S MHD, charged fluid, uniform B_stat = (0,0,0.01), z slabs"},
    "cfg5": {"n": (2048, 2048, 256), "vs": "D3Q19", "ft": "FP16C", "mhd": True, "scaling": "weak",
             "what": "2048x2048x256 per GPU D3Q19 FP16C MHD, charged fluid, uniform B_stat = (0,0,0.01), z slabs"},
}
QSET = {"D2Q9": 9, "D3Q15": 15, "D3Q19": 19, "D3Q27": 27}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic():
    """DRAM bytes per launch of the step's kernels from the committed ncu captures of THIS command (profiles/r2_traffic.json,
    written by tests/tools/make_profiles.py from `ncu --set full`); None when no capture matches the configuration."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return {}


def bytes_per_cell(cfg):
    """Algorithmic bytes per cell update: stream_collide and update_e_b_dynamic."""
    q, s = QSET[cfg["vs"]], (4 if cfg["ft"] == "FP32" else 2)
    sc = 1 + 4 * q * s + 14 * s + 24 + 4 if cfg["mhd"] else 1 + 2 * q * s
    return sc, (49 if cfg["mhd"] else 0)


def host_threads():
    """All host cores this process may use -- not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(device):
    """Pin this rank's thread (and with it the pages of every pinned staging buffer it allocates afterwards: first touch) to the
    CPUs of the NUMA node its GPU hangs off -- host <-> device copies of eight ranks otherwise cross the socket interconnect at
    random.  Returns (description for the JSON line, the previous affinity mask or None)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "bound": False, "why": "the platform reports no NUMA node for the GPU"}, None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        before = os.sched_getaffinity(0)
        allowed = before & cpus
        if not allowed:
            return {"numa_node": node, "bound": False, "why": "none of the node's CPUs is in this process's cpuset"}, None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": True, "cpus": len(allowed), "cpus_before": len(before)}, before
    except Exception as e:  # a sandbox without sysfs: leave the scheduler alone
        return {"numa_node": None, "bound": False, "why": f"{type(e).__name__}: {e}"}, None


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own kernels on host cores, on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------------------
def sample_lattice(cfg):
    """Bounded sample for the CPU arms: same kernels (velocity set, storage, extensions, LOD depth -> the same 8^depth source
    terms per cell in update_e_b_dynamic) on a lattice the host finishes in seconds per step."""
    if not cfg["mhd"]:
        return (64, 64, 64)
    return (128, 128, 128) if cfg["vs"] != "D3Q27" else (96, 96, 96)


def reference_sample_config(cfg, lod_depth):
    from oracle import ref_host as rh
    n = sample_lattice(cfg)
    c = rh.RefConfig(velocity_set=cfg["vs"], float_type=cfg["ft"], n_x=n[0], n_y=n[1], n_z=n[2], ext_volume_force=cfg["mhd"],
                     ext_magneto_hydro=cfg["mhd"], mhd_lod_depth=lod_depth, graphics_active=False)
    if cfg["mhd"]:
        c.units.set(float(cfg["n"][0]), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
        c.nu = float(c.units.nu_si_lu(1.48E-5))
    else:
        c.nu = float(c.units.nu_si_lu(0.1))
    return c


def sample_description(cfg, lod_depth):
    n = sample_lattice(cfg)
    if not cfg["mhd"]:
        return f"{n[0]}x{n[1]}x{n[2]} {cfg['vs']} {cfg['ft']} Taylor-Green (the full cfg1 lattice)"
    return (f"{n[0]}x{n[1]}x{n[2]} {cfg['vs']} {cfg['ft']} MHD sample of the workload: same kernels, extensions, units and LOD depth {lod_depth}, "
            f"charged fluid Q=0.002 u=(0.1,0.01,0), uniform B_stat=(0,1e-4,0) in place of the voxelised magnet (static-field values do not change "
            f"the per-cell work)")


def fill_sample(lbm_like, is_oracle, cfg):
    """Identical initial state on the oracle's numpy buffers or on the product's device buffers."""
    for d in lbm_like.domains:
        n = d.g.n if is_oracle else d.n
        if not cfg["mhd"]:  # cfg1: Taylor-Green vortices of setup.rs:458-543 (quirk Q11 included), identical arrays on both sides
            import cases
            u, rho = cases.taylor_green_numpy(sample_lattice(cfg)[0])
            if is_oracle:
                d.u[:] = u
                d.rho[:] = rho
            else:
                d.write(2, u)
                d.write(1, rho)
            continue
        q = np.full(n, 0.002, np.float32)
        u = np.zeros(3 * n, np.float32)
        u[:n] = 0.1
        u[n:2 * n] = 0.01
        b = np.zeros(3 * n, np.float32)
        b[n:2 * n] = 1e-4
        if is_oracle:
            d.qc[:] = q
            d.u[:] = u
            d.b_stat[:] = b
        else:
            d.write(11, q)  # ION_FIELD_Q
            d.write(2, u)   # ION_FIELD_U
            d.write(6, b)   # ION_FIELD_B_STAT


def cpu_reference_run(cfg, steps, warmup, lod_depth, keep_state=False):
    """Times `steps` Lbm::do_time_step of the reference kernels (oracle/_ref: sim_kernels.cl compiled for the host; the C
    restatement if that library was not prebuilt) on the sample lattice.  Returns a dict; with keep_state the oracle object after
    initialize + ONE step is returned too (bench parity leg)."""
    from oracle import build_ref, ref_host as rh
    c = reference_sample_config(cfg, lod_depth)
    geo = rh.domain_geometry(c, 0, 0, 0, 0)
    kind = "reference" if (os.path.isfile(build_ref.lib_path_for(rh.device_defines(c, geo))) or build_ref.reference_root()) else "port"
    cores = host_threads()
    lbm = rh.RefLbm(c, threads=cores, backend="ref" if kind == "reference" else "port")
    fill_sample(lbm, True, cfg)
    lbm.initialize()
    state = None
    t_first = time.perf_counter()
    lbm.do_time_step()
    t_first = time.perf_counter() - t_first
    if keep_state:
        d = lbm.domains[0]
        # rho / u are only written by the step when graphics are active (UPDATE_FIELDS, domain.rs:856) -- the benchmark runs with
        # graphics off like a production run, so the DDFs carry the flow state; Q, E, B are written every step
        state = {k: np.array(getattr(d, k), copy=True) for k in (("fi", "qc", "e_dyn", "b_dyn") if cfg["mhd"] else ("fi",))}
    for _ in range(max(warmup - 1, 0)):
        lbm.do_time_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        lbm.do_time_step()
    dt = (time.perf_counter() - t0) / max(steps, 1) if steps else t_first
    n = lbm.domains[0].g.n
    return {"mlups": n / dt / 1e6, "ms": dt * 1e3, "kind": kind, "cores": cores, "state": state, "config": c}


def run_reference_arm(args, rank, world):
    if rank != 0:  # under torchrun rank 0 alone runs the CPU arm
        return
    cfg = CONFIGS[args.config]
    r = cpu_reference_run(cfg, args.steps, args.warmup, args.lod_depth)
    sample = f"{sample_description(cfg, args.lod_depth)}; {args.steps} full time steps after {args.warmup} warm-up"
    conf = workload_config(args, world)
    conf["reference_arm_runs"] = sample  # the CPU arm times a bounded sample of the workload, not the full lattice
    conf["same_lattice_as_gpu_arm"] = False
    line = {
        "impl": "reference", "metric": METRIC, "value": r["mlups"], "unit": "MLUPs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": conf,
        "cpu_baseline": {"value": r["mlups"], "unit": "MLUPs/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["mlups"], "unit": "MLUPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def scaling_of(args):
    return args.scaling or CONFIGS[args.config]["scaling"]


def lattice_of(args, world):
    """(global lattice, per-GPU lattice) for the run."""
    n = CONFIGS[args.config]["n"]
    if args.cells_z:
        n = (n[0], n[1], args.cells_z)
    if scaling_of(args) == "weak":
        return (n[0], n[1], n[2] * world), n
    return n, (n[0], n[1], n[2] // world)


def workload_config(args, world):
    cfg = CONFIGS[args.config]
    g, l = lattice_of(args, world)
    ws = float(np.prod(l)) * (sum(bytes_per_cell(cfg)) + (60 if cfg["mhd"] else 16)) / 1e9
    return {"workload": f"{args.config}: {cfg['what']}" + (f", mhd_lod_depth={args.lod_depth}" if cfg["mhd"] else ""),
            "lattice_global": list(g), "lattice_per_gpu": list(l), "decomposition": f"d_z={world} z-slabs, one per GPU",
            "cache": (f"working set ~{ws:.1f} GB per GPU >> 126 MB L2: every step streams from HBM, no L2 flush needed" if ws > 1.0 else
                      f"working set ~{ws * 1e3:.0f} MB per GPU: L2 resident (reported, not used for the HBM claim)")}


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.stop, self.source = [], set(), None, threading.Event(), "nvidia-smi"
        self.index = index
        self.window = [None, None]  # perf_counter bounds of the timed region; the sampler itself starts before the warm-up
        self.thread = threading.Thread(target=self._run, daemon=True)

    def wait_ready(self, timeout=3.0):
        """Blocks until the sampler delivers (NVML initialisation takes ~100 ms): a timed region of a few milliseconds is still sampled."""
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.002)

    def begin(self):
        self.window[0] = time.perf_counter()

    def end(self):
        self.window[1] = time.perf_counter()

    def _record(self, mhz, reasons):
        now = time.perf_counter()
        self.samples.append((now, mhz, tuple(reasons)))

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
                self.max_mhz = float(out[1])
                self._record(float(out[0]), [n for n, v in zip(names, out[2:]) if v.strip().lower().startswith("active")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def _run_nvml(self):
        """Same quantities through NVML (nvidia_ml_py): a query takes microseconds, so a short timed region still gets tens of
        samples.  Returns False (-> nvidia-smi loop) when NVML is not usable."""
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
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                self._record(mhz, [n for b, n in bits.items() if r & b])
            except Exception:
                pass
            self.stop.wait(0.002)
        return True

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        t0, t1 = self.window
        inside = [s for s in self.samples if t0 is not None and t1 is not None and t0 <= s[0] <= t1]
        use = inside or self.samples  # a timed region shorter than one sampling period: fall back to everything sampled under load
        mhz = [s[1] for s in use]
        reasons = sorted({r for s in use for r in s[2]})
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz, "sm_mhz_min": float(np.min(mhz)) if mhz else None,
                "reasons": reasons, "samples": len(use), "samples_inside_timed_region": len(inside), "source": self.source}


def make_lbm(L, cfg_l, rank, world, device):
    if world == 1:
        return L.Lbm(cfg_l, devices=[device])
    import torch
    import torch.distributed as dist
    ident = [L.Lbm.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    lbm = L.Lbm.new_distributed(cfg_l, rank, world, device, ident[0])
    torch.cuda.synchronize()
    return lbm


def build_scene(args, rank, world, device):
    """The workload through the public host API (LbmConfig / Lbm of mod.rs, scene code of setup.rs): an initialised Lbm."""
    from ionsolver_b200 import lbm as L
    cfg = CONFIGS[args.config]
    g, _ = lattice_of(args, world)
    vs = {"D3Q19": L.VelocitySet.D3Q19, "D3Q27": L.VelocitySet.D3Q27}[cfg["vs"]]
    ft = {"FP32": L.FloatType.FP32, "FP16S": L.FloatType.FP16S, "FP16C": L.FloatType.FP16C}[cfg["ft"]]
    c = L.LbmConfig(velocity_set=vs, relaxation_time=L.RelaxationTime.Srt, float_type=ft, n_x=g[0], n_y=g[1], n_z=g[2], d_z=world,
                    ext_volume_force=cfg["mhd"], ext_magneto_hydro=cfg["mhd"], mhd_lod_depth=args.lod_depth,
                    graphics_config=L.GraphicsConfig(False))
    if args.config == "cfg1":
        c.nu = c.units.nu_si_lu(0.1)
        lbm = make_lbm(L, c, rank, world, device)
        lbm.set_taylor_green(1)
        lbm.initialize()
        return lbm
    if args.config == "cfg3":
        c.units.set(256.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 10e-8, 1.0, 50000.0)  # setup.rs:397 with lbm_length x2
        c.nu = c.units.nu_si_lu(0.05)
        c.ecr_freq = c.units.time_lu_si(2.45E9)
        lbm = make_lbm(L, c, rank, world, device)
        s = 2.0
        zc = 64.0 * s * world if scaling_of(args) == "weak" else 64.0 * s  # the thruster sits in the middle of the global lattice
        parts = [("deeva_disk_magnet.stl", 64.001 * s, 0.0, L.ModelType.Magnet, (0.0, 1000000.0, 0.0)),
                 ("deeva_inlet.stl", 64.0 * s, 0.0, L.ModelType.Solid, None),
                 ("deeva_quartz_tube.stl", 64.001 * s, 0.0, L.ModelType.Solid, None),
                 ("deeva_ring_magnet.stl", 64.001 * s, -0.5 * s, L.ModelType.Magnet, (0.0, 500000.0, 0.0)),
                 ("deeva_e_plate1.stl", 64.0 * s, 0.0, L.ModelType.ChargedECR, 0.00000000000021844213 / 2.0),
                 ("deeva_e_plate2.stl", 64.0 * s, 0.0, L.ModelType.ChargedECR, -0.00000000000021844213 / 2.0)]
        for i, (f, ox, oy, kind, val) in enumerate(parts):
            lbm.import_mesh(os.path.join(REF_STL, f), 1.0, ox, oy, zc, 0.0, 0.0, 0.0)
            lbm.voxelise_mesh(i, kind, val)
        for d in lbm.domains:  # 5.4 M magnet cells x 33.5 M outputs = 100 s (fast direct sum) or 163 s (reference order): FFT convolution, 0.4 s
            d.set_precompute_mode(args.precompute_mode if args.precompute_mode >= 0 else 2)
        lbm.precompute_B()
        lbm.precompute_E()
        for d in lbm.domains:
            d.write(11, np.full(d.n, 0.002, np.float32))
        lbm.setup_velocity_field((0.0, 0.05, 0.0), 1.0)  # propellant flowing along the tube axis
        lbm.initialize()
        return lbm
    # cfg2 / cfg4 / cfg5: charged fluid of setup_bfield_spin (setup.rs:144,182-197)
    c.units.set(float(cfg["n"][0]), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
    c.nu = c.units.nu_si_lu(1.48E-5)
    lbm = make_lbm(L, c, rank, world, device)
    if args.config == "cfg2":
        n = cfg["n"][0]
        lbm.import_mesh_reposition(os.path.join(REF_STL, "disk-magnet.stl"), 0.5 * n + 0.1, 0.5 * n + 0.1, 0.5 * g[2], 0.0, 0.0, 0.0, 0.5 * n - 1.0)
        lbm.voxelise_mesh(0, L.ModelType.Magnet, (0.0, 1000000.0, 0.0))
        if args.precompute_mode >= 0:
            for d in lbm.domains:
                d.set_precompute_mode(args.precompute_mode)
        t_pre = time.perf_counter()
        lbm.precompute_B()
        print(f"precompute_B (mode {max(args.precompute_mode, 0)}): {time.perf_counter() - t_pre:.3f} s", file=sys.stderr)
    else:
        for d in lbm.domains:  # uniform B_stat = (0, 0, 0.01) LU written like a scene would (setup.rs:182-190)
            b = np.zeros(3 * d.n, np.float32)
            b[2 * d.n:] = 0.01
            d.write(6, b)
            del b
    for d in lbm.domains:
        d.write(11, np.full(d.n, 0.002, np.float32))  # ION_FIELD_Q
    lbm.setup_velocity_field((0.1, 0.01, 0.0), 1.0)
    lbm.initialize()
    return lbm


def gpu_sample_parity(args, device, cpu_state, ref_cfg):
    """Parity of the timed path against the reference at the CPU arm's sample size: the SAME sample scene is built on the GPU,
    both sides run initialize + one time step from identical state, and the fields are compared (relative L2)."""
    from ionsolver_b200 import lbm as L
    import cases
    from oracle_util import rel_l2
    cfg = CONFIGS[args.config]
    g = L.Lbm(cases.to_lbm_config(ref_cfg), devices=[device])
    fill_sample(g, False, cfg)
    g.initialize()
    g.do_time_step()
    g.finish_queues()
    d = g.domains[0]
    ids = {"fi": 0, "rho": 1, "u": 2, "qc": 11, "e_dyn": 7, "b_dyn": 8}
    out = {}
    for k, want in cpu_state.items():
        got = np.asarray(d.read(ids[k]))
        got = got.view(want.dtype) if got.dtype != want.dtype else got
        out[k] = {"bit_exact": bool(got.tobytes() == want.tobytes())}
        if want.dtype == np.float32:
            out[k]["rel_l2"] = float(rel_l2(got, want))
        else:
            out[k]["stored_words_differing"] = int((got != want).sum())
    fft = d.eb_fft_info() if cfg["mhd"] else (0, 0)
    g.close()
    return {"what": "GPU default path vs the reference kernels on the CPU: same sample scene, initialize + 1 step from identical state",
            "lattice": [ref_cfg.n_x, ref_cfg.n_y, ref_cfg.n_z], "fields": out, "polyphase_fft_tasks": int(fft[1])}


def run_ours(args, rank, world, local_rank):
    import torch
    from ionsolver_b200 import capi
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; ionsolver_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    cfg = CONFIGS[args.config]
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
    numa, affinity_before = bind_to_gpu_numa_node(device)
    t_build = time.perf_counter()
    lbm = build_scene(args, rank, world, device)
    lbm.finish_queues()
    t_build = time.perf_counter() - t_build
    dom = lbm.domains[0]
    stream = torch.cuda.ExternalStream(dom.stream(), device=device)
    g_lat, l_lat = lattice_of(args, world)
    cells_local = int(np.prod(l_lat))
    cells_global = int(np.prod(g_lat))
    mhd = cfg["mhd"]

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
    with ClockSampler(device) as clocks:  # started before the warm-up so that NVML is initialised when the timed region begins
        clocks.wait_ready()
        for _ in range(max(args.warmup, 3)):
            lbm.do_time_step()
        barrier()
        l0 = capi.kernel_launch_count()
        clocks.begin()
        ms_total = timed(lbm.do_time_step, args.steps)
        clocks.end()
        launches = capi.kernel_launch_count() - l0
    ms_step = ms_total / args.steps
    value = cells_global / (ms_step * 1e-3) / 1e6

    # ---- per-kernel durations inside the step (CUDA events on the launching stream) ----
    names = ["clear_qu_lod", "stream_collide", "lod_fold", "update_e_b_dynamic"] if mhd else ["stream_collide"]
    kern_ms = {k: 0.0 for k in names}
    k_prof = min(args.steps, 10)
    evs = []
    barrier()
    t = lbm.get_time_step()
    for s in range(k_prof):  # every rank times its own kernels (no collective inside these calls)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record(stream)
        if mhd:
            dom.enqueue_clear_qu_lod()
        e[1].record(stream)
        dom.enqueue_stream_collide_range(t + s, 0, dom.n_z, False)   # the kernel alone ...
        e[2].record(stream)
        dom.enqueue_stream_collide_range(t + s, 0, 0, True)           # ... and the fold of its LOD deposits (k_lod_fold)
        e[3].record(stream)
        if mhd:
            dom.enqueue_update_e_b_dyn()
        e[4].record(stream)
        evs.append(e)
    lbm.set_time_step(t + k_prof)
    barrier()
    for e in evs:
        if mhd:
            kern_ms["clear_qu_lod"] += e[0].elapsed_time(e[1]) / k_prof
            kern_ms["lod_fold"] += e[2].elapsed_time(e[3]) / k_prof
            kern_ms["update_e_b_dynamic"] += e[3].elapsed_time(e[4]) / k_prof
        kern_ms["stream_collide"] += e[1].elapsed_time(e[2]) / k_prof
    if dist is not None:  # slabs differ in how many foreign LOD sources they sum over: report the slowest rank's kernels
        tk = torch.tensor([kern_ms[k] for k in sorted(kern_ms)], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        kern_ms = {k: float(v) for k, v in zip(sorted(kern_ms), tk.tolist())}
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    sc_bytes, eb_bytes = bytes_per_cell(cfg)
    cells_halo = dom.n  # the kernels also stream the two halo layers of a slab; algorithmic bytes are counted on lattice cells only
    traffic = ncu_traffic().get(args.config, {}) if world == 1 else {}
    fft_bytes, fft_tasks = dom.eb_fft_info() if mhd else (0, 0)

    def hbm_view(name, bpc, extra=None):
        ms = kern_ms[name]
        gbs = cells_local * bpc / (ms * 1e-3) / 1e9
        o = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
             "algorithmic_bytes_per_cell": bpc, "algorithmic_bytes_per_launch": cells_local * bpc, "ms_per_launch": ms,
             "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind}; burst copy figure)",
             "traffic": traffic.get(name, {}).get("dram_bytes"), "traffic_source": traffic.get(name, {}).get("source")}
        if extra:
            o.update(extra)
        return o

    kernels = {"stream_collide": hbm_view("stream_collide", sc_bytes)}
    kernels["stream_collide"]["share_of_step"] = kern_ms["stream_collide"] / ms_step
    if mhd:
        eb = hbm_view("update_e_b_dynamic", eb_bytes, {
            "algorithm": ("polyphase FFT convolution (eb_fft.cu): k_eb_src + k_eb_fft, one block per in-block offset"
                          if fft_tasks else "direct summation over the LOD sources (fields.cu)"),
            "polyphase_fft_tasks": int(fft_tasks), "static_kernel_spectra_bytes": int(fft_bytes),
            "streamed_bytes_per_launch": cells_local * eb_bytes + int(fft_bytes),
            "streamed_frac_of_hbm_peak": (cells_local * eb_bytes + int(fft_bytes)) / (kern_ms["update_e_b_dynamic"] * 1e-3) / 1e9 / hbm_peak,
            "equivalent_direct_pair_terms_per_s": cells_local * (8 ** args.lod_depth) / (kern_ms["update_e_b_dynamic"] * 1e-3),
            "note": "`achieved` counts SURVEY 8d's 49 B per cell only; the kernel also streams its static kernel spectra (102 B per cell) once "
                    "per step -- `streamed_*` includes them.  The reference algorithm sums 8^depth source terms per cell here."})
        eb["share_of_step"] = kern_ms["update_e_b_dynamic"] / ms_step
        kernels["update_e_b_dynamic"] = eb
        kernels["clear_qu_lod"] = {"ms": kern_ms["clear_qu_lod"], "share_of_step": kern_ms["clear_qu_lod"] / ms_step}
        kernels["lod_fold"] = {"ms": kern_ms["lod_fold"], "share_of_step": kern_ms["lod_fold"] / ms_step,
                               "what": "k_lod_fold: replicas of the finest LOD level -> QU_lod, launched by ion_enqueue_stream_collide after the kernel"}
    dom_name = max((k for k in kernels if "achieved" in kernels[k]), key=lambda k: kernels[k]["ms_per_launch"])
    roofline = {k: v for k, v in kernels[dom_name].items() if k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source",
                                                                     "algorithmic_bytes_per_launch", "ms_per_launch", "peak_source")}
    step_bytes = cells_local * (sc_bytes + eb_bytes)
    roofline["whole_step"] = {"algorithmic_bytes_per_cell": sc_bytes + eb_bytes, "achieved_gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                              "frac_of_hbm_peak": step_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak,
                              "note": "north_star's figure: fused MHD stream_collide + field update against the HBM roofline"}
    roofline["other_kernels"] = {k: {kk: v[kk] for kk in ("achieved", "frac", "ms_per_launch", "traffic")} for k, v in kernels.items()
                                 if k != dom_name and "achieved" in v}

    # ---- end to end through the public API with HOST buffers: load state -> initialize -> step -> save state ----
    e2e = None
    host_bytes = dom.n * 21 if mhd else dom.n * 17
    if host_bytes < 12e9:  # at N GPUs every rank uploads / downloads the sections of its own slab
        n = dom.n
        sections = [("flags", n, torch.uint8, 3), ("rho", n, torch.float32, 1), ("u", 3 * n, torch.float32, 2)]
        if mhd:
            sections.append(("q", n, torch.float32, 11))
        host = {name: (torch.empty(sz, dtype=dt, pin_memory=True), fid) for name, sz, dt, fid in sections}   # the state that is loaded
        saved = {name: torch.empty(sz, dtype=dt, pin_memory=True) for name, sz, dt, fid in sections}           # where a state is saved to
        for name, (tns, fid) in host.items():
            tns.numpy()[:] = dom.read(fid)
        lib = capi.load()
        k = len(sections)
        c_fields = (ctypes.c_int * k)(*[fid for _, _, _, fid in sections])
        c_out = (ctypes.c_void_p * k)(*[saved[name].data_ptr() for name, _, _, _ in sections])
        c_in = (ctypes.c_void_p * k)(*[host[name][0].data_ptr() for name, _, _, _ in sections])
        c_bytes = (ctypes.c_size_t * k)(*[host[name][0].numel() * host[name][0].element_size() for name, _, _, _ in sections])

        def e2e_step():
            # save the state the previous step left (file.rs:221-268) and load this step's (file.rs:118-152) from pinned host memory:
            # one call, downloads and uploads overlapped on two streams
            capi.check(lib.ion_buffer_swap(dom.handle, k, c_fields, c_out, c_in, c_bytes))
            lbm.initialize()                  # main.rs:274-278: a loaded state is re-initialised
            lbm.do_time_step()
            lbm.finish_queues()

        e2e_step()
        k_e2e = max(3, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if dist is not None:
            tt = torch.tensor([dt], device=f"cuda:{device}", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        nbytes = sum(t.numel() * t.element_size() for t, _ in host.values()) * world
        e2e = {"value": cells_global / dt / 1e6, "unit": "MLUPs/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "host_link_gbs_aggregate": 2 * nbytes / dt / 1e9, "numa": numa,
               "what": "per step: ion_buffer_swap (download flags/rho/u/Q as the previous step left them, upload this step's from pinned host "
                       "memory; the two directions overlap), Lbm::initialize, Lbm::do_time_step -- the save / load / step cycle of file.rs through the C ABI"}
        del saved
        del host
    else:
        e2e = {"value": None, "unit": "MLUPs/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "what": f"skipped: {host_bytes / 1e9:.0f} GB of pinned host staging per rank"}

    cpu, parity = None, None
    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)  # the CPU arm uses every core this process may run on
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(cfg, 2, 1, args.lod_depth, keep_state=True)
        cpu = {"value": r["mlups"], "unit": "MLUPs/s", "cores": r["cores"], "kind": r["kind"], "ms_per_step": r["ms"],
               "sample": f"{sample_description(cfg, args.lod_depth)}; 2 timed full time steps after 1 warm-up"}
        parity = gpu_sample_parity(args, device, r["state"], r["config"])

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MLUPs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernels": kernels, "cpu_baseline": cpu, "parity": parity, "scene_build_s": t_build,
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
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: the configuration's own (cfg4 strong, others weak)")
    ap.add_argument("--lod-depth", type=int, default=4, help="mhd_lod_depth (reference default 4, mod.rs:126)")
    ap.add_argument("--cells-z", type=int, default=0, help="override the z extent of the configuration's lattice (memory-limited boxes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precompute-mode", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="static-field precompute of the scene build: 0 reference order, 1 fast direct sum, 2 FFT convolution (default: 0, cfg3: 2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
