"""Scratch: rough kernel timings (wall clock around ion_finish) for 256^3."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_host as rh
from ionsolver_b200 import capi
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from first_check import params_from  # noqa

def bench(cfg, steps=20, label=""):
    g = rh.domain_geometry(cfg, 0, 0, 0, 0)
    dom = capi.Domain(params_from(cfg, g))
    n = g.n
    rng = np.random.default_rng(0)
    dom.write(capi.FIELD_U, (0.05 * rng.standard_normal(3 * n)).astype(np.float32))
    mhd = cfg.ext_magneto_hydro
    if mhd:
        dom.write(capi.FIELD_Q, np.full(n, 0.002, np.float32))
    dom.enqueue_initialize()
    if mhd: dom.enqueue_update_e_b_dyn()
    for t in range(3):
        if mhd: dom.enqueue_clear_qu_lod()
        dom.enqueue_stream_collide(t)
    dom.finish()
    t0 = time.perf_counter()
    for t in range(3, 3 + steps):
        if mhd: dom.enqueue_clear_qu_lod()
        dom.enqueue_stream_collide(t)
    dom.finish()
    dt = (time.perf_counter() - t0) / steps
    q = rh.SET_VALUES[cfg.velocity_set][1]; s = rh.FLOAT_SIZE[cfg.float_type]
    bpc = (1 + 4 * q * s + 14 * s + 28) if mhd else (1 + 2 * q * s)
    print(f"{label} stream_collide: {dt*1e3:.3f} ms  {n/dt/1e6:.0f} MLUPs  {n*bpc/dt/1e9:.0f} GB/s ({bpc} B/cell)", flush=True)
    if mhd:
        dom.enqueue_update_e_b_dyn(); dom.finish()
        t0 = time.perf_counter()
        for _ in range(3): dom.enqueue_update_e_b_dyn()
        dom.finish()
        dt = (time.perf_counter() - t0) / 3
        pairs = n * (8 ** cfg.mhd_lod_depth)
        print(f"{label} update_e_b D={cfg.mhd_lod_depth}: {dt*1e3:.2f} ms  {n/dt/1e6:.0f} MLUPs  {pairs/dt/1e12:.3f} Tpairs/s", flush=True)
    dom.close()

C = rh.RefConfig
N = 256
bench(C(velocity_set="D3Q19", float_type="FP32", n_x=N, n_y=N, n_z=N, nu=0.1), label="plain FP32")
bench(C(velocity_set="D3Q19", float_type="FP16S", n_x=N, n_y=N, n_z=N, nu=0.1), label="plain FP16S")
bench(C(velocity_set="D3Q19", float_type="FP16C", n_x=N, n_y=N, n_z=N, nu=0.1), label="plain FP16C")
for D in (3, 4):
    c = C(velocity_set="D3Q19", float_type="FP32", n_x=N, n_y=N, n_z=N, nu=0.1, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=D)
    c.units.set(256.0, 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 1e-10, 1.0)
    bench(c, label=f"MHD FP32 D={D}")
c = C(velocity_set="D3Q27", float_type="FP16S", n_x=N, n_y=N, n_z=N, nu=0.1, ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3)
bench(c, label="MHD D3Q27 FP16S")
