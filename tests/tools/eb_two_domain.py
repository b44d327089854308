"""Two z slabs of 256^3 on ONE GPU (the reference's own multi-domain fallback, opencl.rs:56-61): domain 1 has a lower neighbour, so its
update_e_b_dynamic runs the two-source-set FFT kernel (k_eb_fft<16,2>) plus nothing else (no far slabs).  Used to profile that kernel
on one GPU (`ncu -k regex:k_eb_fft`) and to time both domains' field updates with CUDA events."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ionsolver_b200 import lbm as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dz = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, relaxation_time=L.RelaxationTime.Srt, float_type=L.FloatType.FP32, n_x=n, n_y=n, n_z=n * dz, d_z=dz,
                ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=4, graphics_config=L.GraphicsConfig(False))
c.units.set(float(n), 1.0, 1.0, 1.0, 1.0, 0.1, 1.0, 1.2250, 0.0000000001, 1.0)
c.nu = c.units.nu_si_lu(1.48E-5)
lbm = L.Lbm(c, devices=[0] * dz)
for d in lbm.domains:
    d.write(11, np.full(d.n, 0.002, np.float32))
    b = np.zeros(3 * d.n, np.float32)
    b[2 * d.n:] = 0.01
    d.write(6, b)
lbm.setup_velocity_field((0.1, 0.01, 0.0), 1.0)
lbm.initialize()
for _ in range(3):
    lbm.do_time_step()
lbm.finish_queues()
for i, d in enumerate(lbm.domains):
    st = torch.cuda.ExternalStream(d.stream(), device=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        d.enqueue_update_e_b_dyn()
    e1.record(st)
    torch.cuda.synchronize()
    print(f"domain {i}: update_e_b_dynamic {e0.elapsed_time(e1) / steps:.3f} ms, fft info {d.eb_fft_info()}", flush=True)
lbm.close()
