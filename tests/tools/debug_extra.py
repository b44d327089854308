"""debug: first differences of the x_mhd_d3q19_fp32_trt_lod2 case between oracle (port) and GPU, step by step"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cases
from oracle import ref_host as rh
from oracle_util import buffer_names
from ionsolver_b200 import lbm as L
name = sys.argv[1] if len(sys.argv) > 1 else "x_mhd_d3q19_fp32_trt_lod2"
cfg = dict(cases.extra_oracle_cases())[name]
if len(sys.argv) > 2 and sys.argv[2] == "srt":
    cfg.relaxation_time = "SRT"
ref = rh.RefLbm(cfg, threads=1, backend="port")
cases.fill_inputs(ref, cfg, seed=9)
gpu = L.Lbm(cases.to_lbm_config(cfg, True), devices=[0])
cases.upload_inputs(ref, gpu)
ref.initialize(); gpu.initialize()
cases.seed_electron_gas(ref, gpu)
for step in range(4):
    ref.do_time_step(); gpu.do_time_step(); gpu.finish_queues()
    rd, gd = ref.domains[0], gpu.domains[0]
    line = [f"step {step}:"]
    for n in buffer_names(cfg):
        want = getattr(rd, n); got = gd.read(cases.FIELD_OF[n])
        got = np.asarray(got).view(want.dtype) if got.dtype != want.dtype else got
        neq = np.flatnonzero(got.view(np.uint32 if want.dtype.itemsize == 4 else np.uint8) != want.view(np.uint32 if want.dtype.itemsize == 4 else np.uint8))
        if neq.size:
            k = neq[0]
            line.append(f"{n}: {neq.size} differ, first idx {k} ref {want[k]!r} gpu {got[k]!r} nan_ref {int(np.isnan(want.astype(np.float64)).sum()) if want.dtype.kind=='f' else 0}")
    print(" | ".join(line))
