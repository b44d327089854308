"""Scratch: per-step divergence of deterministic mode vs the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import ref_host as rh
from ionsolver_b200 import lbm as L

def diff(tag, ref, gpu, names):
    for rd, gd in zip(ref.domains, gpu.domains):
        for n in names:
            want = getattr(rd, n); got = gd.read(cases.FIELD_OF[n])
            if n.startswith("transfer"): got = got[:want.size]
            if got.tobytes() != want.tobytes():
                idx = np.nonzero(got != want)[0]
                print(f"  {tag} dom{rd.g.d_i} {n}: {len(idx)} differ, first {idx[:5]} got {got[idx[:3]]} want {want[idx[:3]]}")

ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
NSTEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for name, cfg in cases.mhd_cases() + cases.multi_domain_mhd_cases():
    if ONLY and ONLY not in name:
        continue
    print("CASE", name, flush=True)
    ref = rh.RefLbm(cfg, threads=1, backend="port"); cases.fill_inputs(ref, cfg)
    gpu = L.Lbm(cases.to_lbm_config(cfg, True), devices=[0]); cases.upload_inputs(ref, gpu)
    ref.initialize(); gpu.initialize()
    names = ["qu_lod", "e_dyn", "b_dyn", "qc", "fi", "ei", "fqi", "u"]
    diff("init", ref, gpu, names)
    cases.seed_electron_gas(ref, gpu)
    for s in range(NSTEPS):
        # piecewise step
        for d in ref.domains: d.enqueue_clear_qu_lod()
        for d in ref.domains: d.enqueue_stream_collide()
        for d in gpu.domains: d.enqueue_clear_qu_lod()
        for d in gpu.domains: d.enqueue_stream_collide(gpu.get_time_step())
        diff(f"step{s} after stream_collide", ref, gpu, ["qu_lod", "qc", "fi", "ei", "fqi", "u", "rho"])
        ref.communicate_fi(); gpu.communicate_fi()
        if len(ref.domains) > 1:
            for d in ref.domains: d.enqueue_lod_part_2_gather()
            for d in gpu.domains: d.enqueue_lod_part_2_gather()
        ref.communicate_fqi(); ref.communicate_ei(); ref.communicate_qu_lods()
        gpu.communicate_fqi(); gpu.communicate_ei(); gpu.communicate_qu_lods()
        diff(f"step{s} after comm", ref, gpu, ["qu_lod", "fi", "ei", "fqi"])
        ref.update_e_b_dynamic()
        for d in gpu.domains: d.enqueue_update_e_b_dyn()
        diff(f"step{s} after update_e_b", ref, gpu, ["e_dyn", "b_dyn"])
        ref.increment_timestep(1); gpu.set_time_step(gpu.get_time_step() + 1)
    gpu.close()
