"""N > 1 coverage: world_size-2 runs of the time step.

CPU (gloo, runs everywhere): each rank owns one oracle domain and exchanges halos and LOD pyramids over torch.distributed
using the product's own exchange plan (ion_neighbor_domains / ion_lod_exchange_plan / ion_lbm_make_params); the result has
to be bit-identical to the single-process multi-domain oracle.  GPU (nccl, needs 2 devices): the product's
one-process-per-GPU path (Lbm.new_distributed, ion_comm_*) against the same oracle."""
import os
import socket
import subprocess
import sys

import pytest

import cases
from oracle import ref_host as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(mode, case, world=2, steps=3, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), WORKER, "--mode", mode, "--case", case, "--steps", str(steps)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, f"{' '.join(cmd)}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r.stdout


@pytest.mark.parametrize("case", ["z2_d3q19_fp32", "mhd_z2_d3q19_fp32_lod2"])
def test_two_ranks_gloo_exchange_plan_matches_single_process_oracle(case):
    out = launch("gloo", case)
    assert out.count("bit-identical to the single-process oracle") == 2, out


def test_exchange_plan_against_reference_arithmetic():
    """ion_neighbor_domains / ion_lod_exchange_plan vs the restatement of mod.rs:386-404,448-465 in oracle/ref_host.py,
    on a 3x2x2 decomposition (no GPU)."""
    from ionsolver_b200 import capi
    cfg = cases._mhd(rh.RefConfig(velocity_set="D3Q19", float_type="FP32", n_x=48, n_y=32, n_z=32, d_x=3, d_y=2, d_z=2, nu=0.05,
                                  ext_volume_force=True, ext_magneto_hydro=True, mhd_lod_depth=3))
    d_n = cfg.d_x * cfg.d_y * cfg.d_z
    pc = cases.to_lbm_config(cfg)

    def get_offset(depth):
        return sum((1 << i) ** 3 for i in range(0, depth + 1))

    for d in range(d_n):
        x, y, z = rh.domain_coords(d, cfg.d_x, cfg.d_y)
        want = {0: ((x + 1) % cfg.d_x) + (y + z * cfg.d_y) * cfg.d_x, 1: x + (((y + 1) % cfg.d_y) + z * cfg.d_y) * cfg.d_x,
                2: x + (y + ((z + 1) % cfg.d_z) * cfg.d_y) * cfg.d_x}
        for axis in range(3):
            dp, dm = capi.neighbor_domains(cfg.d_x, cfg.d_y, cfg.d_z, d, axis)
            assert dp == want[axis]
            assert capi.neighbor_domains(cfg.d_x, cfg.d_y, cfg.d_z, dm, axis)[0] == d
        params = pc.make_params(d)
        g = rh.domain_geometry(cfg, x, y, z, d)
        offset = g.n_lod_own
        for dc in range(d_n):
            src, cnt, dst = capi.lod_exchange_plan(params, dc)
            if dc == d:
                assert cnt == 0
                continue
            dx, dy, dz = rh.domain_coords(dc, cfg.d_x, cfg.d_y)
            depth = max(0, cfg.mhd_lod_depth - max(abs(z - dz), abs(y - dy), abs(x - dx)))
            rs, re = get_offset(depth - 1), get_offset(depth)
            assert (src, cnt, dst) == (rs, re - rs, offset)
            offset += re - rs
        assert offset == g.n_lod == params.n_lod


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["z2_d3q19_fp32", "mhd_z2_d3q19_fp32_lod2", "mhd_z2_d3q19_fp32_lod4"])
def test_two_ranks_nccl(case, gpu_lib):
    from ionsolver_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = launch("nccl", case)
    assert out.count("parity ok") == 2, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["z2_d3q19_fp32", "z3_d3q19_fp16s_trt", "mhd_z2_d3q19_fp32_lod2"])
def test_domains_on_two_gpus_peer_copies(name, gpu_lib):
    """One process, domains spread over two GPUs: halos travel by cudaMemcpyPeerAsync (ion_exchange_transfer)."""
    import numpy as np
    from ionsolver_b200 import capi
    from oracle_util import same_bits
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from test_gpu_parity import product
    cfg = dict(cases.all_cases())[name]
    mhd = cfg.ext_magneto_hydro
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    gpu = product(cfg, devices=[0, 1])
    cases.upload_inputs(ref, gpu)
    ref.initialize()
    gpu.initialize()
    if mhd:
        cases.seed_electron_gas(ref, gpu)
        for rd, gd in zip(ref.domains, gpu.domains):
            gd.write(cases.FIELD_OF["e_dyn"], rd.e_dyn)
            gd.write(cases.FIELD_OF["b_dyn"], rd.b_dyn)
    for _ in range(1 if mhd else 4):
        ref.do_time_step()
        gpu.do_time_step()
    gpu.finish_queues()
    for rd, gd in zip(ref.domains, gpu.domains):
        for n in ["fi", "flags"] + (["ei", "fqi", "qc"] if mhd else ["rho", "u"]):
            want = getattr(rd, n)
            assert same_bits(np.asarray(gd.read(cases.FIELD_OF[n])).view(want.dtype), want), f"domain {rd.g.d_i} {n}"
    gpu.close()
