"""One rank of a multi-process run of the time step (launched by tests/test_distributed.py through
`python -m torch.distributed.run`; RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* come from the environment).

--mode gloo  (CPU, no GPU needed): every rank owns ONE oracle domain and performs the exchange of
             Lbm::communicate_field / communicate_qu_lods (mod.rs:371-468) over torch.distributed with the PRODUCT's
             exchange plan -- ion_neighbor_domains, ion_lod_exchange_plan and IonParams from ion_lbm_make_params, the same
             functions ion_lbm_* / ion_comm_* use -- and the pairing rule of ion_comm_exchange_transfer (my +face becomes
             the +neighbour's transfer_m, my -face the -neighbour's transfer_p).
--mode nccl  (one GPU per rank): the product itself, Lbm.new_distributed -> NCCL send/recv + all-gather over NVLink.
Both compare the rank's own domain with the single-process multi-domain oracle run on the same seeded inputs."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import cases
from oracle import ref_host as rh
from oracle_util import rel_l2, same_bits


class DistOracle:
    """Lbm of mod.rs restricted to the domain of this rank; exchanges go through torch.distributed."""

    def __init__(self, cfg, rank, world):
        from ionsolver_b200 import capi
        from oracle.port import PortDomain
        self.capi, self.cfg, self.rank, self.world = capi, cfg, rank, world
        x, y, z = rh.domain_coords(rank, cfg.d_x, cfg.d_y)
        self.dom = PortDomain(cfg, x, y, z, rank, threads=1)
        self.params = cases.to_lbm_config(cfg).make_params(rank)  # product: get_device_defines as IonParams
        assert (self.params.n_lod, self.params.n_lod_own) == (self.dom.g.n_lod, self.dom.g.n_lod_own)

    def communicate_field(self, field, bytes_per_cell):
        c, d = self.cfg, self.dom
        for axis, dn in enumerate((c.d_x, c.d_y, c.d_z)):
            if dn <= 1:
                continue
            d.enqueue_transfer_extract_field(field, axis)
            dp, dm = self.capi.neighbor_domains(c.d_x, c.d_y, c.d_z, self.rank, axis)
            nbytes = d.get_area(axis) * bytes_per_cell
            new_p, new_m = np.zeros_like(d.transfer_p), np.zeros_like(d.transfer_m)
            t = torch.from_numpy
            ops = [dist.P2POp(dist.isend, t(d.transfer_p)[:nbytes], dp, tag=0), dist.P2POp(dist.isend, t(d.transfer_m)[:nbytes], dm, tag=1),
                   dist.P2POp(dist.irecv, t(new_m)[:nbytes], dm, tag=0), dist.P2POp(dist.irecv, t(new_p)[:nbytes], dp, tag=1)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            d.transfer_p, d.transfer_m = new_p, new_m
            d.enqueue_transfer_insert_field(field, axis)

    def communicate_fi(self):
        self.communicate_field(rh.TF_FI, rh.FLOAT_SIZE[self.cfg.float_type] * self.dom.transfers)

    def communicate_ei(self):
        self.communicate_field(rh.TF_EI, rh.FLOAT_SIZE[self.cfg.float_type] * self.dom.transfers)

    def communicate_fqi(self):
        self.communicate_field(rh.TF_QI, rh.FLOAT_SIZE[self.cfg.float_type])

    def communicate_rho_u_flags(self):
        self.communicate_field(rh.TF_RHO_U_FLAGS, 17)

    def communicate_qu_lods(self):
        d = self.dom
        own = torch.from_numpy(d.qu_lod[:4 * d.g.n_lod_own].copy())
        parts = [torch.empty_like(own) for _ in range(self.world)]
        dist.all_gather(parts, own)
        for dc in range(self.world):
            src, cnt, dst = self.capi.lod_exchange_plan(self.params, dc)
            if cnt:
                d.qu_lod[4 * dst:4 * (dst + cnt)] = parts[dc].numpy()[4 * src:4 * (src + cnt)]

    def initialize(self):  # mod.rs:214-231
        d, mhd = self.dom, self.cfg.ext_magneto_hydro
        d.t += 1
        self.communicate_rho_u_flags()
        d.enqueue_initialize()
        self.communicate_rho_u_flags()
        self.communicate_fi()
        if mhd:
            self.communicate_fqi()
            self.communicate_ei()
            self.communicate_qu_lods()
            d.enqueue_update_e_b_dyn()
        d.t = 0

    def do_time_step(self):  # mod.rs:250-272
        d, mhd = self.dom, self.cfg.ext_magneto_hydro
        if mhd:
            d.enqueue_clear_qu_lod()
        d.enqueue_stream_collide()
        if self.cfg.graphics_active:
            self.communicate_rho_u_flags()
        self.communicate_fi()
        if mhd:
            d.enqueue_lod_part_2_gather()
            self.communicate_fqi()
            self.communicate_ei()
            self.communicate_qu_lods()
            d.enqueue_update_e_b_dyn()
        d.t += 1


def case_by_name(name):
    for n, cfg in cases.all_cases():
        if n == name:
            return cfg
    raise KeyError(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["gloo", "nccl"], required=True)
    ap.add_argument("--case", required=True)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    cfg = case_by_name(args.case)
    assert cfg.d_x * cfg.d_y * cfg.d_z == world, "one domain per rank"
    mhd = cfg.ext_magneto_hydro

    # the single-process multi-domain oracle: what every rank's domain has to equal
    ref = rh.RefLbm(cfg, threads=1, backend="port")
    cases.fill_inputs(ref, cfg)
    rd = ref.domains[rank]
    inputs = {n: getattr(rd, n).copy() for n in ("rho", "u", "flags") + (("qc", "b_stat", "e_stat") if mhd else ()) + (("f",) if cfg.ext_force_field else ())}

    if args.mode == "gloo":
        dist.init_process_group("gloo")
        me = DistOracle(cfg, rank, world)
        for n, v in inputs.items():
            getattr(me.dom, n)[:] = v
        ref.initialize()
        me.initialize()
        if mhd:
            cases.seed_electron_gas(ref)
            me.dom.ei[:] = cases.electron_gas_at_rest(me.dom, cfg)
        for _ in range(args.steps):
            ref.do_time_step()
            me.do_time_step()
        names = ["fi", "rho", "u", "flags", "transfer_p", "transfer_m"] + (["ei", "fqi", "qc", "qu_lod", "e_dyn", "b_dyn"] if mhd else [])
        bad = [n for n in names if not same_bits(getattr(me.dom, n), getattr(rd, n))]
        ok = not bad
        print(f"rank {rank}: gloo {args.case} {args.steps} steps: {'bit-identical to the single-process oracle' if ok else 'MISMATCH ' + str(bad)}", flush=True)
    else:
        from ionsolver_b200 import capi, lbm as L
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ident = [L.Lbm.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        gpu = L.Lbm.new_distributed(cases.to_lbm_config(cfg), rank, world, local, ident[0])
        gd = gpu.domains[0]
        assert gd.d_i == rank
        for n, v in inputs.items():
            gd.write(cases.FIELD_OF[n], v)
        ref.initialize()
        gpu.initialize()
        if mhd:
            cases.seed_electron_gas(ref)
            gd.write(cases.FIELD_OF["ei"], cases.electron_gas_at_rest(rd, cfg))
            gd.write(cases.FIELD_OF["e_dyn"], rd.e_dyn)
            gd.write(cases.FIELD_OF["b_dyn"], rd.b_dyn)
        steps = 1 if mhd else args.steps  # MHD: E/B feed back into the DDFs, compare after one step like test_multi_domain_mhd
        for _ in range(steps):
            ref.do_time_step()
            gpu.do_time_step()
        gpu.finish_queues()
        exact = ["fi", "flags"] + (["ei", "fqi", "qc"] if mhd else ["rho", "u"])
        bad = []
        for n in exact:
            got, want = gd.read(cases.FIELD_OF[n]), getattr(rd, n)
            if not same_bits(np.asarray(got).view(want.dtype), want):
                bad.append(n)
        if mhd:
            for n, tol in (("qu_lod", 4e-6), ("e_dyn", 1e-5), ("b_dyn", 1e-5)):
                err = rel_l2(gd.read(cases.FIELD_OF[n]), getattr(rd, n))
                if not err < tol:
                    bad.append(f"{n} rel-L2 {err:.3g}")
        ok = not bad
        print(f"rank {rank}: nccl {args.case} {steps} steps on cuda:{local}: {'parity ok' if ok else 'MISMATCH ' + str(bad)}; "
              f"kernels launched {capi.kernel_launch_count()}", flush=True)
        gpu.close()
    flag = torch.tensor([0 if ok else 1], device="cuda" if args.mode == "nccl" else "cpu")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
