#!/usr/bin/env python
"""Turns the raw outputs of tests/tools/gpu_final.sh (gpurun_out/) into the markdown kept under profiles/:
   r1_launches.md       from launches_<tag>.csv (ncu --metrics gpu__time_duration.sum) + profiles/r1_bench.json
   r1_kernel_matrix.md  from profiles/r1_kernel_matrix.jsonl
usage: python tests/tools/make_profiles.py <tag>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FIRST = {  # the same rows at the first GPU run of the round (git history of profiles/r1_kernel_matrix.md)
    "cfg1-like 64^3 D3Q19 FP32 SRT (fits L2)": 0.572, "D3Q19 FP32 SRT": 0.661, "D3Q19 FP32 TRT": 0.531, "D3Q19 FP16S SRT": 0.362,
    "D3Q19 FP16C SRT": 0.236, "D3Q27 FP32 SRT": 0.727, "cfg2 D3Q19 FP32 MHD": 0.827, "cfg3 shape 512x256x256 D3Q19 FP32 MHD": 0.775,
    "cfg4 kernel D3Q27 FP16S MHD": 0.369, "cfg5 kernel D3Q19 FP16C MHD": 0.229}


def launches(tag):
    with open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"], {"n": 0, "ns": 0.0, "block": row["Block Size"], "grid": row["Grid Size"]})
        a["n"] += 1
        a["ns"] += v
    out = ["# r1 — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (256^3 D3Q19 FP32 MHD, LOD depth 4), final build of the round", "",
           f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_{tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`.",
           "Per-launch times under ncu are serialised and cold-cache; what must agree with bench.py is each kernel's SHARE of a step.", "",
           "| kernel | launches | avg ns / launch | block | grid |", "|---|---:|---:|---|---|"]
    for k, a in agg.items():
        out.append(f"| `{k[:150]}` | {a['n']} | {a['ns'] / a['n']:.0f} | {a['block']} | {a['grid']} |")

    def avg(sub):
        tot = [(a["ns"], a["n"]) for k, a in agg.items() if sub in k]
        return sum(t[0] for t in tot) / max(sum(t[1] for t in tot), 1)
    sc, eb, cl, bs, fo = (avg(s) for s in ("k_stream_collide", "k_update_e_b", "k_clear_qu_lod", "k_build_sources", "k_lod_fold"))
    step = sc + eb + cl + bs + fo
    j = json.load(open(os.path.join(ROOT, "profiles", "r1_bench.json")))
    ks = j["kernels"]
    out += ["", f"One time step = clear_qu_lod + stream_collide + lod_fold + build_sources + update_e_b: {step / 1e6:.3f} ms under ncu.",
            f"Shares under ncu: update_e_b {eb / step:.3f}, stream_collide (+ lod_fold) {(sc + fo) / step:.3f}.",
            f"Shares from bench.py CUDA events (same build, no profiler, profiles/r1_bench.json): update_e_b {ks['update_e_b_dynamic']['share_of_step']:.3f}, "
            f"stream_collide {ks['stream_collide']['share_of_step']:.3f} (step {j['ms_per_step']:.3f} ms, {j['value']:.1f} MLUPs/s).",
            "", "One-off scene construction in the same run: `k_psi` (magnetic scalar potential of the voxelised disk magnet, compacted sources), `k_voxelize`, "
            "`k_initialize`; `k_fma_peak` is bench.py's FP32 issue-peak probe."]
    open(os.path.join(ROOT, "profiles", "r1_launches.md"), "w").write("\n".join(out) + "\n")


def matrix():
    rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r1_kernel_matrix.jsonl"))]
    out = ["# r1 — stream_collide over the configurations of BASELINE.json (one B200; `python tests/tools/kernel_matrix.py --shapes 512x512x512`)", "",
           "CUDA events on the domain stream, 20 back-to-back steps (clear_qu_lod + stream_collide + lod_fold for the MHD rows, LOD depth 3) after 3 warm-up steps;",
           "bytes per cell = SURVEY.md 8(d); peak = 6534.5 GB/s (MEASURED_PEAKS.json, measured copy bandwidth).  Raw lines: `profiles/r1_kernel_matrix.jsonl`.",
           "`first` = the same row at the first GPU run of the round (before: pinned loads, compile-time parity, packed two-species math, FP16C codec by",
           "multiplication, LOD replicas, per-family occupancy, four-cells-per-thread vector kernel for the plain FP32 rows).", "",
           "| configuration | cells | ms / launch | MLUPs/s | B / cell | achieved GB/s | frac of HBM copy peak | first |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for j in rows:
        o = FIRST.get(j["config"])
        out.append(f"| {j['config']} | {j['cells']} | {j['ms']} | {j['mlups']} | {j['bytes_per_cell']} | {j['achieved_gbs']} | **{j['frac_of_hbm_peak']}** | {o if o else '—'} |")
    out += ["", "Reading:", "",
            "* FP32 kernels are HBM-bound: 0.97 (plain D3Q19, four cells per thread with 128-bit accesses), 0.82 (D3Q27), 0.90-0.95 (MHD) of the measured copy peak,",
            "  DRAM traffic = algorithmic bytes (ncu).",
            "* The MHD kernel no longer degrades with lattice size (512^3: 0.63 -> 0.90): the loss was the serialisation of same-address LOD reductions in L2, removed by",
            "  private replicas + one 16-byte vector reduction per warp run (`profiles/r1_stream_collide_ab.md`).",
            "* FP16S / FP16C kernels move half the bytes with the same FP32 arithmetic per cell, so they are instruction-issue bound (ncu: issue slots 70-80 % busy,",
            "  ALU pipe 50-60 %): 0.45-0.53 of the HBM peak, i.e. 1.2-1.4x the FP32 kernels' MLUPs/s.  TRT adds ~20 % instructions per cell (0.72).",
            "* The 64^3 row fits the 126 MB L2 and is launch/latency dominated (9 us per launch); it is reported, not used for the HBM claim."]
    open(os.path.join(ROOT, "profiles", "r1_kernel_matrix.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    launches(sys.argv[1] if len(sys.argv) > 1 else "r1e")
    matrix()
