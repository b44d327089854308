#!/usr/bin/env python
"""Turns the raw outputs of the round-2 GPU passes (gpurun_out/) into the files kept under profiles/:

    r2_bench.json, r2_bench_n<N>.json     bench.py lines (copied)
    r2_launches.md + r2_launches_raw.csv  ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline`
    r2_ncu_eb_fft.md, r2_ncu_*.md         ncu --set full summaries (tests/tools/ncu_summary.py)
    r2_traffic.json                       DRAM bytes per launch from those captures; bench.py reads it for roofline.traffic
    r2_drift_curve.md                     error-vs-step curves of the default path (tests/tools/drift_curve.py)
    r2_scaling.md                         N = 1, 2, 4, 8 table

usage: python tests/tools/make_profiles_r2.py  (every input is optional; what is missing is skipped)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "gpurun_out")
PRO = os.path.join(ROOT, "profiles")


def first_existing(*names):
    for n in names:
        p = os.path.join(OUT, n)
        if os.path.isfile(p) and os.path.getsize(p) > 2:
            return p
    return None


def load_line(path):
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith("{"):
                return json.loads(line)
    return None


def launches(src, bench):
    rows = []
    with open(src) as f:
        for r in csv.reader(l for l in f if not l.startswith("==")):
            if len(r) > 10 and r[0].isdigit():
                rows.append(r)
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[4], {"n": 0, "ns": 0.0, "block": r[7] if len(r) > 7 else "", "grid": r[8] if len(r) > 8 else ""})
        a["n"] += 1
        a["ns"] += float(r[-1].replace(",", ""))
    shutil.copy(src, os.path.join(PRO, "r2_launches_raw.csv"))
    out = ["# r2 -- ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (cfg2: 256^3 D3Q19 FP32 MHD, LOD depth 4)", "",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/<tag>_launches.csv python bench.py "
           "--steps 4 --warmup 3 --no-cpu-baseline` (raw: `profiles/r2_launches_raw.csv`).", "Per-launch times under ncu are serialised and cold-cache; what must "
           "agree with bench.py is each kernel's SHARE of a step.", "", "| kernel | launches | avg ns / launch |", "|---|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        out.append(f"| `{k[:140]}` | {a['n']} | {a['ns'] / a['n']:.0f} |")

    def avg(sub):
        t = [(a["ns"], a["n"]) for k, a in agg.items() if sub in k]
        return sum(x[0] for x in t) / max(sum(x[1] for x in t), 1)
    sc, fold, clr = avg("k_stream_collide"), avg("k_lod_fold"), avg("k_clear_qu_lod")
    eb = avg("k_eb_src") + avg("k_eb_fft") + avg("k_eb_combine")
    step = sc + fold + clr + eb
    out += ["", f"One time step = clear_qu_lod + stream_collide + lod_fold + (k_eb_src + k_eb_fft + k_eb_combine): {step / 1e6:.3f} ms under ncu.",
            f"Shares under ncu: update_e_b_dynamic {eb / step:.3f}, stream_collide (+ lod_fold) {(sc + fold) / step:.3f}."]
    if bench:
        ks = bench["kernels"]
        out.append(f"Shares from bench.py CUDA events (same build, no profiler, profiles/r2_bench.json): update_e_b_dynamic "
                   f"{ks['update_e_b_dynamic']['share_of_step']:.3f}, stream_collide {ks['stream_collide']['share_of_step']:.3f} "
                   f"(step {bench['ms_per_step']:.3f} ms, {bench['value']:.1f} MLUPs/s).")
    out += ["", "One-off launches in the same run: `k_eb_khat` (static kernel spectra), `k_psi` / `k_voxelize` / `k_initialize` (scene construction)."]
    open(os.path.join(PRO, "r2_launches.md"), "w").write("\n".join(out) + "\n")


def ncu_summary(rep, title, dst):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ncu_summary.py"), rep, title], capture_output=True, text=True)
    if r.returncode == 0:
        open(os.path.join(PRO, dst), "w").write(r.stdout)
    return r.returncode == 0


def ncu_dram(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(k)]]
            tot += float(d[k]) * mult
        res.append((d.get("Kernel Name", "?"), tot))
    return res


def drift(src):
    shutil.copy(src, os.path.join(PRO, "r2_drift_curve.jsonl"))
    out = ["# r2 -- N-step drift of the DEFAULT (benchmarked) path against the deterministic path", "",
           "`python tests/tools/drift_curve.py --steps 100` on one B200: both runs start from the same seeded state; after every step rho, u, Q, E_dyn, B_dyn are",
           "compared (relative L2).  The deterministic path is bit-identical to the reference kernels (tests/test_gpu_parity.py); the default path differs by",
           "summation order only -- LOD deposit by warp trees + replicas, update_e_b_dynamic as a polyphase FFT convolution.  Raw: `profiles/r2_drift_curve.jsonl`.", "",
           "Scenes `fp32_lod4 ... fp16c_lod4` use `tests/cases.py::drift_scene` (well posed: velocity unit 1e8 m/s, no static fields); `fp32_lod4_reference_units` uses the",
           "unit set of the reference's scenes, where the electron gas is driven bang-bang on the sign of E (acceleration E * 0.5 / KKGE with KKGE ~ -1e-14) and the run",
           "leaves the finite range after ~17 steps IN BOTH MODES.", "",
           "Tolerances (SURVEY 8c): FP32 1e-5 (rho, u) / 1e-4 (Q, E, B); FP16S / FP16C 2e-3.", ""]
    steps_shown = (1, 2, 5, 10, 15, 20, 30, 50, 75, 100)
    for line in open(src):
        j = json.loads(line)
        c = j["rel_l2_vs_deterministic"]
        fp16 = j["float_type"] != "FP32"
        tol = {"rho": 2e-3 if fp16 else 1e-5, "u": 2e-3 if fp16 else 1e-5, "qc": 2e-3 if fp16 else 1e-4, "e_dyn": 2e-3 if fp16 else 1e-4, "b_dyn": 2e-3 if fp16 else 1e-4}
        n_ok = 0
        for s in range(j["steps"]):
            if all(c[f][s] is not None and c[f][s] <= tol[f] for f in c):
                n_ok = s + 1
            else:
                break
        out += [f"## {j['scene']}: {j['lattice'][0]}x{j['lattice'][1]}x{j['lattice'][2]} {j['float_type']}, LOD depth {j['lod_depth']}, {j['polyphase_fft_tasks']} FFT tasks; "
                f"largest N inside the tolerance: **{n_ok}**" + (f"; first non-finite step {j['first_non_finite_step'] + 1}" if j.get("first_non_finite_step") is not None else ""), "",
                "| field | " + " | ".join(f"step {s}" for s in steps_shown if s <= j["steps"]) + " |", "|---|" + "---:|" * len([s for s in steps_shown if s <= j["steps"]])]
        for f in c:
            out.append(f"| {f} | " + " | ".join(("%.1e" % c[f][s - 1]) if c[f][s - 1] is not None else "nan" for s in steps_shown if s <= j["steps"]) + " |")
        out.append("")
    open(os.path.join(PRO, "r2_drift_curve.md"), "w").write("\n".join(out) + "\n")


def scaling(lines):
    out = ["# r2 -- cfg2 weak scaling (one 256^3 slab per GPU, d_z = N), `bench.py --gpus N --steps 20 --warmup 5` under torchrun", "",
           "| GPUs | ms / step | MLUPs/s | vs N x 1 GPU | stream_collide ms | update_e_b_dynamic ms (slowest rank) | e2e MLUPs/s | e2e vs N x 1 GPU |", "|---:|---:|---:|---:|---:|---:|---:|---:|"]
    base = lines.get(1)
    for n in sorted(lines):
        j = lines[n]
        k = j["kernels"]
        eff = j["value"] / (n * base["value"]) if base else float("nan")
        e2e = j["e2e"]["value"] if j.get("e2e") and j["e2e"].get("value") else None
        e2e_eff = (e2e / (n * base["e2e"]["value"])) if (base and e2e and base["e2e"].get("value")) else float("nan")
        out.append(f"| {n} | {j['ms_per_step']:.3f} | {j['value']:.0f} | {eff:.2f} | {k['stream_collide']['ms_per_launch']:.3f} | "
                   f"{k['update_e_b_dynamic']['ms_per_launch']:.3f} | {e2e:.0f} | {e2e_eff:.2f} |")
    open(os.path.join(PRO, "r2_scaling.md"), "w").write("\n".join(out) + "\n")


def main():
    os.makedirs(PRO, exist_ok=True)
    bench = None
    b = first_existing("r2_final_bench.json", "r2i_bench.json", "r2h_bench.json")
    if b:
        bench = load_line(b)
        json.dump(bench, open(os.path.join(PRO, "r2_bench.json"), "w"), indent=1)
    lines = {1: bench} if bench else {}
    for n in (2, 4, 8):
        p = first_existing(f"r2_final_bench_n{n}.json", f"r2_scale_bench_n{n}.json")
        if p:
            lines[n] = load_line(p)
            json.dump(lines[n], open(os.path.join(PRO, f"r2_bench_n{n}.json"), "w"), indent=1)
    if len(lines) > 1:
        scaling(lines)
    l = first_existing("r2_final_launches.csv", "r2f_launches.csv")
    if l:
        launches(l, bench)
    traffic = {}
    rep = first_existing("r2_final_eb_fft.ncu-rep", "r2f_eb_fft.ncu-rep")
    if rep and ncu_summary(rep, "r2 -- k_eb_fft<16,1> (update_e_b_dynamic, polyphase FFT) at cfg2, one launch inside bench.py", "r2_ncu_eb_fft.md"):
        d = ncu_dram(rep)
        if d:
            traffic.setdefault("cfg2", {})["update_e_b_dynamic"] = {
                "dram_bytes": d[0][1], "source": "profiles/r2_ncu_eb_fft.md: dram__bytes_read.sum + dram__bytes_write.sum of k_eb_fft (k_eb_combine adds ~1.2 GB, "
                                                 "profiles/r2_ncu_eb_combine.md when captured)"}
    rep = first_existing("r2_final_sc.ncu-rep")
    if rep and ncu_summary(rep, "r2 -- k_stream_collide<D3Q19, FP32, MHD> at cfg2, one launch inside bench.py (inside the step)", "r2_ncu_stream_collide.md"):
        d = ncu_dram(rep)
        if d:
            traffic.setdefault("cfg2", {})["stream_collide"] = {"dram_bytes": d[0][1], "source": "profiles/r2_ncu_stream_collide.md: dram__bytes_read.sum + dram__bytes_write.sum"}
    if "stream_collide" not in traffic.get("cfg2", {}):
        traffic.setdefault("cfg2", {})["stream_collide"] = {"dram_bytes": 6.477e9, "source": "profiles/r1_ncu_stream_collide.md (kernel unchanged since): dram__bytes_read.sum 3.442 GB + "
                                                                                               "dram__bytes_write.sum 3.035 GB per launch"}
    rep = first_existing("r2_final_eb_combine.ncu-rep")
    if rep:
        ncu_summary(rep, "r2 -- k_eb_combine<16,1> at cfg2", "r2_ncu_eb_combine.md")
    rep = first_existing("r2_final_eb_fft2.ncu-rep", "r2h_eb_fft2.ncu-rep")
    if rep:
        ncu_summary(rep, "r2 -- k_eb_fft<16,1> and k_eb_fft<16,2> (two source sets) on two 256^3 slabs sharing one GPU (tests/tools/eb_two_domain.py)", "r2_ncu_eb_fft_two_sets.md")
    json.dump(traffic, open(os.path.join(PRO, "r2_traffic.json"), "w"), indent=1)
    d = first_existing("r2_final_drift.jsonl", "r2c_drift.jsonl")
    if d:
        drift(d)
    # the other configurations, the reference arm, the test logs: copied under fixed names
    for src, dst in (("r2_final_bench_cfg1.json", "r2_bench_cfg1.json"), ("r2_final_bench_cfg3.json", "r2_bench_cfg3.json"),
                     ("r2_final_bench_cfg4.json", "r2_bench_cfg4.json"), ("r2_final_bench_lod3.json", "r2_bench_lod3.json"),
                     ("r2_final_bench_reference.json", "r2_bench_reference_arm.json"), ("r2v_bench_cfg5.json", "r2_bench_cfg5_1gpu.json"),
                     ("r2p_bench_cfg5.json", "r2_bench_cfg5_1gpu_unmirrored.json"), ("r2_final_cfg5_n8.json", "r2_bench_cfg5_n8_unmirrored.json"),
                     ("r2v_bench_mirror.json", "r2_bench_mirror_forced.json")):
        p = first_existing(src)
        if p and load_line(p):
            json.dump(load_line(p), open(os.path.join(PRO, dst), "w"), indent=1)
    for src, dst in (("r2_final_pytest.log", "r2_pytest_gpu_1gpu.log"), ("r2_final_pytest_2gpus.log", "r2_pytest_gpu_2gpus.log"),
                     ("r2q_cfg1_launches.csv", "r2_cfg1_launches_raw.csv")):
        p = first_existing(src)
        if p:
            shutil.copy(p, os.path.join(PRO, dst))
    big = [load_line(p) for p in (first_existing("r2_final_cfg4_n8.json"), first_existing("r2_final_cfg5_n8_mirror.json", "r2_final_cfg5_n8.json"), first_existing("r2_final_bench_cfg4.json")) if p]
    if len(big) == 3:
        with open(os.path.join(PRO, "r2_scaling_cfg4_cfg5_n8.jsonl"), "w") as f:
            for j in big:
                f.write(json.dumps(j) + "\n")
    print("profiles written:", sorted(f for f in os.listdir(PRO) if f.startswith("r2_")))


if __name__ == "__main__":
    main()
