# quick GPU check: parity tests + bench at LOD depth 4 and 3 (no CPU baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
for d in 4 3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --lod-depth $d --no-cpu-baseline > gpurun_out/bench_q$d.json 2> gpurun_out/bench_q$d.err; echo bench rc=$?
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_q$d.json"))
print("lod", $d, "MLUPs", round(j["value"],1), "ms/step", round(j["ms_per_step"],3), {k:(round(v["ms"],3)) for k,v in j["kernels"].items()}, "pairs/s %.3e" % j["kernels"]["update_e_b_dynamic"]["pairs_per_s"])
PY
done
if [ -n "$SCALAR_TOO" ]; then ION_EB_SCALAR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; j=json.load(sys.stdin); print('scalar lod4', j['kernels']['update_e_b_dynamic']['ms'])"; fi
