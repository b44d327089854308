# One GPU-box round: smoke, parity tests, bench (both LOD depths), ncu launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo bench rc=$?
cat gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 900 python bench.py --steps 10 --warmup 3 --lod-depth 3 --no-cpu-baseline > gpurun_out/bench_r1_lod3.json 2>> gpurun_out/bench_r1.err; echo bench3 rc=$?
cat gpurun_out/bench_r1_lod3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/ncu_bench.log
