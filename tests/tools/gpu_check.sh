# Re-validation on a GPU box: tensor-path micro-benchmark, smoke, parity tests, one bench line. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_peak tests/tools/mma_peak.cu && timeout 120 /tmp/mma_peak > gpurun_out/mma_peak.jsonl; cat gpurun_out/mma_peak.jsonl
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?
tail -4 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo bench rc=$?
cat gpurun_out/bench_check.json; tail -5 gpurun_out/bench_check.err
