# round 2, pass k (1 GPU): all parity tests (boundary-first schedule, far path with 8^3 blocks), three-domain timing, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2k_pytest.log
timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2k_three_domain.log 2>&1; cat gpurun_out/r2k_three_domain.log | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo bench rc=$?
