# round 2, pass l (1 GPU): all parity tests (adaptive far blocks, ion_buffer_swap e2e), cfg4-like thin slabs timing, full bench with CPU leg
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo bench rc=$?; tail -2 gpurun_out/r2l_bench.err
timeout 600 python bench.py --config cfg1 --steps 200 --warmup 20 > gpurun_out/r2l_bench_cfg1.json 2> gpurun_out/r2l_bench_cfg1.err; echo cfg1 rc=$?
timeout 900 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg3.json 2> gpurun_out/r2l_bench_cfg3.err; echo cfg3 rc=$?; tail -2 gpurun_out/r2l_bench_cfg3.err
timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg4.json 2> gpurun_out/r2l_bench_cfg4.err; echo cfg4 rc=$?; tail -2 gpurun_out/r2l_bench_cfg4.err
