# round 2, pass p (1 GPU): K^ evaluation in FP32 with two real lines per x FFT -- MHD parity tests, bench cfg2 (static), cfg2 streamed, cfg5-like slab
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -k "mhd or streamed or default_mode or drift" > gpurun_out/r2p_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2p_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo bench rc=$?
ION_EB_FFT_BATCH=1024 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench_streamed.json 2> gpurun_out/r2p_bench_streamed.err; echo streamed rc=$?
timeout 1200 python bench.py --config cfg5 --cells-z 192 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_cfg5.json 2> gpurun_out/r2p_bench_cfg5.err; echo cfg5 rc=$?; tail -3 gpurun_out/r2p_bench_cfg5.err
