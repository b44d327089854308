# round 2, pass n (1 GPU): all parity tests (streamed spectra), smoke(), bench cfg2, cfg2 with streamed spectra (A/B), cfg5-like slab on one GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r2n_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r2n_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo bench rc=$?
ION_EB_FFT_BATCH=1024 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_streamed.json 2> gpurun_out/r2n_bench_streamed.err; echo streamed rc=$?
timeout 1200 python bench.py --config cfg5 --cells-z 192 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_cfg5.json 2> gpurun_out/r2n_bench_cfg5.err; echo cfg5 rc=$?; tail -3 gpurun_out/r2n_bench_cfg5.err
