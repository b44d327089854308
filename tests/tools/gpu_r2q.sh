# round 2, pass q (1 GPU): FFT precompute mode (tests + scene build times), chunked ion_buffer_swap / NUMA binding in e2e, cfg1 kernel time
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "precompute or voxeli or stl or swap or e2e" > gpurun_out/r2q_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo bench rc=$?; grep precompute_B gpurun_out/r2q_bench.err
for m in 1 2; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precompute-mode $m > gpurun_out/r2q_bench_pm$m.json 2> gpurun_out/r2q_bench_pm$m.err; echo "mode $m rc=$?"; grep precompute_B gpurun_out/r2q_bench_pm$m.err
done
timeout 900 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline --precompute-mode 2 > gpurun_out/r2q_bench_cfg3_fft.json 2> gpurun_out/r2q_bench_cfg3_fft.err; echo cfg3 rc=$?; tail -2 gpurun_out/r2q_bench_cfg3_fft.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2q_cfg1_launches.csv python bench.py --config cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_cfg1_ncu.log 2>&1; echo ncu rc=$?
timeout 300 python bench.py --config cfg1 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/r2q_bench_cfg1.json 2> gpurun_out/r2q_bench_cfg1.err; echo cfg1 rc=$?
