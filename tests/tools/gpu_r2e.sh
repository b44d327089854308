# round 2, pass e (1 GPU): parity tests, bench with CPU baseline + parity leg, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo launches rc=$?
