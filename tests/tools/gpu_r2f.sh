# round 2, pass e (1 GPU): parity tests, bench with CPU baseline + parity leg, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 3 -c 1 -f -o gpurun_out/r2f_eb_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1; echo ncu rc=$?
