# round 2, pass r (1 GPU): FP64 transforms in the FFT precompute mode (test with printed errors, scene build times); streamed kernel spectra
# with two source sets (three 256^3 slabs on one GPU): event times per domain and an ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "precompute or voxeli or stl" > gpurun_out/r2r_pytest.log 2>&1; echo pytest rc=$?; grep -E "relative L2|passed|failed" gpurun_out/r2r_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precompute-mode 2 > gpurun_out/r2r_bench_pm2.json 2> gpurun_out/r2r_bench_pm2.err; echo "mode 2 rc=$?"; grep precompute_B gpurun_out/r2r_bench_pm2.err
timeout 900 python bench.py --config cfg3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_cfg3.json 2> gpurun_out/r2r_bench_cfg3.err; echo cfg3 rc=$?; tail -c 200 gpurun_out/r2r_bench_cfg3.json
timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2r_three_static.log 2>&1; cat gpurun_out/r2r_three_static.log
ION_EB_FFT_BATCH=1024 timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2r_three_streamed.log 2>&1; cat gpurun_out/r2r_three_streamed.log
ION_EB_FFT_BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_three_streamed_launches.csv -k regex:k_eb python tests/tools/eb_two_domain.py 256 3 1 > gpurun_out/r2r_ncu.log 2>&1; echo ncu rc=$?
