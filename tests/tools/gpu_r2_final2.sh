# round 2, final 1-GPU pass (final build): parity tests, smoke, bench (all configurations that fit one GPU), launch list, ncu captures
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2_final_pytest.log
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r2_final_smoke.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo bench rc=$?; tail -2 gpurun_out/r2_final_bench.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo launches rc=$?
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 3 -c 1 -f -o gpurun_out/r2_final_eb_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu1.log 2>&1; echo ncu1 rc=$?
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -f -o gpurun_out/r2_final_sc python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu2.log 2>&1; echo ncu2 rc=$?
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_combine -s 3 -c 1 -f -o gpurun_out/r2_final_eb_combine python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu3.log 2>&1; echo ncu3 rc=$?
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 6 -c 2 -f -o gpurun_out/r2_final_eb_fft2 python tests/tools/eb_two_domain.py 256 2 2 > gpurun_out/r2_final_ncu4.log 2>&1; echo ncu4 rc=$?
timeout -k 10 600 python bench.py --config cfg1 --steps 2000 --warmup 100 > gpurun_out/r2_final_bench_cfg1.json 2> gpurun_out/r2_final_bench_cfg1.err; echo cfg1 rc=$?
timeout -k 10 900 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_cfg3.json 2> gpurun_out/r2_final_bench_cfg3.err; echo cfg3 rc=$?
timeout -k 10 900 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_cfg4.json 2> gpurun_out/r2_final_bench_cfg4.err; echo cfg4 rc=$?
timeout -k 10 900 python bench.py --lod-depth 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_bench_lod3.json 2> gpurun_out/r2_final_bench_lod3.err; echo lod3 rc=$?
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo reference rc=$?
