# round 2, pass h (1 GPU): two-domain field update timing + ncu of the two-set kernel, bench N=1 with packed butterflies
mkdir -p gpurun_out
timeout 600 python tests/tools/eb_two_domain.py 256 2 5 > gpurun_out/r2h_two_domain.log 2>&1; cat gpurun_out/r2h_two_domain.log | tail -4
timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2h_three_domain.log 2>&1; cat gpurun_out/r2h_three_domain.log | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 6 -c 2 -f -o gpurun_out/r2h_eb_fft2 python tests/tools/eb_two_domain.py 256 2 2 > gpurun_out/r2h_ncu.log 2>&1; echo ncu rc=$?
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo bench rc=$?
timeout 600 python -m pytest tests -m gpu -q -k "mhd or drift or default or tracks" > gpurun_out/r2h_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2h_pytest.log
