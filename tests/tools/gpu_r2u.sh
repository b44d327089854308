# round 2, pass u (1 GPU): far-slab reduction moved into eb_fft_core.cuh (far_reduce_x / far_eval_x) -- multi-domain MHD parity, error behaviour, three slabs timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "multi_domain_mhd or error_behaviour or precompute" > gpurun_out/r2u_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2u_pytest.log
timeout -k 10 300 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2u_three_static.log 2>&1; cat gpurun_out/r2u_three_static.log
