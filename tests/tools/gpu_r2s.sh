# round 2, pass s (1 GPU): far-slab polynomials reduced to x per row in k_eb_combine -- multi-domain MHD parity (tall slabs use the far path),
# three 256^3 slabs on one GPU (event times per domain), a cfg5-shaped three-slab lattice (2048 x 2048 x 3*64, FP32 state would not fit: FP16C via bench is 8-GPU only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "multi_domain_mhd or streamed" > gpurun_out/r2s_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2s_pytest.log
timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2s_three_static.log 2>&1; cat gpurun_out/r2s_three_static.log
timeout 600 python tests/tools/eb_two_domain.py 256 8 5 > gpurun_out/r2s_eight_static.log 2>&1; cat gpurun_out/r2s_eight_static.log
