# round 2, pass w (1 GPU): the mirrored-spectra test alone (after the B_dyn = 0 fix of its assertion)
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -q -k "mirrored" > gpurun_out/r2w_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2w_pytest.log
