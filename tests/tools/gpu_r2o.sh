# round 2, pass o (1 GPU): all parity tests after the window guard of eb_fft_supported went away (ragged slabs now use the FFT path)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r2o_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r2o_smoke.log
