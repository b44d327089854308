# round 2, pass j (1 GPU): MHD parity subset, three-domain field update timing + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "mhd or tracks or default" > gpurun_out/r2j_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2j_pytest.log
timeout 600 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2j_three_domain.log 2>&1; cat gpurun_out/r2j_three_domain.log | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2j_launches3.csv python tests/tools/eb_two_domain.py 256 3 2 > /dev/null 2>&1; echo launches rc=$?
