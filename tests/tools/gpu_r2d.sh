# round 2, third GPU pass: parity tests, bench, ncu of k_eb_fft, drift curves
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?; cat gpurun_out/r2d_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step']); print({k:v.get('ms') for k,v in j['kernels'].items()}); print(j['e2e'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 3 -c 1 -f -o gpurun_out/r2d_eb_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1; echo ncu rc=$?

