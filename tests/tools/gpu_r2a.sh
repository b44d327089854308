# round 2, first GPU pass of the polyphase-FFT field update: parity tests, bench, launch list, ncu of k_eb_fft
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo bench rc=$?; cat gpurun_out/r2a_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step']); print({k:v.get('ms') for k,v in j['kernels'].items()}); print(j['e2e'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 3 -c 1 -f -o gpurun_out/r2a_eb_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1; echo ncu rc=$?
