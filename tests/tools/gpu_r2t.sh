# round 2, pass t (1 GPU): k_eb_fft with dynamically scheduled products (idle warps form the next group's products) and named barriers
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -x -k "mhd or streamed or default_mode" > gpurun_out/r2t_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2t_pytest.log
timeout -k 10 300 python tests/tools/eb_two_domain.py 256 3 5 > gpurun_out/r2t_three_static.log 2>&1; cat gpurun_out/r2t_three_static.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo bench rc=$?
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2t_bench.json'))
print(j['value'], j['ms_per_step'], {k:v.get('ms_per_launch', v.get('ms')) for k,v in j['kernels'].items()})
PY
