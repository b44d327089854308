# round 2, pass v (1 GPU): mirrored kernel spectra (tasks with 2 ox > dsx read their partner's slots) -- parity, cfg5-shaped slab, cfg2 unchanged
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -k "mirrored or streamed or multi_domain_mhd or mhd_single or default_mode" > gpurun_out/r2v_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r2v_pytest.log
timeout -k 10 900 python bench.py --config cfg5 --cells-z 192 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_cfg5.json 2> gpurun_out/r2v_bench_cfg5.err; echo cfg5 rc=$?; tail -2 gpurun_out/r2v_bench_cfg5.err
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo bench rc=$?
ION_EB_FFT_MIRROR=1 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_mirror.json 2> gpurun_out/r2v_bench_mirror.err; echo mirror rc=$?
python - <<'PY'
import json
for f in ("r2v_bench_cfg5","r2v_bench","r2v_bench_mirror"):
    try:
        j=json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(j['value']), round(j['ms_per_step'],3), {k:round(v.get('ms_per_launch', v.get('ms')),3) for k,v in j['kernels'].items()}, j['kernels']['update_e_b_dynamic'].get('static_kernel_spectra_bytes'))
    except Exception as e: print(f, 'ERR', e)
PY
