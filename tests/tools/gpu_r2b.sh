# round 2, second GPU pass: debug of the FP16C z-split deterministic case, parity tests (continue past failures), bench, ncu
mkdir -p gpurun_out
timeout 300 python tests/tools/debug_det.py mhd_z2_d3q19_fp16c_lod4 8 > gpurun_out/r2b_debug_det.log 2>&1; head -40 gpurun_out/r2b_debug_det.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo bench rc=$?; cat gpurun_out/r2b_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step']); print({k:v.get('ms') for k,v in j['kernels'].items()}); print(j['e2e'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eb_fft -s 3 -c 1 -f -o gpurun_out/r2b_eb_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1; echo ncu rc=$?
timeout 900 python tests/tools/drift_curve.py --steps 60 > gpurun_out/r2b_drift.jsonl 2> gpurun_out/r2b_drift.err; echo drift rc=$?; tail -3 gpurun_out/r2b_drift.err
