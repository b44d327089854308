# A/B timing of update_e_b_dynamic variants (ION_EB_VARIANT) + parity tests on the default
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-1 2 3 4 0}; do
  for d in ${DEPTHS:-4 3}; do
  ION_EB_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --lod-depth $d --no-cpu-baseline 2> gpurun_out/bench_v$v.err | python -c "
import json,sys
j=json.load(sys.stdin)
print('variant $v lod $d MLUPs', round(j['value'],1), 'ms/step', round(j['ms_per_step'],3), {k:(round(x['ms'],3)) for k,x in j['kernels'].items()}, 'pairs/s %.3e' % j['kernels']['update_e_b_dynamic']['pairs_per_s'])"
  done
done
