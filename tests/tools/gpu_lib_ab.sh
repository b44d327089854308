# A/B of library builds (ION_LIB) over the stream_collide kernel matrix; parity tests on the default build first.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
for lib in "" ${LIBS}; do
  echo "== lib=${lib:-default}"
  ION_LIB=${lib:+$PWD/ionsolver_b200/$lib} timeout 600 python tests/tools/kernel_matrix.py ${MATRIX_ARGS} 2> gpurun_out/km.err | tee gpurun_out/kernel_matrix_${lib:-default}.jsonl | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print('%-42s ms %.4f  GB/s %7.1f  frac %.3f' % (j['config'][:42], j['ms'], j['achieved_gbs'], j['frac_of_hbm_peak']))"
done
