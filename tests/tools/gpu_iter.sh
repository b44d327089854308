# one iteration on the GPU box: packed-arithmetic check, parity tests, stream_collide matrix, optional ncu captures
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -fmad=false -o /tmp/packed_check tests/tools/packed_check.cu 2>/dev/null && /tmp/packed_check
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/kernel_matrix.py ${MATRIX_ARGS} 2> gpurun_out/km.err | tee gpurun_out/kernel_matrix_iter.jsonl | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print('%-42s ms %.4f  GB/s %7.1f  frac %.3f' % (j['config'][:42], j['ms'], j['achieved_gbs'], j['frac_of_hbm_peak']))"
for sel in ${NCU_SELS}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide --launch-skip 5 -c 1 -f -o gpurun_out/sc_${sel}_${TAG:-iter} python tests/tools/kernel_matrix.py --only "shape $sel D3Q19 FP32 MHD" --shapes $sel --steps 4 > gpurun_out/ncu_$sel.log 2>&1; echo ncu $sel rc=$?
done
