# A/B timing of stream_collide: block size cap (ION_SC_BLOCK) x DDF cache hint build (ION_LIB)
mkdir -p gpurun_out
for lib in "" "$PWD/ionsolver_b200/libionsolver_b200_cs.so"; do
for b in ${BLOCKS:-256 128 64}; do
  ION_LIB=$lib ION_SC_BLOCK=$b timeout 600 python bench.py --steps 20 --warmup 3 --lod-depth 3 --no-cpu-baseline 2> gpurun_out/sc_ab.err | python -c "
import json,sys
j=json.load(sys.stdin)
k=j['kernels']['stream_collide']
print('lib=${lib##*/} block=$b stream_collide ms %.4f GB/s %.0f frac %.3f' % (k['ms'], k['achieved_gbs'], k['frac_of_hbm_peak']))"
done
done
