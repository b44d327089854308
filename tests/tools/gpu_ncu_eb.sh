# ncu --set full of update_e_b for the variants in $VARIANTS
mkdir -p gpurun_out
for v in ${VARIANTS:-2 1}; do
ION_EB_VARIANT=$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update_e_b --launch-skip 4 -c 1 -f -o gpurun_out/eb_v$v python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lod-depth ${DEPTH:-4} > gpurun_out/ncu_eb_v$v.log 2>&1; echo rc=$?
done
ls -la gpurun_out/*.ncu-rep
