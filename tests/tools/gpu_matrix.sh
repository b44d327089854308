# stream_collide over the configuration matrix + ncu --set full of selected variants. Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r1b}
timeout 600 python tests/tools/kernel_matrix.py > gpurun_out/kernel_matrix_$TAG.jsonl 2> gpurun_out/kernel_matrix.err; echo rc=$?
cat gpurun_out/kernel_matrix_$TAG.jsonl
for sel in "D3Q19 FP16C MHD" "D3Q19 FP16S SRT" "D3Q19 FP32 SRT"; do
  name=$(echo $sel | tr ' ' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide --launch-skip 5 -c 1 -f -o gpurun_out/sc_${name}_$TAG python tests/tools/kernel_matrix.py --only "$sel" --steps 4 > gpurun_out/ncu_$name.log 2>&1; echo ncu $name rc=$?
done
ls -la gpurun_out
