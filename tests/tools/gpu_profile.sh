# ncu --set full captures of the two step kernels + the FP32 issue-peak micro-benchmark. Outputs under gpurun_out/.
mkdir -p gpurun_out
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_peak tests/tools/fp32_peak.cu && /tmp/fp32_peak > gpurun_out/fp32_peak.jsonl; cat gpurun_out/fp32_peak.jsonl
TAG=${TAG:-r1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide --launch-skip 4 -c 1 -f -o gpurun_out/sc_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sc.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update_e_b --launch-skip 4 -c 1 -f -o gpurun_out/eb_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_eb.log 2>&1; echo rc=$?
ls -la gpurun_out
