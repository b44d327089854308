# Round-end evidence on one GPU: bench line (+ LOD depth 3), ncu launch list of the same command, ncu --set full of the two step
# kernels, stream_collide matrix, whole-step numbers of the other BASELINE configs.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r1d}
nvidia-smi -L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench rc=$?
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --steps 20 --warmup 3 --lod-depth 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_lod3.json 2>> gpurun_out/bench_$TAG.err; echo bench3 rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo ncu launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide --launch-skip 4 -c 1 -f -o gpurun_out/sc_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sc.log 2>&1; echo ncu sc rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update_e_b --launch-skip 4 -c 1 -f -o gpurun_out/eb_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_eb.log 2>&1; echo ncu eb rc=$?
timeout 600 python tests/tools/kernel_matrix.py --shapes 512x512x512 > gpurun_out/kernel_matrix_$TAG.jsonl 2> gpurun_out/km.err; echo matrix rc=$?
cat gpurun_out/kernel_matrix_$TAG.jsonl
if [ -n "$CONFIGS" ]; then timeout 1500 python tests/tools/bench_configs.py > gpurun_out/bench_configs_$TAG.jsonl 2> gpurun_out/bench_configs.err; echo configs rc=$?; cat gpurun_out/bench_configs_$TAG.jsonl; fi
