# 8-GPU round (gpurun --gpus 8): cfg2 weak scaling (bench.py), cfg4 strong scaling, cfg5 weak scaling
N=${N:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?
cat gpurun_out/bench_n$N.json | cut -c1-400
timeout 600 $TR --master-port 29522 tests/tools/bench_scaling.py --config cfg4 > gpurun_out/cfg4_n$N.json 2> gpurun_out/cfg4_n$N.err; echo cfg4 rc=$?; cat gpurun_out/cfg4_n$N.json; tail -3 gpurun_out/cfg4_n$N.err
timeout 900 $TR --master-port 29523 tests/tools/bench_scaling.py --config cfg5 --steps 3 > gpurun_out/cfg5_n$N.json 2> gpurun_out/cfg5_n$N.err; echo cfg5 rc=$?; cat gpurun_out/cfg5_n$N.json; tail -3 gpurun_out/cfg5_n$N.err
