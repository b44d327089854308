# multi-GPU round (gpurun --gpus N): distributed parity tests + weak-scaling bench at N GPUs
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --lod-depth 3 > gpurun_out/bench_n${N}_lod3.json 2>> gpurun_out/bench_n$N.err; echo bench3 rc=$?
cat gpurun_out/bench_n${N}_lod3.json
