# round 2, pass e (2 GPUs): distributed parity tests, bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed.py -m gpu -q > gpurun_out/r2e2_pytest_dist.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2e2_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e2_bench_n2.json 2> gpurun_out/r2e2_bench_n2.err; echo bench2 rc=$?; tail -3 gpurun_out/r2e2_bench_n2.err
