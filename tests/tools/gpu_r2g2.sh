# round 2, pass g (2 GPUs): all parity tests incl. the 2-GPU ones, bench at N = 2
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2g2_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2g2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g2_bench_n2.json 2> gpurun_out/r2g2_bench_n2.err; echo bench2 rc=$?; tail -2 gpurun_out/r2g2_bench_n2.err
