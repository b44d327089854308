# round 2: cfg2 weak scaling at N GPUs (N = number of GPUs of the box), plus the distributed parity tests when N == 2
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" = "2" ]; then timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_2gpus.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2_final_pytest_2gpus.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err; echo "bench N=$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_final_ref_n$N.json 2> gpurun_out/r2_final_ref_n$N.err; echo "ref arm N=$N rc=$?"; cut -c1-200 gpurun_out/r2_final_ref_n$N.json
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config cfg4 --steps 10 --warmup 3 > gpurun_out/r2_final_cfg4_n8.json 2> gpurun_out/r2_final_cfg4_n8.err; echo "cfg4 N=8 rc=$?"
fi
