# round 2, final 2-GPU pass: the whole -m gpu suite on a 2-GPU box (nothing skipped), cfg2 weak at N = 2
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_2gpus.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2_final_pytest_2gpus.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err; echo "n2 rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/r2_final_bench_n2.json')); print(j['value'], j['ms_per_step'], {k:v.get('ms_per_launch', v.get('ms')) for k,v in j['kernels'].items()}, j['e2e'].get('value'))"
