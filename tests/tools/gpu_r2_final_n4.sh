# round 2, final build: cfg2 weak at N = 4
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n4.json 2> gpurun_out/r2_final_bench_n4.err; echo "n4 rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/r2_final_bench_n4.json')); print(j['value'], j['ms_per_step'], {k:v.get('ms_per_launch', v.get('ms')) for k,v in j['kernels'].items()}, j['e2e'].get('value'))"
