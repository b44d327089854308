# round 2, final 8-GPU pass: cfg2 weak, cfg4 strong, cfg5 weak (streamed kernel spectra) at N = 8
mkdir -p gpurun_out
run() { # N tag args...
  n=$1; tag=$2; shift 2
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
  echo "$tag rc=$?"; python -c "
import json,sys
j=json.load(open('gpurun_out/${tag}.json')); print(j['value'], j['ms_per_step'], {k:v.get('ms_per_launch', v.get('ms')) for k,v in j['kernels'].items()}, j['e2e'].get('value'))"
}
run 8 r2_final_bench_n8 --steps 20 --warmup 5
run 8 r2_final_cfg4_n8 --config cfg4 --steps 10 --warmup 3
run 8 r2_final_cfg5_n8 --config cfg5 --cells-z 192 --steps 4 --warmup 3
