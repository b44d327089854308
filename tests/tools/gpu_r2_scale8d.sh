# round 2: cfg5 weak at N = 8 with streamed + mirrored kernel spectra (final build)
mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config cfg5 --cells-z 192 --steps 3 --warmup 3 > gpurun_out/r2_final_cfg5_n8_mirror.json 2> gpurun_out/r2_final_cfg5_n8_mirror.err
echo "rc=$?"; python -c "
import json
j=json.load(open('gpurun_out/r2_final_cfg5_n8_mirror.json')); print(j['value'], j['ms_per_step'], {k:v.get('ms_per_launch', v.get('ms')) for k,v in j['kernels'].items()})"
