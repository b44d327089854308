# round 2: second 8-GPU pass -- cfg2 weak at N = 8 (chunked ion_buffer_swap in e2e), cfg5 weak at N = 8 with streamed kernel spectra
mkdir -p gpurun_out
run() { # N tag args...
  n=$1; tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
  echo "$tag rc=$?"; tail -c 300 gpurun_out/${tag}.json | head -c 300; echo
}
run 8 r2b_scale_bench_n8 --steps 20 --warmup 5
run 8 r2b_scale_cfg5_n8 --config cfg5 --cells-z 192 --steps 4 --warmup 3
