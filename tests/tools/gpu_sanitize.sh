# compute-sanitizer memcheck + racecheck on small scenes (single domain MHD depth 3/4, split MHD with the overlapped halo schedule)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "test_mhd_single_step_parity or test_multi_domain_mhd or ecr_d3q19 or test_single_domain_bit_exact and d3q19_fp16c" > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -4
done
