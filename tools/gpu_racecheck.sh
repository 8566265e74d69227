# compute-sanitizer racecheck (shared-memory hazards) over one small parity test per persistent kernel; summary -> gpurun_out/r2_racecheck.txt
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
out=gpurun_out/r2_racecheck.txt
: > $out
run() {
  echo "=== $1 ===" >> $out
  GSTK_DECODER=$2 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 \
    python -m pytest "$3" -x -q -p no:cacheprovider > gpurun_out/rc_tmp.log 2>&1
  echo "exit code $?" >> $out
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Race reported|hazard" gpurun_out/rc_tmp.log | sort | uniq -c | head -20 >> $out
}
run "bf16 barrier decoder (decoder_bf16_kernel), free running 3x37x20" barrier "tests/test_decoder_v2_gpu.py::test_v2_free_running_external_randomness_matches_oracle[barrier-3-37-20]"
run "bf16 dataflow decoder (decoder_bf16_v2_kernel), free running 3x37x20" dataflow "tests/test_decoder_v2_gpu.py::test_v2_free_running_external_randomness_matches_oracle[dataflow-3-37-20]"
run "fp32 decoder + GST kernels (golden sma_r1)" barrier "tests/test_golden_gpu.py::test_decoder_against_reference_goldens"
run "GST tensor-core conv stack (1x188)" barrier "tests/test_gst_gpu.py::test_tensor_core_conv_stack_matches_oracle[1-188]"
cat $out
