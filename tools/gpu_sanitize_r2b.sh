# compute-sanitizer over the kernels added late in round 2: racecheck (shared-memory hazards) and memcheck on one small parity test each
# -> gpurun_out/r2_sanitize_new_kernels.txt
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
out=gpurun_out/r2_sanitize_new_kernels.txt
: > $out
run() {   # tool, label, test id
  echo "=== $1: $2 ===" >> $out
  if [ "$1" = racecheck ]; then opts="--racecheck-report analysis"; else opts="--leak-check no"; fi
  timeout 900 compute-sanitizer --tool $1 $opts --print-limit 10 python -m pytest "$3" -x -q -p no:cacheprovider > gpurun_out/san_tmp.log 2>&1
  echo "exit code $?" >> $out
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Race reported|hazard|Invalid|out of bounds" gpurun_out/san_tmp.log | sort | uniq -c | head -12 >> $out
}
for tool in racecheck memcheck; do
  run $tool "small-batch decoder (decoder_bf16_sb_kernel<3>), free running 2x37x20" "tests/test_decoder_sb_gpu.py::test_small_batch_kernel_matches_oracle[2-37-20]"
  run $tool "Vocoder_Taco1 fp32 (conv f32, pool, highway FFMA, BiLSTM persistent, Dense 513) 5x7" "tests/test_vocoder_gpu.py::test_vocoder_matches_oracle[fp32-5-7]"
  run $tool "Vocoder_Taco1 tensor-core mode (tcgen05 convs, highway mma, BiLSTM tc) 5x7" "tests/test_vocoder_gpu.py::test_vocoder_matches_oracle[bf16-5-7]"
  run $tool "Griffin-Lim ragged batch (frames FFT, overlap-add, de-emphasis)" "tests/test_vocoder_gpu.py::test_ragged_batch_equals_one_utterance_at_a_time[fp32]"
  run $tool "GST conv stack with the shared tap box (1x188)" "tests/test_gst_gpu.py::test_tensor_core_conv_stack_matches_oracle[1-188]"
done
cat $out
