# compute-sanitizer over the kernels added after tools/gpu_sanitize_r2b.sh ran: the cluster BiLSTM (distributed shared memory, barrier.cluster),
# the two-n-tile small-batch decoder, the tcgen05 value projection  -> gpurun_out/r2_sanitize_cluster_kernels.txt
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
out=gpurun_out/r2_sanitize_cluster_kernels.txt
: > $out
run() {
  echo "=== $1: $2 ===" >> $out
  if [ "$1" = racecheck ]; then opts="--racecheck-report analysis"; else opts="--leak-check no"; fi
  timeout 900 compute-sanitizer --tool $1 $opts --print-limit 10 python -m pytest $3 -x -q -p no:cacheprovider > gpurun_out/san_tmp.log 2>&1
  echo "exit code $?" >> $out
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|Race reported|hazard|Invalid|out of bounds" gpurun_out/san_tmp.log | sort | uniq -c | head -12 >> $out
}
for tool in racecheck memcheck; do
  run $tool "Encoder, tensor-core mode (cluster BiLSTM, fp16 gate pre-activations)" "tests/test_encoder_gpu.py -k ragged_sizes"
  run $tool "Vocoder_Taco1, tensor-core mode 5x7 (cluster BiLSTM)" "tests/test_vocoder_gpu.py::test_vocoder_matches_oracle[bf16-5-7]"
  run $tool "small-batch decoder, two n-tiles (12 x 150 x 6) + tcgen05 value projection" "tests/test_decoder_sb_gpu.py::test_small_batch_kernel_matches_oracle[12-150-6]"
done
cat $out
