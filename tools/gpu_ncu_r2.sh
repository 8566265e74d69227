# round-2 ncu evidence: launch list of the bench command + --set full capture of the persistent decoder on the bench workload
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_bf16.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
GSTK_DECODER=barrier timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_bf16_kernel -s 1 -c 1 -o gpurun_out/r2_decoder_bf16 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2_decoder_bf16.ncu-rep --page raw --csv > gpurun_out/r2_decoder_bf16_raw.csv 2>/dev/null
GSTK_COMMIT=${GSTK_COMMIT:-unknown} python tools/ncu_traffic.py gpurun_out/r2_decoder_bf16_raw.csv gpurun_out/r2_ncu_traffic.json gpurun_out/r2_ncu_full_decoder_bf16.csv
python tools/summarize_launches.py gpurun_out/r2_launches_bench_bf16.csv > gpurun_out/r2_launches_summary.md 2>&1; head -20 gpurun_out/r2_launches_summary.md
