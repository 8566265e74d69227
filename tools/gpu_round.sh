# full GPU round: all GPU tests, bench (both arms), per-phase timers, ncu launch list + full capture of the decoder on the bench workload
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r_bench_ref.json 2> gpurun_out/r_bench_ref.err; echo "bench ref rc=$?"
timeout 300 python tools/profile_phases.py 256 150 200 > gpurun_out/r_phases.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/r_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_bf16_kernel -s 1 -c 1 -o gpurun_out/r_decoder_bf16 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/r_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/r_pytest_gpu.log; cat gpurun_out/r_bench.json; cat gpurun_out/r_bench_ref.json; cat gpurun_out/r_phases.txt
