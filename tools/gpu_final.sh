# end-of-session check: all GPU tests, smoke(), bench (ours + reference arm)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; cat gpurun_out/f_bench.json; tail -3 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/f_bench_ref.json
# launch lists (per-kernel durations; numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-pipeline --no-extras --no-latency --no-cpu-baseline > /dev/null 2>&1; echo "ncu bench rc=$?"
REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file gpurun_out/f_launches_vocoder.csv python tools/bench_vocoder.py 256 1000 bf16 > /dev/null 2>&1; echo "ncu vocoder rc=$?"
timeout 300 python tools/profile_phases_sb.py 1 82 200 > gpurun_out/f_sb_phases.txt 2>&1; timeout 300 python tools/profile_phases_sb.py 8 150 200 >> gpurun_out/f_sb_phases.txt 2>&1
timeout 300 python tools/bench_vocoder.py 256 1000 bf16 > gpurun_out/f_vocoder_bench.txt 2>&1; timeout 300 python tools/bench_encoder.py >> gpurun_out/f_vocoder_bench.txt 2>&1; timeout 300 python tools/bench_gst.py >> gpurun_out/f_vocoder_bench.txt 2>&1
