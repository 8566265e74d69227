# end-of-session check: all GPU tests, smoke(), bench (ours + reference arm)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; cat gpurun_out/f_bench.json; tail -3 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/f_bench_ref.json
