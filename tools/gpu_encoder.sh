cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -s > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/e_pytest.log
timeout 300 python tools/bench_encoder.py 256 150 5 > gpurun_out/e_bench.txt 2>&1; cat gpurun_out/e_bench.txt
timeout 300 python tools/bench_encoder.py 1 82 5 2>&1 | tee -a gpurun_out/e_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/e_launches.csv python tools/bench_encoder.py 256 150 1 > gpurun_out/e_ncu.log 2>&1; echo "ncu rc=$?"
grep "encoder\|postnet" gpurun_out/e_launches.csv | head -14 | awk -F'","' '{print $5, $(NF)}'
