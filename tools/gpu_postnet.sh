# Postnet parity tests + timing + ncu launch list (one short GPU call)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_postnet_gpu.py -x -q -s > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/p_pytest.log
timeout 300 python tools/bench_postnet.py 256 1000 5 > gpurun_out/p_bench.txt 2>&1; cat gpurun_out/p_bench.txt
GSTK_POSTNET_TC=0 timeout 300 python tools/bench_postnet.py 256 1000 5 2>&1 | head -1 | sed 's/^/mma.sync only: /' | tee -a gpurun_out/p_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/p_launches.csv python tools/bench_postnet.py 256 1000 1 > gpurun_out/p_ncu.log 2>&1; echo "ncu rc=$?"
grep postnet gpurun_out/p_launches.csv | head -8 | cut -d, -f5,10-
if [ "$NCU" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:postnet_conv_tc -s 1 -c 1 -o gpurun_out/p_postnet_tc -f python tools/bench_postnet.py 256 1000 1 > gpurun_out/p_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
