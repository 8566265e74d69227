# quick GPU check of the bf16 decoder: parity tests, per-phase timers, optional ncu capture (NCU=1)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or golden" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/q_pytest.log
timeout 300 python tools/profile_phases.py 256 150 200 > gpurun_out/q_phases.txt 2>&1; cat gpurun_out/q_phases.txt
timeout 300 python tools/profile_phases.py 1 82 200 2>&1 | head -3
if [ "$NCU" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_bf16_kernel -s 1 -c 1 -o gpurun_out/q_decoder_bf16 -f python tools/profile_phases.py 256 150 100 > gpurun_out/q_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
