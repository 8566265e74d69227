cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GSTK_DEBUG=${DBG:-0} timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_bf16_kernel -s 1 -c 1 -o gpurun_out/q_decoder_bf16 -f python tools/profile_phases.py 256 150 100 > gpurun_out/q_ncu_full.log 2>&1; echo "ncu full rc=$?"
