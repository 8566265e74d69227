cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gst or golden or attention" 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/g_launches.csv
