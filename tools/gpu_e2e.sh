cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "bf16" 2>&1 | tail -3
for n in 4 8; do echo "chunks=$n"; GSTK_TCHUNKS=$n timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['us_per_decoder_step'])"; done
