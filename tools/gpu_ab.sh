cd $GRAFT_REPO_ROOT
for dbg in 0 4 1; do echo "=== GSTK_DEBUG=$dbg"; GSTK_DEBUG=$dbg timeout 300 python tools/profile_phases.py 256 150 200 2>&1 | head -9; done
echo "=== B=128"; timeout 300 python tools/profile_phases.py 128 150 200 2>&1 | head -9
