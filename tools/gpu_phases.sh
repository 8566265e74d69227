cd $GRAFT_REPO_ROOT
timeout 300 python tools/profile_phases.py 256 150 200 > /tmp/ph.txt 2>&1; grep -v "A1:" /tmp/ph.txt | head -12
timeout 300 python tools/profile_phases.py 1 82 200 2>&1 | head -1
