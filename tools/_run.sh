cd $GRAFT_REPO_ROOT/_tpr
export PLSLAM_LSD_WATCHDOG_MS=3000
(cd oracle && make -s 2>&1 | tail -2)
echo "== check"; timeout 200 python tools/lsd_check.py --frames 20 --batch 20 2>&1 | tail -3
echo "== 1 frame"; timeout 100 python tools/prof_line.py --frames 1 --chunk 1 --iters 10 2>&1 | tail -5
echo "== 300 room"; timeout 100 python tools/prof_line.py --frames 300 --chunk 300 --room 2>&1 | tail -12
echo "== 300 room few"; PLSLAM_LSD_SHAPE=few timeout 100 python tools/prof_line.py --frames 300 --chunk 300 --room 2>&1 | tail -5
