cd $GRAFT_REPO_ROOT
export PLSLAM_LSD_WATCHDOG_MS=3000
echo "== default"; timeout 100 python tools/lsd_check.py --frames 20 --batch 20 --repeat 3 2>&1| grep -v "frame 9:\|done$" | tail -4
for w in 1024 2048; do echo "== slots $w"; PLSLAM_LSD_SLOTS=$w timeout 100 python tools/prof_line.py --frames 1 --chunk 1 --iters 10 2>&1 | grep -A2 "grow kernel" ; done
timeout 100 python tools/prof_line.py --frames 300 --chunk 300 --room 2>&1 | tail -9
PLSLAM_LSD_SHAPE=few timeout 100 python tools/prof_line.py --frames 300 --chunk 300 --room 2>&1 | tail -9
