#!/bin/bash
# Captures the ncu evidence for profiles/: (1) the launch list of a short bench run, (2) --set full of the top kernels.
# Run under gpurun:  gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh'
set -u
mkdir -p gpurun_out
BENCH="python bench.py --frames 300 --steps 1 --warmup 1 --cpu-sample 4 --kitti-frames 0 --c5-frames 0"
$BENCH > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
LINE="python tools/prof_line.py --room --frames 300 --chunk 300 --iters 1"
$LINE > gpurun_out/line_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_lsd_grow2|k_lsd_nfa|k_lsd_scale|k_lsd_grad|k_lsd_scatter|k_line_finalize|k_blur5_sobel3_tma|k_lbd_rows|k_lbd_finish' -c 9 -o gpurun_out/prof_line -f $LINE > gpurun_out/ncu_line.log 2>&1
echo "line full rc=$?"
ORB="python tools/prof_orb.py --frames 300 --chunk 300 --iters 1"
$ORB > gpurun_out/orb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pyr_resize|k_fast_cells|k_blur7|k_orient_brief|k_octree|k_pyr_base' -c 12 -o gpurun_out/prof_orb -f $ORB > gpurun_out/ncu_orb.log 2>&1
echo "orb full rc=$?"
ls -la gpurun_out
