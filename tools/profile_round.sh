#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): launch lists and one `ncu --set full` capture per kernel.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r01b "hpel lowres residual"'
# $1 = tag for the output names, $2 = kernels (xd_<name>) to capture in full (default: all).
# Each ncu command only runs after the same command line exited 0 without ncu (B200_PROFILING.md).
TAG=${1:-r01}
KERNELS=${2:-"hpel_kernel lowres_kernel residual_kernel mc_frame_kernel me_sized_kernel me_search_kernel deblock_kernel deblock_strength_kernel la_multi_kernel la_intra_kernel load_i420_kernel expand_border_kernel filtered_border_kernel"}
OUT=gpurun_out
mkdir -p $OUT
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv"
NCU_FULL="ncu --set full --clock-control none --import-source on"

B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-me"
P="python tools/bench_paths.py --pairs 4 --reps 1 --no-oracle"

$B > $OUT/${TAG}_bench_plain.log 2>&1 &&
$NCU_LIST --log-file $OUT/${TAG}_bench_launches.csv $B > $OUT/${TAG}_bench_ncu.log 2>&1
$P > $OUT/${TAG}_paths_plain.log 2>&1 &&
$NCU_LIST --log-file $OUT/${TAG}_paths_launches.csv $P > $OUT/${TAG}_paths_ncu.log 2>&1
for k in $KERNELS; do
    CMD="$P"
    [ "$k" = "la_inter_kernel" ] && CMD="$B"
    [ "$k" = "la_multi_kernel" ] && CMD="$B"
    [ "$k" = "la_intra_kernel" ] && CMD="$B"
    # skip the warm-up launch of each kernel so that the capture is a steady-state one
    $NCU_FULL -k regex:xd_$k -s 1 -c 1 -f -o $OUT/${TAG}_full_$k $CMD > $OUT/${TAG}_full_$k.log 2>&1
    tail -2 $OUT/${TAG}_full_$k.log
done
ls -la $OUT | tail -40
