#!/bin/bash
# Run on the GPU box (via gpurun): ncu launch list of the bench command, full captures of the dominant forward kernel
# (tensor-core finest level) and of the fused backward kernel, compute-sanitizer logs, and the bench lines themselves.
# Outputs land in gpurun_out/;  tools/summarize_profiles.py turns them into profiles/<round>_*.txt.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-train --no-other-configs \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) full capture of the dominant kernel (finest level, both flow directions, rotating buffers), 2 launches
ncu --set full --clock-control none --import-source on -k regex:warp_corr_fwd -s 15 -c 2 \
    -o gpurun_out/prof_fwd_l4_${TAG} python tools/profile_level.py 4 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
# (3) full capture of the fused backward kernel (HRNet training level, batch 8)
ncu --set full --clock-control none --import-source on -k regex:corr_bwd_fused -s 1 -c 1 \
    -o gpurun_out/prof_bwd_${TAG} python tools/profile_backward.py > gpurun_out/ncu_bwd_${TAG}.log 2>&1
# (3b) full captures of the tensor-core backward kernel and of the shared-memory-window splat kernel (same shape)
ncu --set full --clock-control none --import-source on -k regex:corr_bwd_tc -s 1 -c 1 \
    -o gpurun_out/prof_bwd_tc_${TAG} python tools/profile_backward_tc.py > gpurun_out/ncu_bwd_tc_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flow_warp_bwd_box -s 1 -c 1 \
    -o gpurun_out/prof_splat_${TAG} python tools/profile_backward_tc.py > gpurun_out/ncu_splat_${TAG}.log 2>&1
# (4) compute-sanitizer over cases that touch every kernel path
for tool in memcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/sanitizer_${tool}_${TAG}.txt 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_${TAG}.txt
done
timeout 420 compute-sanitizer --tool racecheck python tools/race_cluster_cases.py > gpurun_out/sanitizer_racecheck_${TAG}.txt 2>&1
tail -3 gpurun_out/sanitizer_racecheck_${TAG}.txt
# (5) the bench lines proper (never under a profiler)
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json; echo; tail -c 300 gpurun_out/bench_ref_${TAG}.json; echo
