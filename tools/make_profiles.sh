#!/bin/bash
# Run on the GPU box (via gpurun): ncu launch list of the bench command + full capture of the
# dominant forward kernel + the bench lines themselves.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) full capture of the dominant kernel (finest level, rotating buffers), 2 launches
ncu --set full --clock-control none --import-source on -k regex:warp_corr_fwd -s 15 -c 2 \
    -o gpurun_out/prof_fwd_l4_${TAG} python tools/profile_level.py 4 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
# (3) the bench lines proper (never under a profiler)
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json; echo; tail -c 300 gpurun_out/bench_ref_${TAG}.json; echo
