#!/bin/bash
# Profiling recipe used for profiles/ (run on the GPU box through gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   $1 = tag (e.g. r01_v3), $2... = extra bench.py args; NSKIP / NLIST / NCAP = launches to skip, list, capture (one cfg2 step = 13 kernels)
# 1. plain run (must exit 0)  2. launch list with device times  3. one --set full capture of one step's kernels
TAG=$1; shift
K='regex:^(zero_bytes|copy_bytes|stft1024|ns_scan|ns_mask_smooth|smooth_mask|istft|kweight_energy|peak_factor|gate|gate_grid|logmel|mel_finalize|noise_pass|noise_finalize|frame_stats|select_clip|fill_int|gather_lufs|apply_gain)_kernel'
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s ${NSKIP:-39} -c ${NLIST:-26} --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s ${NSKIP:-39} -c ${NCAP:-13} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-1200
