#!/bin/bash
# round 2, final evidence (run under gpurun): launch lists + step shares of both training steps at the end of the round and a
# full capture of the forward attention kernel in class-token-outside-the-tiles mode
TAG=${1:-r02m}
mkdir -p gpurun_out
bash tools/prof_launches.sh $TAG
POSE_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attn_fwd_tc_kernel<.int.64" -s 4 -c 1 \
    --profile-from-start off -o gpurun_out/${TAG}_attn_fwd_tail \
    python tools/bench_train.py --model vit --batch 64 --steps 1 --warmup 3 --cuda-profiler > gpurun_out/ncu_attn_fwd_tail.log 2>&1
ncu -i gpurun_out/${TAG}_attn_fwd_tail.ncu-rep --page raw --csv > gpurun_out/${TAG}_attn_fwd_tail_full_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_attn_fwd_tail.ncu-rep
ls -la gpurun_out/${TAG}_*
