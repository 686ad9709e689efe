#!/bin/bash
# ncu launch lists (gpu__time_duration.sum) of ONE training step of each model (eager launches: POSE_TRAIN_GRAPH=0, the
# per-kernel durations are the same as inside the graph), after the plain run exited 0.  usage: prof_launches.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
for m in vit cnn; do
  b=64; [ $m = cnn ] && b=128
  POSE_TRAIN_GRAPH=0 python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 3 > gpurun_out/plain_$m.log 2>&1 || exit 1
  POSE_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/${TAG}_${m}_train_launches.csv \
      python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 3 --cuda-profiler > gpurun_out/ncu_$m.log 2>&1
  python tools/step_share.py gpurun_out/${TAG}_${m}_train_launches.csv "$m training step ($TAG)" > gpurun_out/${TAG}_${m}_train_step_share.md
done
