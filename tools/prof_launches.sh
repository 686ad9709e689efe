#!/bin/bash
# ncu launch lists (gpu__time_duration.sum) of one training step of each model, after the plain run exited 0
mkdir -p gpurun_out
for m in vit cnn; do
  b=64; [ $m = cnn ] && b=128
  python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 2 > gpurun_out/plain_$m.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/r01b_${m}_train_launches.csv \
      python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 2 > gpurun_out/ncu_$m.log 2>&1
done
