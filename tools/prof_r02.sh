#!/bin/bash
# round 2 (run under gpurun): per-launch trace + ncu launch list of the CNN step, full captures of the convolution kernels
TAG=${1:-r02a}
mkdir -p gpurun_out
python tools/trace_plan.py --model cnn --batch 128 --top 90 > gpurun_out/${TAG}_cnn_trace.txt 2>&1
cap() {  # name regex model batch skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $5 -c $6 -o gpurun_out/${TAG}_$1 \
      python tools/bench_train.py --model $3 --batch $4 --steps 1 --warmup 1 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_full_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
# conv1.1 forward / data gradient (3x3 64 -> 64 at 128^2): gemm<64,6,64,MODE 1>; conv weight gradients: gemm<128,4,64,MODE 2>
cap conv_fwd64 'gemm_bf16_tn_kernel<.int.64, .int.6, .int.64, .int.1' cnn 128 1 1
cap conv_wgrad 'gemm_bf16_tn_kernel<.int.128, .int.4, .int.64, .int.2' cnn 128 0 2
cap bn_bwd_reduce 'bn_bwd_reduce_kernel<.int.2' cnn 128 40 1
ls -la gpurun_out/${TAG}_*
