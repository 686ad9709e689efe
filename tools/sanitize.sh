#!/bin/bash
# compute-sanitizer pass over the tiny-shape GPU tests (run under gpurun).  memcheck on everything selected below;
# racecheck + synccheck on the kernels that use clusters / DSMEM (augment), mbarrier pipelines and TMEM alloc / dealloc
# (GEMM, convolution, attention).  Summaries -> gpurun_out/sanitize_*.log (copied to profiles/ by hand).
# usage: bash tools/sanitize.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
# small shapes only: the sanitizer serialises and instruments every access (10-100x slower)
SEL_SMALL='test_loss_matches_oracle or test_loss_other_joint_counts or (test_heatmap_matches_oracle and not 500) or (test_augment_matches_oracle_all_stages and 64-64-16) or test_infer_prep or test_collator or test_eval_metrics'
SEL_TC='(test_gemm_tcgen05_matches_fp32_reference and (128-128-64 or 1-51-512)) or test_regression_head_training_mode or (test_conv_epilogue_emits_batchnorm_statistics and 4-64-64-64-3) or (test_gemm_epilogue_emits_batchnorm and 300-32-64) or (test_attention_forward_backward and 16-256)'
run() {  # tool selection name
  timeout 1500 $SAN --tool $1 --error-exitcode 77 --print-limit 20 --launch-timeout 120 \
      python -m pytest tests -x -q -m gpu -k "$2" -p no:cacheprovider > $OUT/sanitize_${TAG}_$3.log 2>&1
  echo "$3 rc=$?" | tee -a $OUT/sanitize_${TAG}_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $OUT/sanitize_${TAG}_$3.log | tail -5 | tee -a $OUT/sanitize_${TAG}_summary.txt
}
rm -f $OUT/sanitize_${TAG}_summary.txt
run memcheck "$SEL_SMALL" memcheck_small
run memcheck "$SEL_TC" memcheck_tc
run racecheck "(test_augment_matches_oracle_all_stages and 64-64-16) or test_loss_matches_oracle" racecheck_small
run racecheck "$SEL_TC" racecheck_tc
run synccheck "(test_augment_matches_oracle_all_stages and 64-64-16) or $SEL_TC" synccheck
