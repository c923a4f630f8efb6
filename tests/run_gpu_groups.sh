#!/bin/bash
# Runs the GPU kernel tests group by group, each under its own timeout, so one hung kernel cannot eat the box.
# usage: tests/run_gpu_groups.sh [outdir]
out=${1:-gpurun_out}
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/smi.txt" 2>&1
i=0
for k in "gemm_linear" "gemm_column or gemm_epilogue or gemm_n_store" "gemm_geglu" "gemm_lora" "gemm_conv3x3" \
         "gemm_tconv3" "gemm_large" "groupnorm" "layernorm" "attention_self" "attention_cross" "attention_svd" \
         "attention_temporal" "small_linear or pack_unpack or upsample or cfg_euler or fp32_residual"; do
  i=$((i+1))
  timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" -p no:cacheprovider > "$out/group_$i.log" 2>&1
  echo "group $i [$k] exit $?" | tee -a "$out/summary.txt"
  tail -n 3 "$out/group_$i.log" | tee -a "$out/summary.txt"
done
