#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU): launch lists of a C3 / C4 / C5 step (python tools/ncu_agg.py <csv>
# aggregates them by kernel), per-launch DRAM traffic of the GEMM
# launches of that step, and --set full captures of the dominant GEMM / conv / attention launches.
# usage: tools/ncu_round.sh <tag>      -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches_c3.csv $B > $out/${tag}_launches_c3.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm_tcgen05 --csv --log-file $out/${tag}_gemm_traffic_c3.csv $B \
    > $out/${tag}_gemm_traffic_c3.log 2>&1
B4="python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu-baseline --profiler-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches_c4.csv $B4 > $out/${tag}_launches_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches_vae_decode8.csv python tools/bench_vae.py 8 --profiler > $out/${tag}_launches_vae.log 2>&1
B5="python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --profiler-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches_c5.csv $B5 > $out/${tag}_launches_c5.log 2>&1     # one eager training step
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"gn_bwd|small_linear_bwd|colsum|geglu|ln_bwd|gn_apply4" --csv \
    --log-file $out/${tag}_train_small_kernels.csv python tools/bench_train_small.py > /dev/null 2>&1
for spec in "0 gemm_ff1_L0_geglu" "11 gemm_conv_L0_res32" "3 gemm_proj_L0_res32"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f \
      -o $out/${tag}_$2 python tools/bench_gemm.py --only $1 --iters 1 > $out/${tag}_$2.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 2 -c 1 -f \
    -o $out/${tag}_attn_L0 python tools/bench_attn.py 0 > $out/${tag}_attn_L0.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 2 -c 1 -f \
    -o $out/${tag}_gn_apply_L0 python tools/bench_norm.py > $out/${tag}_gn_apply_L0.log 2>&1
ls -la $out
