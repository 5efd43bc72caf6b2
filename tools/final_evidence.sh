#!/bin/bash
# Final-build measurements for profiles/ (run on the B200 box: gpurun -- bash tools/final_evidence.sh [part]).
set -u
OUT=gpurun_out
mkdir -p $OUT
part=${1:-all}
if [ "$part" = "a" ] || [ "$part" = "all" ]; then
  python bench.py --config 1 --steps 5 --warmup 3 > $OUT/fin_bench_config1.json 2> $OUT/fin_bench_config1.err
  cp $OUT/profile_ops_fp32_b32.csv $OUT/fin_per_op_fp32_b32.csv 2>/dev/null
  python bench.py --batch 1 --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/fin_bench_b1_bf16.json 2> $OUT/fin_bench_b1_bf16.err
  cp $OUT/profile_ops_bf16_b1.csv $OUT/fin_per_op_bf16_b1.csv 2>/dev/null
  python bench.py --batch 1 --dtype fp32 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/fin_bench_b1_fp32.json 2> $OUT/fin_bench_b1_fp32.err
  cp $OUT/profile_ops_fp32_b1.csv $OUT/fin_per_op_fp32_b1.csv 2>/dev/null
  python bench.py --batch 2 --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/fin_bench_b2_bf16.json 2> $OUT/fin_bench_b2_bf16.err
fi
if [ "$part" = "b" ] || [ "$part" = "all" ]; then
  python bench.py --config 2 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $OUT/fin_bench_config2.json 2> $OUT/fin_bench_config2.err
  cp $OUT/profile_ops_bf16_b256.csv $OUT/fin_per_op_bf16_b256.csv 2>/dev/null
  python bench.py --config 3 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $OUT/fin_bench_config3.json 2> $OUT/fin_bench_config3.err
fi
if [ "$part" = "c" ] || [ "$part" = "all" ]; then
  for dt in bf16 fp32; do
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file $OUT/fin_ncu_launches_${dt}_b1.csv python tools/ncu_target.py --batch 1 --evals 1 --dtype $dt > $OUT/fin_ncu_${dt}_b1.log 2>&1
    python tools/summarize_ncu.py $OUT/fin_ncu_launches_${dt}_b1.csv $OUT/fin_kernel_summary_${dt}_b1.json >> $OUT/fin_ncu_${dt}_b1.log 2>&1
  done
  # one full capture of the split-K cluster kernel (8 x 10 and 16 x 20 levels of one clip)
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_ks_kernel -s 4 -c 6 -o $OUT/fin_ncu_conv_ks_bf16 -f \
      python tools/ncu_target.py --batch 1 --evals 1 --dtype bf16 > $OUT/fin_ncu_conv_ks.log 2>&1
  python tools/ncu_extract.py $OUT/fin_ncu_conv_ks_bf16.ncu-rep $OUT/fin_ncu_conv_ks_bf16.csv >> $OUT/fin_ncu_conv_ks.log 2>&1
fi
ls -la $OUT | grep fin_ | head -40
