#!/bin/bash
# ncu evidence of round 2 (run under gpurun on one B200; every profiled command is first run WITHOUT ncu).
#   bash profiles/capture_r02.sh        -> gpurun_out/r02_*   (summaries are copied into profiles/ afterwards)
set -x
OUT=gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02_ncu_plain_run.json 2> $OUT/r02_ncu_plain_run.err || exit 1
# (1) launch list of one bench command: per-launch duration + DRAM bytes
[ -n "$SKIP_LAUNCH_LIST" ] || ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/r02_launches_bench_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02_ncu_launches.log 2>&1
# (2) full captures of the three epoch kernels at their heaviest epochs (2^18 rows keep the replay short)
for spec in "3:10" "5:22" "7:12"; do
  S=${spec%%:*}; SKIP=${spec##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:epoch_kernel<jb::HamVdp, \(int\)${S}," --launch-skip $SKIP --launch-count 1 -f -o $OUT/r02_epoch_kernel_s${S} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --M 262144 > $OUT/r02_ncu_full_s${S}.log 2>&1
  ncu -i $OUT/r02_epoch_kernel_s${S}.ncu-rep --page details > $OUT/r02_ncu_full_epoch_kernel_s${S}.txt 2>&1
  ncu -i $OUT/r02_epoch_kernel_s${S}.ncu-rep --page raw --csv > $OUT/r02_ncu_raw_epoch_kernel_s${S}.csv 2>&1
done
ls -la $OUT | tail -20
