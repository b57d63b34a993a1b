#!/bin/bash
# ncu evidence for config 4 (MLP Hamiltonian, DMMA tile GEMMs), run under gpurun on one B200 after the plain run exits 0.
#   bash profiles/capture_r02_c4.sh   -> gpurun_out/r02c4_*
set -x
OUT=gpurun_out
python bench.py --workload c4 --M 262144 --steps 1 --warmup 2 --no-cpu-baseline > $OUT/r02c4_plain_run.json 2> $OUT/r02c4_plain_run.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/r02c4_launches.csv python bench.py --workload c4 --M 262144 --steps 1 --warmup 2 --no-cpu-baseline > $OUT/r02c4_ncu_launches.log 2>&1
for K in mlp_grad_kernel mlp_hess_kernel lu16_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" --launch-skip 6 --launch-count 1 \
      -f -o $OUT/r02c4_$K python bench.py --workload c4 --M 262144 --steps 1 --warmup 2 --no-cpu-baseline > $OUT/r02c4_ncu_full_$K.log 2>&1
  ncu -i $OUT/r02c4_$K.ncu-rep --page details > $OUT/r02c4_ncu_full_$K.txt 2>&1
  ncu -i $OUT/r02c4_$K.ncu-rep --page raw --csv > $OUT/r02c4_ncu_raw_$K.csv 2>&1
  ncu -i $OUT/r02c4_$K.ncu-rep --page source --csv > $OUT/r02c4_ncu_source_$K.csv 2>&1
  rm -f $OUT/r02c4_$K.ncu-rep
done
ls -la $OUT | tail
