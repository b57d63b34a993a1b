#!/bin/bash
# Round-2 final evidence (run under gpurun on one B200):  bash profiles/capture_r02_final.sh  -> gpurun_out/r02f_*
set -x
OUT=gpurun_out
( time python -m pytest tests -m gpu -q ) > $OUT/r02f_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02f_smoke.log 2>&1
python bench.py > $OUT/r02f_bench_default.json 2> $OUT/r02f_bench_default.err
python bench.py --fixed-order --no-secondary --no-cpu-baseline > $OUT/r02f_bench_fixed.json 2> $OUT/r02f_bench_fixed.err
python bench.py --workload c3 --no-secondary --no-cpu-baseline > $OUT/r02f_bench_c3_bang.json 2> $OUT/r02f_bench_c3_bang.err
python bench.py --workload c3 --control smooth --no-secondary --no-cpu-baseline > $OUT/r02f_bench_c3_smooth.json 2> $OUT/r02f_bench_c3_smooth.err
python bench.py --workload c2cloud --no-secondary --no-cpu-baseline > $OUT/r02f_bench_c2cloud.json 2> $OUT/r02f_bench_c2cloud.err
python bench.py --workload c4 --no-cpu-baseline > $OUT/r02f_bench_c4.json 2> $OUT/r02f_bench_c4.err
python profiles/measure_sens.py > $OUT/r02f_sens.json 2> $OUT/r02f_sens.err
# launch list of the default command (plain run above exited 0)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/r02f_launches_bench_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02f_ncu_launches.log 2>&1
# lanes active on config 3 (de-fit evidence): the 3-stage epoch kernel in a mid-solve epoch
python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02f_c3_plain.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:epoch_kernel<jb::HamVdpPmp<.*>, \(int\)3," \
    --launch-skip 12 --launch-count 1 -f -o $OUT/r02f_c3_epoch_s3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02f_ncu_c3.log 2>&1
ncu -i $OUT/r02f_c3_epoch_s3.ncu-rep --page details > $OUT/r02f_ncu_full_c3_epoch_kernel_s3.txt 2>&1
rm -f $OUT/r02f_c3_epoch_s3.ncu-rep
# full captures of the three epoch kernels of the headline workload at their heaviest epochs (2^18 rows keep the replay short)
for spec in "3:10" "5:22" "7:12"; do
  S=${spec%%:*}; SKIP=${spec##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:epoch_kernel<jb::HamVdp, \(int\)${S}," --launch-skip $SKIP --launch-count 1 -f -o $OUT/r02f_epoch_kernel_s${S} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --M 262144 > $OUT/r02f_ncu_full_s${S}.log 2>&1
  ncu -i $OUT/r02f_epoch_kernel_s${S}.ncu-rep --page details > $OUT/r02f_ncu_full_epoch_kernel_s${S}.txt 2>&1
  ncu -i $OUT/r02f_epoch_kernel_s${S}.ncu-rep --page raw --csv > $OUT/r02f_ncu_raw_epoch_kernel_s${S}.csv 2>&1
  rm -f $OUT/r02f_epoch_kernel_s${S}.ncu-rep
done
# the two tile kernels of config 4 (final build)
for K in mlp_grad_kernel mlp_hess_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" --launch-skip 6 --launch-count 1 \
      -f -o $OUT/r02f_$K python bench.py --workload c4 --M 262144 --steps 1 --warmup 2 --no-cpu-baseline > $OUT/r02f_ncu_full_$K.log 2>&1
  ncu -i $OUT/r02f_$K.ncu-rep --page details > $OUT/r02f_ncu_full_$K.txt 2>&1
  ncu -i $OUT/r02f_$K.ncu-rep --page raw --csv > $OUT/r02f_ncu_raw_$K.csv 2>&1
  ncu -i $OUT/r02f_$K.ncu-rep --page source --csv > $OUT/r02f_ncu_source_$K.csv 2>&1
  rm -f $OUT/r02f_$K.ncu-rep
done
python profiles/kernel_timeline.py c4 262144 > $OUT/r02f_timeline_c4.txt 2>&1
python profiles/kernel_timeline.py c2 1048576 > $OUT/r02f_timeline_c2.txt 2>&1
python profiles/measure_dopri.py > $OUT/r02f_dopri.json 2>/dev/null
tail -3 $OUT/r02f_pytest.log
