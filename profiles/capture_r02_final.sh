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
tail -3 $OUT/r02f_pytest.log
