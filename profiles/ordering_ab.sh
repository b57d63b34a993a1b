for w in "" "--workload c2cloud" "--workload c3 --M 262144" "--workload c3 --control smooth --M 262144"; do
  for h in 4 2 3; do
    JB200_EPOCH_HORDER=$h python bench.py $w --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/ab_o.json 2> gpurun_out/ab_o.err
    python -c "
import json;d=json.loads(open('gpurun_out/ab_o.json').read().strip().splitlines()[-1]);print('horder $h [$w]', '%.4e'%d['value'], '%.2f ms'%d['ms_per_step'], '%.4f'%d['roofline']['frac'])"
  done
done
