#!/bin/bash
# usage (under gpurun): bash profiles/refresh.sh  -> gpurun_out/r1/*.json : every bench line of the round, re-measured in one sitting
O=gpurun_out/r1; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/gpu.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
for w in c3 c4 c5 sat; do python bench.py --workload $w --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; done
python bench.py --workload sat --no-cpu --no-autoreset > $O/bench_sat_plainstep.json 2>/dev/null
python bench.py --workload ppo > $O/bench_ppo_f32.json 2> $O/bench_ppo_f32.err
python bench.py --workload ppo --tf32 > $O/bench_ppo_tf32.json 2> $O/bench_ppo_tf32.err
bash profiles/host_step_attrib.sh > $O/host_step.txt 2>&1
for f in $O/*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('  value %.4g %s | e2e %s | frac %s' % (d.get('value',0), d.get('unit'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac')))"; done
