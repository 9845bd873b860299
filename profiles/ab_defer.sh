#!/bin/bash
# A/B of the deferred re-observation (UAISA_NO_DEFER=1 switches it off): agent-steps/s and roofline fraction per workload
for w in ${WORKLOADS:-c2 c5 sat c3 c4}; do
  for nd in 0 1; do
    if [ $nd = 1 ]; then export UAISA_NO_DEFER=1; else unset UAISA_NO_DEFER; fi
    python bench.py --workload $w --no-cpu --no-subs --e2e-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', 'no_defer=$nd', '%.4g'%d['value'], 'frac %.3f' % d['roofline']['frac'], 'spl', d['config']['steps_per_launch'])"
  done
done
