#!/bin/bash
# usage (under gpurun): [WORKLOADS="c2 sat"] bash profiles/variants.sh  -> agent-steps/s of the in-tree lib and of every lib in lib/variants (two runs each)
for rep in 1 2; do
for so in 3drvo-marl-collisionavoidance_b200/lib/libuaisa_b200.so 3drvo-marl-collisionavoidance_b200/lib/variants/*.so; do
  for w in ${WORKLOADS:-c2 c5 sat}; do
    UAISA_B200_LIB=$PWD/$so python bench.py --workload $w --no-cpu --no-subs --e2e-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $so)', '$w', '%.4g'%d['value'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
done
