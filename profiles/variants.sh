#!/bin/bash
# usage (under gpurun): bash profiles/variants.sh  -> prints agent-steps/s of every lib in lib/variants for c5/sat
for so in 3drvo-marl-collisionavoidance_b200/lib/variants/*.so; do
  for w in ${WORKLOADS:-c5 sat}; do
    UAISA_B200_LIB=$PWD/$so python bench.py --workload $w --steps 40 --warmup 3 --no-cpu --e2e-steps 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$(basename $so)', '$w', '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'single %.3e'%d['single_step_launch']['value'])"
  done
done
