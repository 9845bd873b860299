#!/bin/bash
# usage (under gpurun): [WORKLOADS="c2 sat"] bash profiles/variants.sh  -> agent-steps/s of the in-tree lib and of every lib in lib/variants
for so in 3drvo-marl-collisionavoidance_b200/lib/libuaisa_b200.so 3drvo-marl-collisionavoidance_b200/lib/variants/*.so; do
  for w in ${WORKLOADS:-c5 sat}; do
    UAISA_B200_LIB=$PWD/$so python bench.py --workload $w --no-cpu --no-sat --e2e-steps 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $so)', '$w', '%.4g'%d['value'])"
  done
done
