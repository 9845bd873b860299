#!/bin/bash
# usage (under gpurun): bash profiles/variants_e2e.sh  -> device-resident value and e2e (host round trip) of the in-tree lib and of every lib in lib/variants
for rep in 1 2; do
for so in 3drvo-marl-collisionavoidance_b200/lib/libuaisa_b200.so 3drvo-marl-collisionavoidance_b200/lib/variants/*.so; do
  for w in ${WORKLOADS:-c2 sat}; do
    UAISA_B200_LIB=$PWD/$so python bench.py --workload $w --no-cpu --no-subs --e2e-steps ${E2E:-1000} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $so)', '$w', '%.4g'%d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'])"
  done
done
done
