V=$PWD/3drvo-marl-collisionavoidance_b200/lib/variants
run() { python bench.py --workload $2 --steps 40 --warmup 3 --no-cpu --e2e-steps 2 --threads-per-block $3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', '$2', 'tpb $3', '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"; }
for w in sat c5; do
  run base $w 128; run base $w 96
  UAISA_B200_LIB=$V/r136.so run r136 $w 96
  UAISA_B200_LIB=$V/r136.so run r136 $w 32
  UAISA_B200_LIB=$V/r136.so run r136 $w 64
  UAISA_B200_LIB=$V/r152.so run r152 $w 32
done
