# configs[3] (64 envs x 256 drones): egos per CTA x neighbour slices sweep  (GEOMS="64:3 32:3" to pick)
mkdir -p gpurun_out
for g in ${GEOMS:-128:1 64:1 32:1 64:2 32:2 32:3 32:4 64:3 0:0}; do
t=${g%%:*}; s=${g##*:}
UAISA_VERBOSE=1 python bench.py --workload c4 --no-cpu --no-subs --e2e-steps 2 --threads-per-block $t --slices $s 2>gpurun_out/geo.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tpb $t slices $s', '%.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'])"
grep -m1 uaisa gpurun_out/geo.err
done
