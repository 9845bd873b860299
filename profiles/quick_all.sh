# quick agent-steps/s of every workload (no CPU baseline, no e2e): bash profiles/quick_all.sh [extra bench flags]
for w in c2 sat c3 c4 c5; do python bench.py --workload $w --no-cpu --no-sat --e2e-steps 2 "$@" 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', '%.4g' % d['value'], 'ms/step %.5f' % d['ms_per_step'])"; done
