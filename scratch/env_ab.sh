# A/B of library builds on the env-only workloads: bash scratch/env_ab.sh libA.so libB.so
for l in "$@"; do for w in c4 c2; do DRONECU_LIB=$PWD/drone_rl_b200/$l python bench.py --workload $w --only --no-e2e --no-cpu --steps 20 --warmup 3 --sustain-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); w=d.get('workloads',{}).get('$w',d); print('$l $w', w.get('value'), w.get('ms_per_step'))"; done; done
