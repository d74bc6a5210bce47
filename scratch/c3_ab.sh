for l in "$@"; do DRONECU_LIB=$PWD/drone_rl_b200/$l python bench.py --workload c3 --only --no-e2e --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); w=d.get('workloads',{}).get('c3',d); print('$l', w['tensor_core_path']['value'], w['tensor_core_path']['ms_per_step'])"; done
