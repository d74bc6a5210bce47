# A/B of two library builds on the PPO workloads (tensor-core and fp32 parity paths); usage: bash scratch/ab_bench.sh libA.so libB.so
for l in "$@"; do
  for w in c3 c5 c1; do
    DRONECU_LIB=$PWD/drone_rl_b200/$l python bench.py --workload $w --only --no-e2e --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read())
w = d.get('workloads', {}).get('$w', d)
tc, fp = w.get('tensor_core_path', {}), w.get('fp32_parity_path', {})
print('$l $w tensor-core', tc.get('value'), tc.get('ms_per_step'), '| fp32', fp.get('value'), fp.get('ms_per_step'))
"
  done
done
