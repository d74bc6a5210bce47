python -m pytest tests/test_gpu_ppo.py tests/test_gpu_dp.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -3
for w in c1 c5; do python bench.py --workload $w --only --no-e2e --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); w=d.get('workloads',{}).get('$w',d); print('$w', w.get('value'), w.get('ms_per_step'), (w.get('fp32_parity_path') or {}).get('value'))"; done
