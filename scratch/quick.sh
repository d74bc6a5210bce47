python -m pytest tests/test_gpu_env.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['value'], d['roofline']['frac'], d['roofline']['achieved'])"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:rollout_kernel -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "inst_executed|duration|issue_active"
