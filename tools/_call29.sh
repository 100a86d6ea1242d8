cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_botsort_gpu.py tests/test_strongsort_gpu.py -q 2>&1 | tail -3) | cut -c1-200
for ex in 0 1; do for c in 16 8; do (TK_SS_EXCLUSIVE_SM=$ex timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-config2 --no-extra --ctas $c | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exclusive', $ex, 'ctas', $c, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['us_per_unit'],1) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])") 2>&1 | tail -1; done; done
