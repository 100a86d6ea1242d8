cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests/test_pairwise_gpu.py tests/test_strongsort_gpu.py tests/test_bpbreid_gpu.py tests/test_deepocsort_gpu.py tests/test_hota_gpu.py tests/test_trackers_edge_gpu.py tests/test_connected_pipeline_gpu.py tests/test_ecc_gpu.py -q -x 2>&1 | tail -12) > gpurun_out/c19_tests.log 2>&1
cat gpurun_out/c19_tests.log
python tools/bench_lsap.py 2>&1 | tail -4
python tools/run_deepocsort_only.py 500 512 2>&1 | tail -1
(timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-config2 --no-extra | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['us_per_unit'],1) for k,v in d['stages'].items()})") 2>&1 | tail -1
(timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-config2 --no-extra --ctas 8 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas8 fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['us_per_unit'],1) for k,v in d['stages'].items()})") 2>&1 | tail -1
