cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests/test_reid_crop_gpu.py tests/test_reid_fused_gpu.py tests/test_connected_pipeline_gpu.py tests/test_strongsort_gpu.py tests/test_deepocsort_gpu.py tests/test_real_engine_gpu.py tests/test_rtdetr_gpu.py -q -x 2>&1 | tail -12) > gpurun_out/c20_tests.log 2>&1
cat gpurun_out/c20_tests.log
(timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config2 --no-extra | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['us_per_unit'],1) for k,v in d['stages'].items()}, d['kernels'].get('crop_resize_norm'))") 2>&1 | tail -1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
(timeout 600 ncu --profile-from-start off --metrics $M --clock-control none -c 2000 -k regex:'crop_resize|maxpool|avgpool' --csv --log-file gpurun_out/r02_pool_crop_after.csv python tools/ncu_forward.py) > gpurun_out/c20_ncu.log 2>&1
grep -E "crop_resize|maxpool|avgpool" gpurun_out/r02_pool_crop_after.csv | cut -d, -f5,13,15 | cut -c1-200 | head -12
