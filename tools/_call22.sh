cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_botsort_gpu.py -q 2>&1 | tail -20) > gpurun_out/c22.log 2>&1
cat gpurun_out/c22.log | cut -c1-400
(timeout 600 python __graft_entry__.py smoke 2>&1 | grep smoke:) 
(timeout 1200 python bench.py --steps 5 --warmup 3) > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
tail -2 gpurun_out/c22_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/c22_bench.json').read().strip().splitlines()[-1])
print('fps', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['traffic']); print(d['trackers_alone']); print(d['hota_vs_generator']['device'])"
