set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/c8_tests.log 2>&1
(timeout 200 python tools/probe_conv3x3_nan.py) > gpurun_out/c8_probe.log 2>&1
(timeout 600 python tools/stress_sweep.py --reps 10) > gpurun_out/r02_stress_sweep_1gpu.jsonl 2> gpurun_out/c8_stress.err
for b in 10 25 50; do (timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-config2 --batch $b | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch', d['config']['detector_batch'], 'fps', d['value'], {k:v['us_per_unit'] for k,v in d['stages'].items()})") >> gpurun_out/c8_batch.log 2>&1; done
tail -8 gpurun_out/c8_tests.log; cat gpurun_out/c8_probe.log | cut -c1-300; grep -E "cosine|lap" gpurun_out/r02_stress_sweep_1gpu.jsonl | cut -c1-300; cat gpurun_out/c8_batch.log
