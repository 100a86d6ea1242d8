cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_botsort_gpu.py -q 2>&1 | tail -20) > gpurun_out/c22.log 2>&1
cat gpurun_out/c22.log | cut -c1-400
python tools/run_botsort_only.py 500 512 2>&1 | tail -2
python tools/run_deepocsort_only.py 500 512 2>&1 | tail -1
