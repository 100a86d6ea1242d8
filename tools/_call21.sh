cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_botsort_gpu.py -q -x 2>&1 | tail -30) > gpurun_out/c21.log 2>&1
cat gpurun_out/c21.log | cut -c1-400
