cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_botsort_gpu.py -q 2>&1 | tail -4)
python tools/run_botsort_only.py 500 512 2>&1 | tail -1
