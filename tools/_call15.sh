set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_deepocsort_gpu.py tests/test_reid_crop_gpu.py -q 2>&1 | tail -30) > gpurun_out/c15.log 2>&1
cat gpurun_out/c15.log
