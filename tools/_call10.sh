set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_deepocsort_gpu.py -q -x 2>&1 | tail -30) > gpurun_out/c10_doc.log 2>&1
(timeout 300 python tools/run_deepocsort_only.py 500 512) > gpurun_out/c10_doc_time.log 2>&1
(timeout 600 python -m pytest tests/test_ocsort_gpu.py tests/test_engine_modules_gpu.py -q 2>&1 | tail -5) > gpurun_out/c10_oc.log 2>&1
cat gpurun_out/c10_doc.log; cat gpurun_out/c10_doc_time.log | tail -5; cat gpurun_out/c10_oc.log
