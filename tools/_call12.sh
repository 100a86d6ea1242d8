set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_deepocsort_gpu.py -q 2>&1 | tail -8) > gpurun_out/c12_doc.log 2>&1
(timeout 300 python tools/run_deepocsort_only.py 500 512 | tail -4) > gpurun_out/c12_doc_time.log 2>&1
cat gpurun_out/c12_doc.log gpurun_out/c12_doc_time.log
