set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/weights
(timeout 900 python -m pytest tests/test_connected_pipeline_gpu.py tests/test_pairwise_gpu.py tests/test_strongsort_gpu.py tests/test_bpbreid_gpu.py tests/test_trackers_edge_gpu.py tests/test_detector_kernels_gpu.py -x -q -s 2>&1 | tail -40) > gpurun_out/c2_tests.log 2>&1
(timeout 600 python tools/strict_parity_probe.py) > gpurun_out/c2_strict.log 2>&1
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/c2_smoke.log 2>&1
(timeout 1200 python tools/train_synth_detector.py --variant m --steps 1200 --out gpurun_out/weights/yolox_m_synth.pt) > gpurun_out/c2_train_m.log 2>&1
tail -25 gpurun_out/c2_tests.log; cat gpurun_out/c2_strict.log | cut -c1-220; tail -5 gpurun_out/c2_smoke.log; tail -3 gpurun_out/c2_train_m.log
