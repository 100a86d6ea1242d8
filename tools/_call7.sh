set -x
cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_conv1x1_tc_gpu.py -x -q 2>&1 | tail -25) > gpurun_out/c7_conv.log 2>&1
(timeout 600 python -m pytest tests/test_real_engine_gpu.py tests/test_nn_epilogue_gpu.py tests/test_detector_kernels_gpu.py tests/test_connected_pipeline_gpu.py -q 2>&1 | tail -25) > gpurun_out/c7_tests.log 2>&1
(timeout 300 python tools/bench_conv1x1.py --net yolox_s --batch 50 2>&1 | tail -40) > gpurun_out/c7_micro_s.log 2>&1
(timeout 300 python tools/bench_conv1x1.py --net yolox_m --batch 20 2>&1 | tail -40) > gpurun_out/c7_micro_m.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline) > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
(TK_NO_TC3=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/c7_bench_notc3.json 2> gpurun_out/c7_bench_notc3.err
tail -25 gpurun_out/c7_conv.log; tail -12 gpurun_out/c7_tests.log; tail -3 gpurun_out/c7_micro_s.log; tail -3 gpurun_out/c7_micro_m.log; cut -c1-300 gpurun_out/c7_bench.json; tail -3 gpurun_out/c7_bench.err
