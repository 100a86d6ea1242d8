set -x
cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_conv1x1_tc_gpu.py -x -q 2>&1 | tail -25) > gpurun_out/c3_conv.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_conv1x1_tc_gpu.py 2>&1 | tail -40) > gpurun_out/c3_tests.log 2>&1
if grep -q passed gpurun_out/c3_conv.log && ! grep -q failed gpurun_out/c3_conv.log; then TC=0; else TC=1; fi
(TK_NO_TC=$TC timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
(TK_NO_TC=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/c3_bench_notc.json 2> gpurun_out/c3_bench_notc.err
tail -25 gpurun_out/c3_conv.log; tail -30 gpurun_out/c3_tests.log; cut -c1-1500 gpurun_out/c3_bench.json; tail -5 gpurun_out/c3_bench.err
