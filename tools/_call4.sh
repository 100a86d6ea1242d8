set -x
cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_conv1x1_tc_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/c4_conv.log 2>&1
for cfg in "8 0" "12 0" "4 0" "8 1"; do set -- $cfg; (TK_C1_EPI_WARPS=$1 TK_C1_ONE_CTA=$2 timeout 300 python tools/bench_conv1x1.py --net yolox_s --batch 50 2>&1 | tail -1) >> gpurun_out/c4_micro.log 2>&1; done
(timeout 300 python tools/bench_conv1x1.py --net yolox_s --batch 50 2>&1 | tail -30) > gpurun_out/c4_micro_s.log 2>&1
(timeout 300 python tools/bench_conv1x1.py --net yolox_m --batch 20 2>&1 | tail -30) > gpurun_out/c4_micro_m.log 2>&1
(timeout 300 python tools/bench_conv1x1.py --net resnet50 --batch 768 2>&1 | tail -30) > gpurun_out/c4_micro_r.log 2>&1
tail -5 gpurun_out/c4_conv.log; cat gpurun_out/c4_micro.log; tail -12 gpurun_out/c4_micro_s.log; tail -12 gpurun_out/c4_micro_m.log; tail -14 gpurun_out/c4_micro_r.log
