set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/weights
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/c1_tests.log 2>&1
(timeout 600 python tools/strict_parity_probe.py) > gpurun_out/c1_strict.log 2>&1
(timeout 900 python tools/train_synth_detector.py --variant s --steps 700 --out gpurun_out/weights/yolox_s_synth.pt) > gpurun_out/c1_train_s.log 2>&1
tail -3 gpurun_out/c1_tests.log; cat gpurun_out/c1_strict.log | tail -30; tail -5 gpurun_out/c1_train_s.log
