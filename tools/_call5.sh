set -x
cd $GRAFT_REPO_ROOT
(timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/c5_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
(timeout 600 python bench.py --impl reference --steps 2) > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err
# launch list of the timed steps only (cold-cache, serialised: shares, not absolutes)
(TK_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline --no-config2) > gpurun_out/c5_ncu_list.log 2>&1
# full capture of the tcgen05 GEMM on three YOLOX-m layer shapes
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv1x1_tc -s 30 -c 6 -o gpurun_out/r02_conv1x1_tc python tools/bench_conv1x1.py --net yolox_m --batch 20 --reps 2) > gpurun_out/c5_ncu_full.log 2>&1
tail -8 gpurun_out/c5_tests.log; cut -c1-600 gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err; cut -c1-400 gpurun_out/c5_bench_ref.json; tail -3 gpurun_out/c5_ncu_list.log; tail -3 gpurun_out/c5_ncu_full.log; ls -la gpurun_out | tail -8
