set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8) > gpurun_out/c16_smoke.log 2>&1
(timeout 1200 python bench.py --steps 5 --warmup 3) > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/c16_bench_ref.json 2> gpurun_out/c16_bench_ref.err
cat gpurun_out/c16_smoke.log; cut -c1-400 gpurun_out/c16_bench.json; tail -3 gpurun_out/c16_bench.err; cut -c1-600 gpurun_out/c16_bench_ref.json
