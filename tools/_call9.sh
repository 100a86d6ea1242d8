set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_hota_gpu.py -q 2>&1 | tail -15) > gpurun_out/c9_hota.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
(timeout 900 ncu --profile-from-start off --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/r02_forward_metrics.csv python tools/ncu_forward.py) > gpurun_out/c9_fwd.log 2>&1
(TK_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_eager.csv python bench.py --steps 1 --warmup 2 --frames 100 --no-e2e --no-cpu-baseline --no-config2 --no-graphs) > gpurun_out/c9_list_eager.log 2>&1
(TK_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_graph.csv python bench.py --steps 1 --warmup 3 --frames 100 --no-e2e --no-cpu-baseline --no-config2) > gpurun_out/c9_list_graph.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
tail -6 gpurun_out/c9_hota.log; tail -2 gpurun_out/c9_fwd.log; wc -l gpurun_out/r02_forward_metrics.csv gpurun_out/r02_launches_eager.csv gpurun_out/r02_launches_graph.csv; tail -2 gpurun_out/c9_list_eager.log gpurun_out/c9_list_graph.log; cut -c1-300 gpurun_out/c9_bench.json; tail -3 gpurun_out/c9_bench.err
