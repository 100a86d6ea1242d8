cd $GRAFT_REPO_ROOT
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
(timeout 500 ncu --profile-from-start off --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/r02_forward_metrics_b50.csv python tools/ncu_forward.py --batch 50 --crops 1800) > gpurun_out/c31.log 2>&1
tail -1 gpurun_out/c31.log; wc -l gpurun_out/r02_forward_metrics_b50.csv
