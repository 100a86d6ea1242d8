cd $GRAFT_REPO_ROOT
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
(timeout 400 ncu --metrics $M --clock-control none -k regex:'crop_resize_norm' -c 4 --csv --log-file gpurun_out/r02_crop_after.csv python tools/ncu_targets.py) > gpurun_out/c32.log 2>&1
tail -2 gpurun_out/c32.log | cut -c1-200; grep crop_resize gpurun_out/r02_crop_after.csv | cut -d, -f5,12- | cut -c1-300 | head -12
