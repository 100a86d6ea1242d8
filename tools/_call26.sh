cd $GRAFT_REPO_ROOT
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/c26_tests.log 2>&1
cat gpurun_out/c26_tests.log | cut -c1-300
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:'deepocsort_video|botsort_video' -c 2 -o gpurun_out/r02_trackers_f python tools/run_botsort_only.py 60 512) > gpurun_out/c26_ncu1.log 2>&1
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:'deepocsort_video' -c 1 -o gpurun_out/r02_deepocsort python tools/run_deepocsort_only.py 60 512) > gpurun_out/c26_ncu2.log 2>&1
(timeout 400 ncu --set full --clock-control none -k regex:'crop_resize_norm|maxpool3x3s2_rows|avgpool_kernel|hota_' -c 12 -o gpurun_out/r02_small_kernels python __graft_entry__.py smoke) > gpurun_out/c26_ncu3.log 2>&1
tail -2 gpurun_out/c26_ncu1.log gpurun_out/c26_ncu2.log gpurun_out/c26_ncu3.log | cut -c1-200
(timeout 1200 python bench.py --steps 5 --warmup 3) > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err
tail -2 gpurun_out/c26_bench.err; cut -c1-200 gpurun_out/c26_bench.json
