cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_reid_crop_gpu.py tests/test_rtdetr_gpu.py "tests/test_deepocsort_gpu.py::test_xyxy_int_crop_rule_matches_pil" tests/test_connected_pipeline_gpu.py -q 2>&1 | tail -3) | cut -c1-200
M=gpu__time_duration.sum
(timeout 200 ncu --metrics $M --clock-control none -k regex:'crop_resize_norm' -c 2 --csv --log-file gpurun_out/r02_crop_after2.csv python tools/ncu_targets.py) > gpurun_out/c33.log 2>&1
grep crop_resize gpurun_out/r02_crop_after2.csv | cut -d, -f9,15 | head -3
