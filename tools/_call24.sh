cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_jpeg_gpu.py -q -s 2>&1 | grep -E "sampling|passed|failed|assert|Error" | head -12) | cut -c1-300
