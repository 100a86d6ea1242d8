cd $GRAFT_REPO_ROOT
(timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) | cut -c1-200
(timeout 300 python __graft_entry__.py smoke 2>&1 | grep "smoke:")
