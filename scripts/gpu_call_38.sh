# last sanity after the comment-only rebuild: smoke + model tests
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_models.py -q --timeout 300 -x 2>&1 | tail -2
