export BFM_QUIET=1
timeout 48 python -m pytest tests/test_gpu_parity.py -x -q -k "independent_sparse" 2>&1 | tail -3
