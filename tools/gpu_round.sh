mkdir -p gpurun_out
export BFM_QUIET=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "galerkin_product or independent_sparse or sim_run_matches" 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>> gpurun_out/err.log | python tools/show_bench.py /dev/stdin
timeout 300 python bench.py --cells 2000x500 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check 2>> gpurun_out/err.log | python tools/show_bench.py /dev/stdin
tail -3 gpurun_out/err.log
