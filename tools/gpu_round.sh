mkdir -p gpurun_out
export BFM_QUIET=1
BFM_JOB_VERBOSE=1 python tools/profile_symbolic.py 10000x2500 2>&1 | grep -v "^\[job\]"
echo "== cold start + bench"
BFM_JOB_VERBOSE=1 BFM_MG_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check 2> gpurun_out/r2_cold_verbose.log | python tools/show_bench.py /dev/stdin
grep "\[plan\]\|\[job\]\|\[hier\] level 0" gpurun_out/r2_cold_verbose.log | sed -n 1,20p
