mkdir -p gpurun_out
echo "== symbolic tests"
timeout 900 python -m pytest tests/test_gpu_symbolic.py -x -q 2>&1 | tail -15
echo "== bench 50M verbose"
BFM_JOB_VERBOSE=1 BFM_MG_VERBOSE=1 timeout 600 python bench.py --cells 10000x2500 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_sym_verbose.log | tee gpurun_out/r2_sym_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['value','ms_per_step','assembly_ms','solve_ms','cg_iterations','solve_setup_ms']}, 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and {k:round(v,1) for k,v in d['e2e']['stages_ms'].items()}, 'symbolic', round(d['symbolic_setup_ms_once_per_mesh']))
"
grep "\[job\]\|\[hier\]\|\[mg\]\|\[sim_run\]" gpurun_out/r2_sym_verbose.log | head -60
