mkdir -p gpurun_out
echo "== deformation.py on the GPU"
timeout 300 python tools/examples_harness/run_reference_example.py --workdir /tmp/rundir examples/deformation.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
run() { # label, env..., cells
  label=$1; shift; cells=$1; shift
  echo "== $label $cells"
  env "$@" timeout 600 python bench.py --cells $cells --steps 2 --warmup 2 --no-cpu-baseline --no-parity-check 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','cg_restarts','cg_true_rel_residual','solve_setup_ms']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and {k:round(v,1) for k,v in d['e2e']['stages_ms'].items()}, 'first', d['e2e'] and round(d['e2e']['first_call_ms']))
"
}
run fp32smooth 10000x2500 BFM_JOB_VERBOSE=1
grep "\[job\]" gpurun_out/err.log | tail -9
run fp32smooth 2000x500 A=1
tail -3 gpurun_out/err.log
