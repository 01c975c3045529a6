mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not independent_sparse" 2>&1 | tail -5
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "independent_sparse" 2>&1 | tail -12
run() { # label, env..., cells
  label=$1; shift; cells=$1; shift
  echo "== $label $cells"
  env "$@" timeout 600 python bench.py --cells $cells --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','cg_restarts','cg_rel_residual','cg_true_rel_residual','cg_backward_error','solve_setup_ms','gpu_launches']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1), 'frac', round(d['roofline']['cg_iteration']['frac'],3))
"
}
run refine 10000x2500 A=1
run norefine 10000x2500 BFM_CG_REFINE=0
run refine_r16_4 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4
run refine 2000x500 A=1
run norefine 2000x500 BFM_CG_REFINE=0
tail -5 gpurun_out/err.log
