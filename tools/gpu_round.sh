mkdir -p gpurun_out
run() { # label, env..., cells
  label=$1; shift; cells=$1; shift
  echo "== $label $cells"
  env "$@" timeout 600 python bench.py --cells $cells --steps 2 --warmup 2 --no-cpu-baseline --no-parity-check --no-e2e 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','cg_restarts','mg_levels','coarse_dim','solve_setup_ms']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1))
"
}
run r16_4_om18 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8
run r16_5_om18 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=5 BFM_MG_OMEGA=1.8
run r20_4_om18 10000x2500 BFM_MG_RATIO0=20 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8
run r16_4_om19 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.9
run r16_4_om18_d2048 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8 BFM_MG_DENSE_NODES=2048
run r16_4_om18 2000x500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8
run r16_4_om18 500x125 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8
run r16_6_om18 2000x500 BFM_MG_RATIO0=16 BFM_MG_OMEGA=1.8
echo "== gear60 / small meshes with r16_4 om1.8 (general path)"
BFM_ONE_CTA=0 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_OMEGA=1.8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sim_run_matches_reference or independent_sparse" 2>&1 | tail -3
tail -3 gpurun_out/err.log
