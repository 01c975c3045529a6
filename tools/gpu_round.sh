mkdir -p gpurun_out
run() { # label, env..., cells
  label=$1; shift; cells=$1; shift
  echo "== $label $cells"
  env "$@" timeout 600 python bench.py --cells $cells --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','coarse_dim','solve_setup_ms','gpu_launches']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1))
"
}
run default 10000x2500 A=1
run gamma1 10000x2500 BFM_MG_GAMMA=1
run gamma221 10000x2500 BFM_MG_GAMMA=2,2,1
run gamma21 10000x2500 BFM_MG_GAMMA=2,1
run r16_4 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4
run r16_4_g2221 10000x2500 BFM_MG_RATIO0=16 BFM_MG_RATIO=4 BFM_MG_GAMMA=2,2,2,1
run r9_6 10000x2500 BFM_MG_RATIO0=9 BFM_MG_RATIO=6
run dense1024 10000x2500 BFM_MG_DENSE_NODES=1024
run om1.8 10000x2500 BFM_MG_OMEGA=1.8
run om1.33 10000x2500 BFM_MG_OMEGA=1.33
run default 2000x500 A=1
run gamma21 2000x500 BFM_MG_GAMMA=2,1
run chunk4 2000x500 BFM_CG_CHUNK=4
echo "== ncu launch list 50M, 6 iterations"
BFM_QUIET=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_mg_50m.csv python tools/profile_target.py 10000x2500 6 2>&1 | tail -3
tail -5 gpurun_out/err.log
