mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # label, env..., cells
  label=$1; shift; cells=$1; shift
  echo "== $label $cells"
  env "$@" timeout 600 python bench.py --cells $cells --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','cg_restarts','solve_setup_ms']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and {k:round(v,1) for k,v in d['e2e']['stages_ms'].items()})
"
}
run pdl 10000x2500 A=1
run nopdl 10000x2500 BFM_PDL=0
run pdl 2000x500 A=1
run nopdl 2000x500 BFM_PDL=0
run pdl 500x125 A=1
run nopdl 500x125 BFM_PDL=0
echo "== irregular numbering probe 6000x1500"
timeout 900 python tools/irregular_probe.py 6000x1500 2>> gpurun_out/err.log | tee gpurun_out/r2_irregular_probe_18m.jsonl
tail -3 gpurun_out/err.log
