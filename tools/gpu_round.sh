mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== dist tests"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6
run() { # label, n, cells, env...
  label=$1; shift; n=$1; shift; cells=$1; shift
  echo "== $label N=$n $cells"
  if [ "$n" = "1" ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711"; fi
  env "$@" timeout 400 $launcher bench.py --gpus $n --cells $cells --steps 3 --warmup 3 2>> gpurun_out/err.log | tee gpurun_out/r2_bench_n${n}_${label}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['value','ms_per_step','assembly_ms','solve_ms','cg_iterations','cg_restarts','mg_levels','cg_rel_residual','cg_true_rel_residual','solve_setup_ms','gpu_launches']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1), 'spmv', round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and {k:round(v,1) for k,v in d['e2e']['stages_ms'].items()}, 'parity', d['parity_check'] and max(d['parity_check']['rel_l2'].values()), 'symbolic', round(d['symbolic_setup_ms_once_per_mesh']), 'clocks', d['clocks'])
"
}
run mg 8 10000x2500
run mg 4 10000x2500
run mg 2 10000x2500
run mg 1 10000x2500
run mg_r4 8 10000x2500 BFM_MG_RATIO=4
grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$\|NCCL version" gpurun_out/err.log | tail -10
