mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== dist test world 2"
timeout 420 python -m pytest tests/test_gpu_dist.py -x -q -k "2-peer or 2-nccl" 2>&1 | tail -30
run() { # label, n, cells, env...
  label=$1; shift; n=$1; shift; cells=$1; shift
  echo "== $label N=$n $cells"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --cells $cells --steps 2 --warmup 2 --no-cpu-baseline 2>> gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(d[k],2) if isinstance(d[k],float) and d[k] > 1e-3 else d[k]) for k in ['ms_per_step','solve_ms','cg_iterations','cg_restarts','mg_levels','cg_rel_residual','cg_true_rel_residual','cg_backward_error','solve_setup_ms','gpu_launches','exchange']}, 'us/iter', round(d['roofline']['cg_iteration']['us'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1))
"
}
run mg 2 2000x500
run mg 2 10000x2500
run legacy 2 10000x2500 BFM_MG=0
tail -20 gpurun_out/err.log
