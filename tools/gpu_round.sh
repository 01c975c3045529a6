mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== full bench line"; timeout 900 python bench.py --steps 3 --warmup 3 2>> gpurun_out/err.log | tee gpurun_out/bench_full.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ['value','ms_per_step','assembly_ms','solve_ms','solve_setup_ms','cg_iterations','cg_restarts','mg_levels','cg_rel_residual','cg_true_rel_residual','cg_backward_error','symbolic_setup_ms_once_per_mesh','gpu_launches','same_size','parity_check','clocks']: print(k, d[k])
print('e2e', d['e2e']); print('roofline', {k:v for k,v in d['roofline'].items() if k!='cg_iteration'}); print('iter', {k:v for k,v in d['roofline']['cg_iteration'].items() if k!='model'}); print('asm', d['roofline_assembly'])
"
echo "== batch"; timeout 300 python bench.py --workload batch --steps 3 --warmup 3 2>> gpurun_out/err.log | tee gpurun_out/bench_batch.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','assembly_ms','solve_ms']}, d['e2e'])
"
tail -5 gpurun_out/err.log
