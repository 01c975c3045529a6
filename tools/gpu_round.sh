mkdir -p gpurun_out
export BFM_QUIET=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "independent_sparse or sim_run_matches or large_plate or deterministic or staged_job" 2>&1 | tail -3
timeout 300 python tools/compare_refinement.py 10000x2500 2>> gpurun_out/err.log | tee gpurun_out/r2_refinement_50m.jsonl
timeout 300 python tools/compare_refinement.py 2000x500 2>> gpurun_out/err.log | tee gpurun_out/r2_refinement_2m.jsonl
tail -3 gpurun_out/err.log
