mkdir -p gpurun_out
echo "== symbolic + renumbering tests"
timeout 900 python -m pytest tests/test_gpu_symbolic.py -x -q 2>&1 | tail -15
echo "== irregular probe 6000x1500"
timeout 600 python tools/irregular_probe.py 6000x1500 2>> gpurun_out/err.log | tee gpurun_out/r2_irregular_probe_18m_renumbered.jsonl
tail -3 gpurun_out/err.log
