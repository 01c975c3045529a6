# what a round ends with on one B200: the GPU suite, the smoke test, the default bench line
mkdir -p gpurun_out
export BFM_QUIET=1
echo "== gpu suite"
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default"
( time timeout 900 python bench.py 2>> gpurun_out/err.log > gpurun_out/r2_bench_n1_final.json ) 2>&1 | tail -3
python tools/show_bench.py gpurun_out/r2_bench_n1_final.json
tail -3 gpurun_out/err.log
