mkdir -p gpurun_out
export BFM_QUIET=1
echo "== full gpu suite (defaults: smoothed aggregation)"
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default"
( time timeout 900 python bench.py 2>> gpurun_out/err.log > gpurun_out/r2_bench_n1_sa.json ) 2>&1 | tail -3
python tools/show_bench.py gpurun_out/r2_bench_n1_sa.json
echo "== bench 2000x500 / 500x125 / plain"
for c in 2000x500 500x125; do timeout 300 python bench.py --cells $c --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check 2>> gpurun_out/err.log | python tools/show_bench.py /dev/stdin; done
BFM_MG_SMOOTH=0 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity-check --no-e2e 2>> gpurun_out/err.log | python tools/show_bench.py /dev/stdin
tail -3 gpurun_out/err.log
