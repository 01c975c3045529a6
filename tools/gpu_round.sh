# What a round ends with on one B200 (under gpurun): the GPU suite, the smoke test, the default bench line.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh'
# Several GPUs:  gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_dist.py -q'  and
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N
mkdir -p gpurun_out
export BFM_QUIET=1
echo "== gpu suite"
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default"
( time timeout 900 python bench.py 2>> gpurun_out/err.log > gpurun_out/bench_n1.json ) 2>&1 | tail -3
python tools/show_bench.py gpurun_out/bench_n1.json
tail -3 gpurun_out/err.log
