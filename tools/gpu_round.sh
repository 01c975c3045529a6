mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== dist tests (world 2)"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6
echo "== bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 2>> gpurun_out/err.log | tee gpurun_out/r2_bench_n2_final.json | python tools/show_bench.py /dev/stdin
grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$\|NCCL version" gpurun_out/err.log | tail -5
