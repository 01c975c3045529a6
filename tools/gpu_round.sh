mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # n
  n=$1
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 3 --warmup 3 2>> gpurun_out/err.log | tee gpurun_out/r2_bench_n${n}_sa.json | python tools/show_bench.py /dev/stdin
}
run 8
run 4
grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$\|NCCL version" gpurun_out/err.log | tail -5
