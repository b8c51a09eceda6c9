# N-GPU evidence (gpurun --gpus N): the 2-rank NCCL parity test, then bench.py weak and strong scaling
N=${N:-2}; TAG=${TAG:-mg}
python -m pytest tests/test_dist_nccl.py -m gpu -x -q > gpurun_out/${TAG}_nccl_test.log 2>&1; tail -2 gpurun_out/${TAG}_nccl_test.log
for sc in weak strong; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --scaling $sc > gpurun_out/${TAG}_bench_${N}gpu_$sc.log 2> gpurun_out/${TAG}_bench_${N}gpu_$sc.err
  tail -c 700 gpurun_out/${TAG}_bench_${N}gpu_$sc.log
done
