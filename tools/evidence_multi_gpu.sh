# Multi-GPU bench lines of a round: bash tools/evidence_multi_gpu.sh N [configs]   (under gpurun --gpus N)
N=$1
for c in ${2:-4 5}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$c bench.py --gpus $N --config $c --steps 3 --warmup 3 \
    > gpurun_out/bench_r2_config${c}_n$N.json 2> gpurun_out/bench_c${c}_n$N.err
  tail -c 300 gpurun_out/bench_r2_config${c}_n$N.json; echo
done
