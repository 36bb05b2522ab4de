# Developer tool: the 2 / 4 / 8-GPU lines of the scaling run in one `gpurun --gpus 8` call (outputs under gpurun_out/).
tag=${1:-scale}
for n in 2 4 8; do
  extra="--no-sub"; [ $n = 8 ] && extra=""
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 100 --warmup 5 $extra > gpurun_out/${tag}_${n}gpu.json 2> gpurun_out/${tag}_${n}gpu.err
  echo "n=$n rc=$?"; tail -c 300 gpurun_out/${tag}_${n}gpu.json | head -c 0
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/nccl_gather_check.py > gpurun_out/${tag}_nccl_gather.log 2>&1; echo "gather rc=$?"
