# NCCL scatter + demod path at 1, 2, 4, 8 GPUs of one box (bench.py --scatter); one JSON line per N
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --scatter --steps 5 2>/dev/null | tail -1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --scatter --steps 5 2>/dev/null | grep '^{' | tail -1; fi
done
