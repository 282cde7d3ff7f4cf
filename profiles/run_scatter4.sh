# NCCL scatter + demod path on 4 GPUs of one box (bench.py --scatter) and the plain channel-sharded bench on the same 4 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --scatter --steps 5 2>gpurun_out/scatter4.err | grep '^{' | tail -1 > gpurun_out/r02_scatter_n4.json
tail -c 700 gpurun_out/r02_scatter_n4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 5 --warmup 3 --no-sweep --no-secondary --no-cpu-baseline 2>>gpurun_out/scatter4.err | grep '^{' | tail -1 > gpurun_out/r02_bench_n4.json
tail -c 300 gpurun_out/r02_bench_n4.json; tail -3 gpurun_out/scatter4.err
