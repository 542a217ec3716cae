#!/bin/bash
# N-GPU bench line exactly as the driver launches it: scripts/bench_n.sh N [extra bench.py flags]
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"
