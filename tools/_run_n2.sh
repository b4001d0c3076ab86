#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 2 --warmup 3 > $O/n2_bench.json 2> $O/n2_bench.err; echo rc=$? >> $O/n2_bench.err
tail -c 600 $O/n2_bench.json; tail -3 $O/n2_bench.err
