export OMP_NUM_THREADS=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; echo rc=$? >> gpurun_out/r2n_bench_n8.err
nvidia-smi --query-gpu=index,memory.used --format=csv > gpurun_out/r2n_mem_before.csv
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --cells 116 --steps 1 --warmup 1 --no-parity-check > gpurun_out/r2n_bench_n8_c116.json 2> gpurun_out/r2n_bench_n8_c116.err; echo rc=$? >> gpurun_out/r2n_bench_n8_c116.err
