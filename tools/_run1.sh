timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2t_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2t_gputests.log
GLSNS_FULL_SIZE_CELLS=64 timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -k full_size_properties > gpurun_out/r2t_full_size_n64.log 2>&1; echo rc=$? >> gpurun_out/r2t_full_size_n64.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2t_bench64.json 2> gpurun_out/r2t_bench64.err; echo rc=$? >> gpurun_out/r2t_bench64.err
