timeout 900 python -m pytest tests -m gpu -q -x -k "spmv or gmres or restart or full_size" > gpurun_out/r2j_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2j_gputests.log
timeout 600 python tools/profile_kernels.py 64 spmv orthog assemble_system ilu_factor > gpurun_out/r2j_kernels64.json 2>&1
timeout 600 python tools/profile_kernels.py 32 spmv orthog > gpurun_out/r2j_kernels32.json 2>&1
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2j_bench64.json 2> gpurun_out/r2j_bench64.err
