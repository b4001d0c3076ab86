timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2l_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2l_gputests.log
timeout 600 python tools/profile_kernels.py 64 spmv ilu_apply > gpurun_out/r2l_kernels64.json 2>&1
