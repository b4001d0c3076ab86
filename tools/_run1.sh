timeout 900 python -m pytest tests -m gpu -q -x -k "hanging or assembly or couette or restart" > gpurun_out/r2u_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2u_gputests.log
