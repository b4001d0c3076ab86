timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2h_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2h_gputests.log
for L in 0 1; do GLSNS_TRSV_TUNE=0 GLSNS_TRSV_LAYOUT=$L timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2h_sweep64_L$L.json 2>&1; done
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_LAYOUT=2 GLSNS_TRSV_HELPERS=6 timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2h_sweep64_L2.json 2>&1
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_LAYOUT=1 GLSNS_TRSV_HELPERS=6 timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2h_sweep64_L1h6.json 2>&1
for L in 0 1; do GLSNS_TRSV_TUNE=0 GLSNS_TRSV_LAYOUT=$L timeout 600 python tools/trsv_sweep.py 32 > gpurun_out/r2h_sweep32_L$L.json 2>&1; done
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_LAYOUT=1 timeout 600 python tools/trsv_trace.py 64 > gpurun_out/r2h_trace64_L1.json 2>&1
