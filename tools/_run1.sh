GLSNS_TRSV_TUNE=0 timeout 300 python tools/trsv_trace.py 32 > gpurun_out/r2_trace32e.json 2>&1
GLSNS_TRSV_TUNE=0 timeout 600 python tools/trsv_trace.py 64 > gpurun_out/r2_trace64e.json 2>&1
