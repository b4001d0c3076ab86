GLSNS_TRSV_TUNE=0 timeout 300 python tools/trsv_trace.py 32 > gpurun_out/r2_trace32e.json 2>&1
GLSNS_TRSV_TUNE=0 timeout 600 python tools/trsv_trace.py 64 > gpurun_out/r2_trace64e.json 2>&1
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_TEAMS=3 GLSNS_TRSV_HELPERS=4 timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2_sweep64_3x4.json 2>&1
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_TEAMS=3 GLSNS_TRSV_HELPERS=4 timeout 600 python tools/trsv_sweep.py 32 > gpurun_out/r2_sweep32_3x4.json 2>&1
GLSNS_TRSV_TUNE=0 GLSNS_TRSV_TEAMS=2 GLSNS_TRSV_HELPERS=6 timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2_sweep64_2x6.json 2>&1
GLSNS_TRSV_TUNE=3 GLSNS_TRSV_DEBUG=1 timeout 600 python tools/trsv_sweep.py 64 > gpurun_out/r2_sweep64_tune.json 2>&1
