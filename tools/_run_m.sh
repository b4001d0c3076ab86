#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/m_gputests.log 2>&1; echo rc=$? >> $O/m_gputests.log
tail -3 $O/m_gputests.log
timeout 200 python bench.py --cells 32 --steps 2 --warmup 3 --no-cpu-baseline > $O/m_bench32_fused.json 2> $O/m_bench32_fused.err
GLSNS_GMRES_FUSED=0 timeout 200 python bench.py --cells 32 --steps 2 --warmup 3 --no-cpu-baseline > $O/m_bench32_unfused.json 2> $O/m_bench32_unfused.err
python - <<'P'
import json
for f in ("fused","unfused"):
    d=json.loads(open("gpurun_out/m_bench32_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["config"]["gmres_iterations"], repr(d["config"]["true_residual"]), d["phases_ms_per_step"]["orthog_ms"], d["ms_per_step"])
P
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
