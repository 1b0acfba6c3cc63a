#!/bin/bash
# whole GPU suite + smoke, logs into gpurun_out/
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/t_all.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
