#!/bin/bash
# round 2, call p (1 GPU): the whole GPU suite on the final code of the round, then smoke()
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 262 python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_$TAG.log
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
