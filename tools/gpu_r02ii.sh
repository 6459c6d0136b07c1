#!/bin/bash
cd "$(dirname "$0")/.."
export LATTE_B200_NO_BUILD=1
timeout 300 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -2
timeout 300 python tools/gpu_train_bench.py 5 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
