#!/bin/bash
# blocks of 8: full parity suite
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_b8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_b8.log 2>&1
