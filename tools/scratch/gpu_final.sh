#!/bin/bash
# end-of-round validation on one GPU: full -m gpu suite, smoke(), default bench
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
