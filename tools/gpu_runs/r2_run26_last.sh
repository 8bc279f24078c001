#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2y_last.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -4 | tee -a gpurun_out/r2y_last.log
