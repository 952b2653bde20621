#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -rf > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s9_smoke.log
tail -8 gpurun_out/s9_pytest.log; tail -2 gpurun_out/s9_smoke.log
