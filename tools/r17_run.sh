#!/bin/bash
# one-off GPU session: parity and timing of sgdx_match_compact with 64-bit cuckoo entries + short seed index
cd "$(dirname "$0")/.."
timeout 35 python -m pytest tests/test_gpu_compact.py -m gpu -x -q > gpurun_out/r18_compact_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r18_compact_pytest.log
timeout 25 python tools/bench_compact.py --no-e2e > gpurun_out/r18_compact.json 2> gpurun_out/r18_compact.err
cat gpurun_out/r18_compact.json; tail -2 gpurun_out/r18_compact.err
