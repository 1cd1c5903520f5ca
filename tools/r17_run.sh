#!/bin/bash
# one-off GPU session for the short seed index of sgdx_match_compact: parity, timing, then ncu of the faster variant
cd "$(dirname "$0")/.."
SGDX_SHORT_SEED=1 timeout 40 python -m pytest tests/test_gpu_compact.py -m gpu -x -q -k "oracle or baseline or batcher" > gpurun_out/r17_short_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r17_short_pytest.log
SGDX_SHORT_SEED=1 timeout 30 python tools/bench_compact.py --no-e2e > gpurun_out/r17_short.json 2> gpurun_out/r17_short.err
cat gpurun_out/r17_short.json
USE=$(python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r17_short.json"))
    ok = d["device"]["compact"]["equals_ascii_kernel"] and d["device"]["compact"]["avg_ms"] < 0.93 and (d["path_flags"] & 32)
    print("1" if ok else "0")
except Exception:
    print("0")
PY
)
echo "profiling with SGDX_SHORT_SEED=$USE"
if [ "$USE" = "1" ]; then export SGDX_SHORT_SEED=1; fi
echo "$USE" > gpurun_out/r17_variant.txt
timeout 50 ncu --set full --clock-control none --import-source on -k regex:sgdx_match_compact -s 3 -c 1 -o gpurun_out/r17_compact_prof -f python tools/bench_compact.py --no-e2e --steps 2 > gpurun_out/r17_ncu.log 2>&1
tail -2 gpurun_out/r17_ncu.log
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r17_launches_compact.csv python tools/bench_compact.py --no-e2e --steps 10 > gpurun_out/r17_ncu2.log 2>&1
tail -1 gpurun_out/r17_ncu2.log
ls -la gpurun_out | grep r17
