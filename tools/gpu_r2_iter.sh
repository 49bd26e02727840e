#!/bin/bash
# Round-2 measurement iteration: parity tests, then A/B variants of the fast-path kernels for every pattern set, then (optional,
# "$1" = launches) the per-kernel launch list of the C2 variants under ncu.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?" > gpurun_out/iter.txt
tail -5 gpurun_out/tests.log
for set in c2 c1 c3 lit; do
  timeout 600 python tools/variants.py --mib 2048 --set $set ${VARIANTS:+--variants $VARIANTS} > gpurun_out/var_$set.log 2>&1; echo "variants $set rc=$?" >> gpurun_out/iter.txt
  cat gpurun_out/var_$set.log | grep -E "variant=|Error|error" | head -20
done
if [ "$1" = "launches" ]; then
  timeout 300 python tools/variants.py --mib 2048 --set c2 --passes 2 > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_variants.csv python tools/variants.py --mib 2048 --set c2 --passes 2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/iter.txt
fi
cat gpurun_out/iter.txt
