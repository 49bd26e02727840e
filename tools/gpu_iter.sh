#!/bin/bash
# One measurement iteration on the GPU box: parity tests, then per-kernel timing of one 2 GiB segment per pattern set.
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?" > gpurun_out/iter.txt
for set in c1 c2 c3 lit; do
  timeout 300 python tools/profile_run.py --mib 2048 --set $set --passes 3 > gpurun_out/prof_$set.log 2>&1; echo "prof $set rc=$?" >> gpurun_out/iter.txt
done
if [ "$1" = "launches" ]; then
  timeout 300 python tools/profile_run.py --mib 2048 --set c2 > gpurun_out/plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/profile_run.py --mib 2048 --set c2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/iter.txt
fi
cat gpurun_out/iter.txt; tail -3 gpurun_out/tests.log
for s in c1 c2 c3 lit; do tail -1 gpurun_out/prof_$s.log; done
for n in 1000 10000; do timeout 600 python tools/profile_run.py --mib 1024 --set c5 --patterns $n --passes 3 2>&1 | tail -1; done
