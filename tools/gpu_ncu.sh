#!/bin/bash
# ncu --set full capture (with source) of the kernels matching $1, from one 1 GiB pass over the C2 set.
mkdir -p gpurun_out
timeout 300 python tools/profile_run.py --mib 1024 --set ${2:-c2} > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-2} -c ${4:-2} -f -o gpurun_out/capture python tools/profile_run.py --mib 1024 --set ${2:-c2} > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu.log
