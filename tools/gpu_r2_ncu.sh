#!/bin/bash
# ncu --set full capture (with source) of the kernels matching $1 from tools/variants.py ($2 = variants, $3 = set, $4 = MiB)
mkdir -p gpurun_out
V=${2:-new}; S=${3:-c2}; M=${4:-1024}
timeout 300 python tools/variants.py --mib $M --set $S --passes 2 --variants $V > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -f -o gpurun_out/capture python tools/variants.py --mib $M --set $S --passes 2 --variants $V > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu.log
