#!/bin/bash
# First GPU bring-up: environment probe, golden + parity tests, smoke, short bench.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
  echo "nproc=$(nproc)"; free -g | head -2
  ldconfig -p | grep -E 'libhs|vectorscan' || echo "no libhs"
  python -c "import hyperscan" 2>&1 | tail -1
  ls baseline/_ref 2>&1 | head -3
  df -h /dev/shm | tail -1
} > gpurun_out/probe.txt 2>&1
timeout 600 python -m pytest tests/test_golden_gpu.py -x -q > gpurun_out/golden.log 2>&1; echo "golden rc=$?" >> gpurun_out/probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/probe.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/probe.txt
timeout 600 python bench.py --gib 2 --steps 3 --warmup 3 > gpurun_out/bench_2g.log 2>&1; echo "bench rc=$?" >> gpurun_out/probe.txt
cat gpurun_out/probe.txt; tail -5 gpurun_out/golden.log; tail -25 gpurun_out/parity.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_2g.log
