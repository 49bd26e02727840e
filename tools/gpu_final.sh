#!/bin/bash
# Round-end evidence: bench line (N=1), launch list and one full ncu capture per hot kernel for the C2 set (2 GiB segment).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" > gpurun_out/final.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/final.txt
timeout 300 python tools/profile_run.py --mib 2048 --set c2 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_run.py --mib 2048 --set c2 > gpurun_out/ncu1.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stream|k_verify_local|k_emit_simple' -s 3 -c 3 -f -o gpurun_out/r1_c2_kernels python tools/profile_run.py --mib 2048 --set c2 > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?" >> gpurun_out/final.txt
# launch list of the bench command itself (first 400 launches: synthetic-text generation is host code, so these are the
# device-resident passes followed by the first end-to-end segments)
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/ncu3.log 2>&1
echo "bench launch list rc=$?" >> gpurun_out/final.txt
cat gpurun_out/final.txt; tail -1 gpurun_out/bench_n1.json | cut -c1-600; tail -1 gpurun_out/bench_ref.json | cut -c1-400
