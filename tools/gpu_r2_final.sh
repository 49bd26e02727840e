#!/bin/bash
# Round-2 evidence on one B200: bench lines (both arms), launch lists, full ncu captures of the hot kernels (C2, 2 GiB segment).
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" > gpurun_out/final.txt
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; echo "ref rc=$?" >> gpurun_out/final.txt
timeout 300 python tools/variants.py --mib 2048 --set c2 --passes 2 --variants new > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c2_2GiB.csv python tools/variants.py --mib 2048 --set c2 --passes 2 --variants new > gpurun_out/ncu1.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_stream|k_verify_smem|k_emit_nlm' -s 3 -c 3 -f -o gpurun_out/r2_c2_kernels python tools/variants.py --mib 2048 --set c2 --passes 2 --variants new > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?" >> gpurun_out/final.txt
# launch list of the bench command itself (device-resident passes first)
timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --cpu-seconds 1 > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-extras --cpu-seconds 1 > gpurun_out/ncu3.log 2>&1
echo "bench launch list rc=$?" >> gpurun_out/final.txt
cat gpurun_out/final.txt; tail -1 gpurun_out/r2_bench_n1.json | cut -c1-400; tail -1 gpurun_out/r2_bench_ref_n1.json | cut -c1-300
