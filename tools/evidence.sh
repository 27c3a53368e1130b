#!/bin/bash
# round-2 evidence on one GPU: ncu launch list of bench.py itself, ncu --set full of the hot kernels, randomised
# parity sweep against the unmodified reference library with its log.  Outputs in gpurun_out/ (copied to profiles/).
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"
ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:score_kernel -c 24 -o gpurun_out/r2_score python tools/prof_run.py 262144 1 c2 > gpurun_out/r2_ncu_score.log 2>&1
ncu --set full --clock-control none -k regex:band_kernel -c 8 -o gpurun_out/r2_band python tools/prof_run.py 1048576 1 c2 > gpurun_out/r2_ncu_band.log 2>&1
ncu --set full --clock-control none -k regex:tiny_kernel -c 1 -o gpurun_out/r2_tiny python tools/prof_run.py 2000000 1 s2 > gpurun_out/r2_ncu_tiny.log 2>&1
for r in score band tiny; do
  ncu -i gpurun_out/r2_$r.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2_$r.ncu-rep
done
( for fam in mixed edge both; do
    SSW_CUDA_TBAND_MIN=0 python tools/fuzz_gpu.py 6000 8 11 $fam flags
    python tools/fuzz_gpu.py 6000 6 12 $fam
  done
  python tools/fuzz_gpu.py 300 4 13 long
  SSW_CUDA_TBAND_MIN=0 python tools/fuzz_gpu.py 400 4 14 big ) > gpurun_out/r2_fuzz_gpu.log 2>/dev/null
tail -3 gpurun_out/r2_fuzz_gpu.log
grep -c "mismatches 0" gpurun_out/r2_fuzz_gpu.log; grep -v "mismatches 0" gpurun_out/r2_fuzz_gpu.log | head
ls -la gpurun_out | tail -12
