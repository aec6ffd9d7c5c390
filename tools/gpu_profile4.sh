#!/usr/bin/env bash
# ncu evidence for one bench step (B=32, 4 new tokens) with the CTA-pair GEMM: launch list + --set full captures
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
CMD="python bench.py --profile-mode --new-tokens 4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_v5.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" | tee -a gpurun_out/summary.txt
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/details_$1.txt 2>/dev/null
}
cap pair_plain 'gemm_tcgen05_pair_kernel<__nv_bfloat16, .{0,6}0>' 100 4
cap pair_silu 'gemm_tcgen05_pair_kernel<__nv_bfloat16, .{0,6}4>' 4 1
cap pair_gelu 'gemm_tcgen05_pair_kernel<__nv_bfloat16, .{0,6}2>' 8 1
du -sh gpurun_out; ls gpurun_out
