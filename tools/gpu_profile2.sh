#!/usr/bin/env bash
# ncu evidence for one bench step (B=32, 4 new tokens): launch list + --set full captures of the hot kernels.
# Reports are reduced to CSV / text on the box (gpurun_out is capped at 64 MiB); only two .ncu-rep files are kept.
set -u
mkdir -p gpurun_out
CMD="python bench.py --profile-mode --new-tokens 4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_v3.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
cap() {  # name regex skip count keep
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" | tee -a gpurun_out/summary.txt
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/details_$1.txt 2>/dev/null
  if [ "$5" = keep ]; then cp /tmp/prof_$1.ncu-rep gpurun_out/; fi
}
cap fmha_llama 'fmha_tcgen05_kernel<__nv_bfloat16, .{0,6}128' 8 1 drop
cap fmha_sam_global 'fmha_tcgen05_kernel<__nv_bfloat16, .{0,6}80' 1 1 keep
cap fmha_window 'fmha_window_kernel' 8 1 keep
cap gemm_prefill 'gemm_tcgen05_kernel<__nv_bfloat16, .{0,6}256, .{0,6}4, .{0,6}0>' 60 4 drop
cap gemm_gelu 'gemm_tcgen05_kernel<__nv_bfloat16, .{0,6}256, .{0,6}4, .{0,6}2>' 8 1 drop
cap gemm_stream 'gemm_stream_kernel' 40 4 drop
du -sh gpurun_out; ls gpurun_out
