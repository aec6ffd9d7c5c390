#!/usr/bin/env bash
# ncu evidence for one bench step (B=32, 4 new tokens) with the decode-step changes of this session:
# launch list + --set full captures of the decode kernels and the dominant GEMM; reduced to CSV / text on the box.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
CMD="python bench.py --profile-mode --new-tokens 4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_v4.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?" | tee -a gpurun_out/summary.txt
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/details_$1.txt 2>/dev/null
}
cap gemm_stream 'gemm_stream_kernel' 40 4
cap attn_decode 'attn_decode_kernel' 8 1
cap gemm_prefill 'gemm_tcgen05_kernel<__nv_bfloat16, .{0,6}256, .{0,6}4, .{0,6}0>' 60 4
cap gemm_silu 'gemm_tcgen05_kernel<__nv_bfloat16, .{0,6}256, .{0,6}4, .{0,6}4>' 4 1
du -sh gpurun_out; ls gpurun_out
