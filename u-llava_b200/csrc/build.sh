#!/usr/bin/env bash
# Builds libullava_sm100.so in-tree (next to this script) for sm_100a only.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
SRCS="runtime.cu gemm_sm100.cu gemm_stream_sm100.cu norm.cu attention.cu fmha_sm100.cu fmha_window_sm100.cu elementwise.cu sampling.cu eval_ops.cu preprocess.cu sam_decoder.cu sam_encoder.cu models.cu partition.cu gemm_chain_sm100.cu capi.cu"
mkdir -p build
pids=()
for f in $SRCS; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ gemm_stream.cuh -nt "$o" ] || [ ullava_internal.h -nt "$o" ] || [ ../../include/ullava_sm100.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=$(for f in $SRCS; do echo build/${f%.cu}.o; done)
$NVCC -shared -o libullava_sm100.so $OBJS -Xcompiler -fPIC -cudart static
echo "built $(pwd)/libullava_sm100.so"
