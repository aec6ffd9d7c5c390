// Issue rate of the instructions an attention softmax is made of, per SM sub-partition (B200, sm_100a).
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates tools/micro/pipe_rates.cu && /tmp/pipe_rates
// Each kernel runs ITER x 8 independent instructions of one kind per thread; prints cycles per warp-instruction per
// sub-partition with W warps resident on it (1 CTA per SM, 128 * W threads).
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITER = 512;

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (OP == 0) {   // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if constexpr (OP == 1) {   // F2FP.BF16.F32.PACK_AB
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        acc ^= r;   // LOP3 on the ALU pipe, measured separately as OP 6
      } else if constexpr (OP == 2) {   // F2FP.F16.F32.PACK_AB
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        acc ^= r;
      } else if constexpr (OP == 3) {   // FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      } else if constexpr (OP == 4) {   // FFMA2 (two per instruction)
        if (i < 4) {
          uint64_t v;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
          asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v));
        }
      } else if constexpr (OP == 5) {   // FMNMX
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      } else if constexpr (OP == 6) {   // LOP3
        acc ^= __float_as_uint(a[i]) + it;
      } else if constexpr (OP == 7) {   // PRMT
        uint32_t r;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(a[i])), "r"(acc));
        acc = r;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4}) {
    k<OP><<<148, 128 * w>>>(out, cyc, 0.5f);
    k<OP><<<148, 128 * w>>>(out, cyc, 0.5f);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP %d: %6.2f cycles per warp-instruction per SMSP\n", name, w,
           double(c) / (double(ITER) * per_iter * w));
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 8);
  run<1>("F2FP.BF16.PACK_AB (+LOP3)", 8);
  run<2>("F2FP.F16.PACK_AB (+LOP3)", 8);
  run<3>("FFMA", 8);
  run<4>("FFMA2", 4);
  run<5>("FMNMX", 8);
  run<6>("IADD+LOP3 (2 instr)", 8);
  run<7>("PRMT (dependent chain)", 8);
  return 0;
}
