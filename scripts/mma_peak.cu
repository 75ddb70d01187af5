// Micro-benchmark: peak issue rate of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16 -> fp32) on this GPU,
// as a function of resident warps per SM.  It is the ceiling of skp_selfattn.cu (mma.sync flash kernels); the tcgen05
// GEMM's ceiling is MEASURED_PEAKS.json.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mma_peak scripts/mma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ACC>
__global__ void mma_loop(float* out, int iters) {
  float c[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 4);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int iters = 4096;
  for (int warps_per_sm : {4, 8, 12, 16, 32}) {
    int threads = 128, ctas = p.multiProcessorCount * (warps_per_sm / 4);
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    mma_loop<8><<<ctas, threads>>>(out, iters);
    cudaEventRecord(s);
    mma_loop<8><<<ctas, threads>>>(out, iters);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    double flops = (double)ctas * 4 * iters * 8 * 4096.0;
    double clk_per_mma_smsp = (ms * 1e-3 * p.clockRate * 1e3) / ((double)iters * 8 * (warps_per_sm / 4.0));
    printf("{\"bench\": \"mma.sync m16n8k16 bf16\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.1f, \"clk_per_mma_per_smsp_at_max_clock\": %.2f}\n",
           warps_per_sm, ms, flops / ms / 1e9, clk_per_mma_smsp);
  }
  return 0;
}
