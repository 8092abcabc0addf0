// FP32-pipe peak probe.  MEASURED_PEAKS.json has HBM and bf16 tensor peaks only; the Chamfer kernel
// is bound by the FP32 CUDA-core pipe, so bench.py measures that roofline denominator live with
// this probe: independent FFMA (1 FMA / lane / issue) and FFMA2 (2 FMA / lane / issue) streams.
#include "common.cuh"

namespace vpn {

__device__ __forceinline__ u64 ppk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int PACKED>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, const float* in, int iters) {
  float s0 = in[0], s1 = in[1];
  float f[8]; u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = in[2 + i] + threadIdx.x; p[i] = ppk(f[i], f[i] + 1.f); }
  u64 ps0 = ppk(s0, s0), ps1 = ppk(s1, s1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
        else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(s0), "f"(s1));
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); acc += f[i] + a + b; }
  if (acc == 123.456f) out[0] = acc;
}

}  // namespace vpn

// scratch: >= 64 floats of device memory whose first 16 hold finite values near 1.0.
// Returns achieved TFLOP/s (FMA = 2 flop) for scalar FFMA and packed FFMA2 streams, best of `reps`.
extern "C" int vpn_fp32_peak_probe(float* scratch, int reps, double* tflops_ffma, double* tflops_ffma2, void* stream) {
  if (!scratch || !tflops_ffma || !tflops_ffma2 || reps < 1) { vpn_set_error("fp32 probe: bad arguments"); return VPN_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { vpn_set_error("fp32 probe: event create failed"); return VPN_ERR_CUDA; }
  const int iters = 2048;
  dim3 grid(sms * 4), block(256);
  double best[2] = {0.0, 0.0};
  for (int packed = 0; packed < 2; ++packed) {
    for (int r = 0; r < reps + 1; ++r) {
      cudaEventRecord(e0, s);
      if (packed) vpn::fp32_probe_kernel<1><<<grid, block, 0, s>>>(scratch + 32, scratch, iters);
      else vpn::fp32_probe_kernel<0><<<grid, block, 0, s>>>(scratch + 32, scratch, iters);
      cudaEventRecord(e1, s);
      if (cudaEventSynchronize(e1) != cudaSuccess) { vpn_set_error("fp32 probe: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return VPN_ERR_CUDA; }
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)grid.x * block.x * (double)iters * 32.0 * (packed ? 2.0 : 1.0);
      double tf = fma * 2.0 / (ms * 1e-3) / 1e12;
      if (r > 0 && tf > best[packed]) best[packed] = tf;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops_ffma = best[0]; *tflops_ffma2 = best[1];
  return VPN_OK;
}
