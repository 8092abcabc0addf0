// Which instruction classes overlap on a B200 SM sub-partition?  Independent streams A and B are
// interleaved 1:1 (no data dependence between them); the time of the mix is compared with A alone and
// B alone.  If mix ~= max(A, B) the two classes use different issue/execute resources; if
// mix ~= A + B they are serialised.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_overlap_probe pipe_overlap_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }

enum { A_NONE, A_FFMA2, A_FADD2, A_FMUL2, A_FFMA };
enum { B_NONE, B_FMNMX, B_FMNMX3, B_IADD, B_LOP, B_FADD, B_FADD2, B_IMNMX, B_FFMA, B_FSETP, B_HMNMX2 };

template <int A, int B, int RA, int RB>   // RA A-ops and RB B-ops per group, 8 groups per inner body
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  float s0 = in[0], s1 = in[1];
  u64 p[8]; float f[8]; int n[8]; unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = in[2 + i] + threadIdx.x; p[i] = pk(f[i], f[i] + 1.f); n[i] = threadIdx.x + i; h[i] = threadIdx.x * 7 + i; }
  u64 ps0 = pk(s0, s0), ps1 = pk(s1, s1);
  int is0 = (int)in[3];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        if (A == A_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps0), "l"(ps1));
        if (A == A_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps0));
        if (A == A_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps0));
        if (A == A_FFMA) { float t = lo(p[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t) : "f"(s0), "f"(s1)); p[i] = pk(t, t); }
      }
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        if (B == B_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(s0));
        if (B == B_FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(s0), "f"(s1));
        if (B == B_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(is0));
        if (B == B_LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(is0));
        if (B == B_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(s0));
        if (B == B_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(s0), "f"(s1));
        if (B == B_IMNMX) asm volatile("min.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(is0));
        if (B == B_FADD2) { u64 t = pk(f[i], f[i]); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(t) : "l"(ps0)); f[i] = lo(t); }
        if (B == B_FSETP) { int t; asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; selp.s32 %0, 1, 0, q;}" : "=r"(t) : "f"(f[i]), "f"(s0)); n[i] += t; }
        if (B == B_HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += f[i] + lo(p[i]) + (float)n[i] + (float)h[i];
  if (acc == 123.456f) out[0] = acc;
}

template <int A, int B, int RA, int RB>
static double run(float* d_out, const float* d_in, int sms) {
  const int iters = 2048;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<A, B, RA, RB><<<sms * 4, 256>>>(d_out, d_in, 16); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0)); k<A, B, RA, RB><<<sms * 4, 256>>>(d_out, d_in, iters); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  // cycles per group (RA A-ops + RB B-ops) per SMSP:  8 warps per SMSP, 8 groups per iter
  double clk = 1.965e9;
  return best * 1e-3 * clk / ((double)iters * 8 /*groups*/ * 8 /*warps per smsp*/);
}

#define ROW(name, A, B, RA, RB) do { \
  double a = RA ? run<A, B_NONE, RA, 0>(d_out, d_in, sms) : 0, b = RB ? run<A_NONE, B, 0, RB>(d_out, d_in, sms) : 0, m = run<A, B, RA, RB>(d_out, d_in, sms); \
  printf("%-34s A=%6.2f  B=%6.2f  mix=%6.2f  (sum %6.2f, max %6.2f) clk per group per SMSP\n", name, a, b, m, a + b, a > b ? a : b); } while (0)

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float h_in[16]; for (int i = 0; i < 16; ++i) h_in[i] = 1.0f + 1e-3f * i;
  float *d_in, *d_out; CK(cudaMalloc(&d_in, sizeof(h_in))); CK(cudaMalloc(&d_out, 64));
  CK(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  ROW("2 FFMA2 + 1 FMNMX", A_FFMA2, B_FMNMX, 2, 1);
  ROW("2 FFMA2 + 2 FMNMX", A_FFMA2, B_FMNMX, 2, 2);
  ROW("2 FFMA2 + 1 FMNMX3", A_FFMA2, B_FMNMX3, 2, 1);
  ROW("2 FFMA2 + 2 FMNMX3", A_FFMA2, B_FMNMX3, 2, 2);
  ROW("2 FFMA2 + 2 IADD", A_FFMA2, B_IADD, 2, 2);
  ROW("2 FFMA2 + 2 LOP", A_FFMA2, B_LOP, 2, 2);
  ROW("2 FFMA2 + 2 IMNMX", A_FFMA2, B_IMNMX, 2, 2);
  ROW("2 FFMA2 + 2 FADD", A_FFMA2, B_FADD, 2, 2);
  ROW("2 FFMA2 + 2 FFMA", A_FFMA2, B_FFMA, 2, 2);
  ROW("2 FFMA2 + 2 FSETP/SEL", A_FFMA2, B_FSETP, 2, 2);
  ROW("2 FFMA2 + 2 HMNMX2", A_FFMA2, B_HMNMX2, 2, 2);
  ROW("2 FADD2 + 2 FMNMX", A_FADD2, B_FMNMX, 2, 2);
  ROW("2 FMUL2 + 2 FMNMX", A_FMUL2, B_FMNMX, 2, 2);
  ROW("2 FADD2 + 2 FMNMX3", A_FADD2, B_FMNMX3, 2, 2);
  ROW("4 FFMA2 + 1 FMNMX", A_FFMA2, B_FMNMX, 4, 1);
  ROW("4 FFMA2 + 2 FMNMX", A_FFMA2, B_FMNMX, 4, 2);
  ROW("4 FFMA2 + 4 FMNMX", A_FFMA2, B_FMNMX, 4, 4);
  return 0;
}
