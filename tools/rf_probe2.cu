// Operand-read cost table for packed FP32 / min instructions on B200 (register-file bandwidth probe).
// Every pattern runs 8 independent chains per thread, 8 warps per SMSP; reports clk per instruction per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define ADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define MIN3(d, a, b, c) asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MIN2(d, a, b) asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))

template <int T>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  u64 p[8], q[8], r[8]; float m[8], a[8], b[8], sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float f = in[i] + threadIdx.x;
    p[i] = pk(f, f + 1.f); q[i] = pk(f * 0.5f, f * 0.25f); r[i] = pk(f * 0.125f, 1.f - f);
    m[i] = f * 3.f; a[i] = f * 5.f; b[i] = f * 7.f; sc[i] = in[8 + (i & 3)] * (i + 1);
  }
  u64 Q = pk(in[1], in[2]), S = pk(in[3], in[3]);
  float s0 = in[4];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (T == 0) FMA2(p[i], q[i], r[i], p[i]);                       // 3 distinct pairs
      if (T == 1) FMA2(p[i], q[i], S, p[i]);                          // pair, shared bcast scalar, acc
      if (T == 2) FMA2(p[i], Q, pk(sc[i], sc[i]), p[i]);              // shared pair, distinct scalar, acc
      if (T == 3) FMA2(p[i], Q, S, p[i]);                             // acc only
      if (T == 4) ADD2(p[i], p[i], S);                                // acc + shared scalar
      if (T == 5) ADD2(p[i], p[i], q[i]);                             // two pairs
      if (T == 6) MIN3(m[i], m[i], a[i], b[i]);                       // 3 distinct regs
      if (T == 7) MIN2(m[i], m[i], a[i]);                             // 2 regs
      if (T == 8) MIN3(m[i], m[i], s0, a[i]);                         // 2 regs + shared
      if (T == 9) FMA1(m[i], a[i], b[i], m[i]);                       // scalar FMA 3 distinct
      if (T == 10) { float x, y; upk(p[i], x, y); MIN3(m[i], m[i], x, y); }   // acc + both halves of a pair
      if (T == 11) { float x, y, z, w; upk(p[i], x, y); upk(q[i], z, w); MIN3(m[i], m[i], x, z); }  // even+even
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); acc += x + y + m[i]; }
  if (acc == 123.456f) out[0] = acc;
}

template <int T> static void run(const char* name, float* d_out, const float* d_in, int sms) {
  const int iters = 4096;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<T><<<sms * 4, 256>>>(d_out, d_in, 16); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rr = 0; rr < 3; ++rr) { CK(cudaEventRecord(e0)); k<T><<<sms * 4, 256>>>(d_out, d_in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  double clk = best * 1e-3 * 1.965e9 / ((double)iters * 8 /*instr*/ * 8 /*warps per SMSP*/);
  printf("%-46s %6.2f clk per instruction per SMSP\n", name, clk);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float h_in[16]; for (int i = 0; i < 16; ++i) h_in[i] = 1.0f + 1e-3f * i;
  float *d_in, *d_out; CK(cudaMalloc(&d_in, sizeof(h_in))); CK(cudaMalloc(&d_out, 64));
  CK(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  run<0>("FFMA2 pair,pair,pair (all distinct)", d_out, d_in, sms);
  run<1>("FFMA2 pair, shared scalar, acc pair", d_out, d_in, sms);
  run<2>("FFMA2 shared pair, distinct scalar, acc pair", d_out, d_in, sms);
  run<3>("FFMA2 shared pair, shared scalar, acc pair", d_out, d_in, sms);
  run<4>("FADD2 acc pair + shared scalar", d_out, d_in, sms);
  run<5>("FADD2 acc pair + distinct pair", d_out, d_in, sms);
  run<6>("FMNMX3 3 distinct regs", d_out, d_in, sms);
  run<7>("FMNMX 2 regs", d_out, d_in, sms);
  run<8>("FMNMX3 2 regs + shared", d_out, d_in, sms);
  run<9>("FFMA scalar 3 distinct regs", d_out, d_in, sms);
  run<10>("FMNMX3 acc + lo,hi of one pair", d_out, d_in, sms);
  run<11>("FMNMX3 acc + lo,lo of two pairs", d_out, d_in, sms);
  return 0;
}
