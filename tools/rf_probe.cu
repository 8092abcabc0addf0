// Register-file pressure probe: packed FMA with distinct (non reuse-cached) operands, alone and mixed
// with 3-input min on live registers.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rf_probe rf_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

// MODE 0: 8 x fma2(p, q_i, r_i) distinct operands          (FMA only)
// MODE 1: + per fma2: one min3(m_i, lo(p_i), hi(p_i))         (col-style: even+odd)
// MODE 2: + per fma2: one min3(m_i, lo(p_i), lo(p_{i+1}))     (row-style: even+even)
// MODE 3: like the real loop: 3 fma2 (broadcast scalar) + add2 + 2 min3 per item
// MODE 4: MODE 3 but mins are 2-input (4 per item)
// MODE 5: MODE 3 without any min (arithmetic only)
// MODE 6: MODE 3 with row-style min3 replaced by (lo, hi of swapped)  i.e. no same-bank pairs
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  u64 p[8], q[8], r[8]; float m[16]; float s[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float f = in[i] + threadIdx.x; p[i] = pk(f, f + 1.f); q[i] = pk(f * 0.5f, f * 0.25f); r[i] = pk(f * 0.125f, 1.f - f); }
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 1e30f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = in[8 + i];
  for (int it = 0; it < iters; ++it) {
    if (MODE <= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i]), "l"(r[i]));
      }
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float a, b; upk(p[i], a, b); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a), "f"(b)); }
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float a, b, c, d; upk(p[i], a, b); upk(p[(i + 1) & 7], c, d); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a), "f"(c)); }
      }
    } else {
      // two columns (s[0..3] stand in for x,y,z,c of a column; a second column reuses them shifted)
      u64 e0[8], e1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        u64 t = r[i];
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(p[i]), "l"(pk(s[0], s[0])));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(q[i]), "l"(pk(s[1], s[1])));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(r[i]), "l"(pk(s[2], s[2])));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(t) : "l"(pk(s[3], s[3])));
        e0[i] = t;
        u64 t2 = r[i];
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t2) : "l"(p[i]), "l"(pk(s[1], s[1])));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t2) : "l"(q[i]), "l"(pk(s[2], s[2])));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t2) : "l"(r[i]), "l"(pk(s[3], s[3])));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(t2) : "l"(pk(s[0], s[0])));
        e1[i] = t2;
      }
      float cm0 = 1e30f, cm1 = 1e30f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a0, b0, a1, b1; upk(e0[i], a0, b0); upk(e1[i], a1, b1);
        if (MODE == 3) {
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[2 * i]) : "f"(a0), "f"(a1));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[2 * i + 1]) : "f"(b0), "f"(b1));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(cm0) : "f"(a0), "f"(b0));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(cm1) : "f"(a1), "f"(b1));
        }
        if (MODE == 4) {
          asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * i]) : "f"(a0)); asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * i]) : "f"(a1));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * i + 1]) : "f"(b0)); asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * i + 1]) : "f"(b1));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(cm0) : "f"(a0)); asm volatile("min.f32 %0, %0, %1;" : "+f"(cm0) : "f"(b0));
          asm volatile("min.f32 %0, %0, %1;" : "+f"(cm1) : "f"(a1)); asm volatile("min.f32 %0, %0, %1;" : "+f"(cm1) : "f"(b1));
        }
        if (MODE == 5) { m[2 * i] += a0 * 1e-30f; }
        if (MODE == 6) {
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[2 * i]) : "f"(a0), "f"(b1));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[2 * i + 1]) : "f"(b0), "f"(a1));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(cm0) : "f"(a0), "f"(b0));
          asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(cm1) : "f"(a1), "f"(b1));
        }
      }
      s[0] += cm0 * 1e-30f; s[1] += cm1 * 1e-30f;
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; upk(p[i], a, b); acc += a + b; }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += m[i];
  acc += s[0] + s[1];
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE> static void run(const char* name, float* d_out, const float* d_in, int sms, int warps_per_smsp, double per) {
  const int iters = 2048;
  int ctas = sms * warps_per_smsp / 2;   // 256 threads = 8 warps = 2 per SMSP
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE><<<ctas, 256>>>(d_out, d_in, 16); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); k<MODE><<<ctas, 256>>>(d_out, d_in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  double clk = best * 1e-3 * 1.965e9 / ((double)iters * warps_per_smsp);
  printf("%-52s warps/SMSP=%d  clk per body per warp-slot = %7.2f   per unit = %6.2f\n", name, warps_per_smsp, clk, clk / per);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float h_in[16]; for (int i = 0; i < 16; ++i) h_in[i] = 1.0f + 1e-3f * i;
  float *d_in, *d_out; CK(cudaMalloc(&d_in, sizeof(h_in))); CK(cudaMalloc(&d_out, 64));
  CK(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  for (int w = 2; w <= 8; w *= 2) {
    run<0>("8 fma2 distinct operands (per fma2)", d_out, d_in, sms, w, 8);
    run<1>("8 fma2 + 8 min3(lo,hi) (per fma2)", d_out, d_in, sms, w, 8);
    run<2>("8 fma2 + 8 min3(lo,lo') (per fma2)", d_out, d_in, sms, w, 8);
    run<5>("loop: 16 items arithmetic only (per item)", d_out, d_in, sms, w, 16);
    run<3>("loop: 16 items, 3 fma2+add2+2 min3 (per item)", d_out, d_in, sms, w, 16);
    run<6>("loop: same, row min3 on (lo,hi') (per item)", d_out, d_in, sms, w, 16);
    run<4>("loop: 16 items, 2-input mins (per item)", d_out, d_in, sms, w, 16);
  }
  return 0;
}
