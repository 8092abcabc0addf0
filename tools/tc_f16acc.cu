// Probe: tcgen05.mma kind::f16 with an FP16 accumulator (instruction descriptor c_format = 0) for the Chamfer filter.
//   1. where the f16 results sit in TMEM (raw 32x32b read) and what tcgen05.ld ... .pack::16b returns;
//   2. error of the f16 hot value against the exact squared distance: relative to the value itself (one final rounding
//      would give 2^-11) and relative to (|P|+|T|)^2 (the accumulation error the f32 accumulator has, 2^-19.85);
//   3. epilogue rate with all 16 epilogue warps of a CTA: 128 f32 columns + 64 FMNMX3 against 64 packed registers +
//      VHMNMX (3-input half2 min), clk per 128 x 128 block per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../volumetric-primitives-net_b200/csrc -o tc_f16acc tc_f16acc.cu
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "chamfer_tc.cu"
void vpn_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
int vpn_check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -1; } return 0; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
using namespace vpn;

constexpr uint32_t kIdescF16Acc = (0u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcBlk >> 4) << 24);   // D f16, N = 128

__device__ __forceinline__ void ld32_pack(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}

// D32: f32 accumulator (128 x 128), RAW: the f16 accumulator's TMEM columns as 32-bit words (128 x 128),
// PK: the same region read with .pack::16b (128 x 64 words)
__global__ void __launch_bounds__(128) err_kernel(const float* __restrict__ Pp, const float* __restrict__ Tp, float* __restrict__ D32,
                                                   uint32_t* __restrict__ RAW, uint32_t* __restrict__ PK) {
  __shared__ __align__(1024) unsigned char rows[kTcBlkBytes];
  __shared__ __align__(1024) unsigned char cols[kTcBlkBytes];
  __shared__ __align__(8) u64 bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(tc_smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { tc_mbar_init(tc_smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  tc_make_operand(rows, tid, Pp[3 * tid], Pp[3 * tid + 1], Pp[3 * tid + 2], true);
  tc_make_operand(cols, tid, Tp[3 * tid], Tp[3 * tid + 1], Tp[3 * tid + 2], false);
  tc_fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    if (tc_elect()) {
      tc_mma(tb, tc_desc(tc_smem_u32(rows)), tc_desc(tc_smem_u32(cols)), 0, kTcIdescHalf);
      tc_mma(tb + 128, tc_desc(tc_smem_u32(rows)), tc_desc(tc_smem_u32(cols)), 0, kIdescF16Acc);
      tc_commit(tc_smem_u32(&bar));
    }
    __syncwarp();
  }
  tc_mbar_wait(tc_smem_u32(&bar), 0);
  tc_fence_after();
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    tc_ld32(tl + c, v); tc_wait_ld();
    for (int k = 0; k < 32; ++k) D32[tid * 128 + c + k] = v[k];
    tc_ld32(tl + 128 + c, v); tc_wait_ld();
    for (int k = 0; k < 32; ++k) RAW[tid * 128 + c + k] = __float_as_uint(v[k]);
  }
  for (int c = 0; c < 128; c += 64) {                     // 64 columns -> 32 packed registers
    uint32_t r[32];
    ld32_pack(tl + 128 + c, r); tc_wait_ld();
    for (int k = 0; k < 32; ++k) PK[tid * 64 + c / 2 + k] = r[k];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tb) : "memory");
}

// ---- epilogue rate: 16 warps, each reading its lane quarter of a 128-column accumulator buffer again and again
__device__ __forceinline__ uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r, t;
  asm("min.s16x2 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(t), "r"(c));      // ptxas fuses the pair into VIMNMX3.S16x2
  return r;
}
__device__ __forceinline__ uint32_t hmin2_(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmin2n_(uint32_t a, uint32_t b) { uint32_t r; asm("min.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// 16 packed registers -> 1 with 2-input half2 mins only (alternating NaN modes keep ptxas from fusing them into VHMNMX)
__device__ __forceinline__ uint32_t hmin16(const uint32_t* v) {
  const uint32_t t0 = hmin2_(v[0], v[1]), t1 = hmin2_(v[2], v[3]), t2 = hmin2_(v[4], v[5]), t3 = hmin2_(v[6], v[7]);
  const uint32_t t4 = hmin2_(v[8], v[9]), t5 = hmin2_(v[10], v[11]), t6 = hmin2_(v[12], v[13]), t7 = hmin2_(v[14], v[15]);
  const uint32_t s0 = hmin2n_(t0, t1), s1 = hmin2n_(t2, t3), s2 = hmin2n_(t4, t5), s3 = hmin2n_(t6, t7);
  return hmin2n_(hmin2_(s0, s1), hmin2_(s2, s3));
}
__device__ __forceinline__ uint32_t imin16(const uint32_t* s) {
  uint32_t x = s[0];
#pragma unroll
  for (int k = 1; k < 15; k += 2) x = min3_s16x2(x, s[k], s[k + 1]);
  uint32_t r; asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s[15]));
  return r;
}
template <int MODE>      // 5: packed + 2-input HMNMX2 only, 6: units 0,1 VIMNMX3.S16x2 + units 2,3 HMNMX2, 7: three units integer + one half2; 0: f32 columns + FMNMX3, 1: packed f16 + 3-input half2 min, 2: packed + 3-input s16x2 min, 3: packed loads only, 4: f32 loads only
__global__ void __launch_bounds__(512, 1) rate_kernel(float* __restrict__ out, long long* __restrict__ clk, int iters) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(tc_smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const int q = warp & 3, buf = warp >> 2;                 // 4 buffers of 128 columns, lane quarter q
  const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + buf * 128;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      float u[4];
      acc += tc_lane_min4x32(tl, 0, 1, u) + u[1];           // lane != 0: no barrier arrive
    } else if (MODE == 2) {
      uint32_t a[32], b[32];
      ld32_pack(tl, a); ld32_pack(tl + 64, b);
      tc_wait_ld();
      uint32_t m[4];
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const uint32_t* s = a + 16 * uu;
        uint32_t x = s[0];
#pragma unroll
        for (int k = 1; k < 15; k += 2) x = min3_s16x2(x, s[k], s[k + 1]);
        asm("min.s16x2 %0, %1, %2;" : "=r"(m[uu]) : "r"(x), "r"(s[15]));
        const uint32_t* s2 = b + 16 * uu;
        uint32_t y = s2[0];
#pragma unroll
        for (int k = 1; k < 15; k += 2) y = min3_s16x2(y, s2[k], s2[k + 1]);
        asm("min.s16x2 %0, %1, %2;" : "=r"(m[2 + uu]) : "r"(y), "r"(s2[15]));
      }
      float u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int lo = (int)(short)(m[k] & 0xffffu), hi = (int)(short)(m[k] >> 16);
        const int w = max(min(lo, hi), 0);                                   // negative (tiny) values clamp to +0
        u[k] = __half2float(__ushort_as_half((unsigned short)w));
      }
      acc += fminf(tc_min3(u[0], u[1], u[2]), u[3]) + u[1];
    } else if (MODE == 5 || MODE == 6 || MODE == 7) {
      uint32_t a[32], b[32];
      ld32_pack(tl, a); ld32_pack(tl + 64, b);
      tc_wait_ld();
      uint32_t m[4];
      m[0] = (MODE == 5) ? hmin16(a) : imin16(a);
      m[1] = (MODE == 5) ? hmin16(a + 16) : imin16(a + 16);
      m[2] = (MODE == 7) ? imin16(b) : hmin16(b);
      m[3] = hmin16(b + 16);
      float u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&m[k]));
        u[k] = fminf(f.x, f.y);
      }
      acc += fminf(tc_min3(u[0], u[1], u[2]), u[3]) + u[1];
    } else if (MODE == 3) {
      uint32_t a[32], b[32];
      ld32_pack(tl, a); ld32_pack(tl + 64, b);
      tc_wait_ld();
      acc += __uint_as_float(a[3] ^ b[17]);
    } else if (MODE == 4) {
      float va[32], vb[32];
      tc_ld32(tl, va); tc_ld32(tl + 32, vb); tc_wait_ld();
      acc += va[1] + vb[2];
      tc_ld32(tl + 64, va); tc_ld32(tl + 96, vb); tc_wait_ld();
      acc += va[3] + vb[7];
    } else {
      uint32_t a[32], b[32];
      ld32_pack(tl, a); ld32_pack(tl + 64, b);
      tc_wait_ld();
      __half2 m[4];
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const uint32_t* s = a + 16 * uu;
        __half2 x = *reinterpret_cast<const __half2*>(&s[0]);
#pragma unroll
        for (int k = 1; k < 15; k += 2) x = __hmin2(__hmin2(x, *reinterpret_cast<const __half2*>(&s[k])), *reinterpret_cast<const __half2*>(&s[k + 1]));
        m[uu] = __hmin2(x, *reinterpret_cast<const __half2*>(&s[15]));
        const uint32_t* s2 = b + 16 * uu;
        __half2 y = *reinterpret_cast<const __half2*>(&s2[0]);
#pragma unroll
        for (int k = 1; k < 15; k += 2) y = __hmin2(__hmin2(y, *reinterpret_cast<const __half2*>(&s2[k])), *reinterpret_cast<const __half2*>(&s2[k + 1]));
        m[2 + uu] = __hmin2(y, *reinterpret_cast<const __half2*>(&s2[15]));
      }
      float u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = fminf(__low2float(m[k]), __high2float(m[k]));
      acc += fminf(tc_min3(u[0], u[1], u[2]), u[3]) + u[1];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 512 + tid] = acc;
  if (tid == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tb) : "memory");
}

static double urand() { return (double)rand() / RAND_MAX; }
int main() {
  float *dP, *dT, *dD; uint32_t *dR, *dK;
  CK(cudaMalloc(&dP, 128 * 3 * 4)); CK(cudaMalloc(&dT, 128 * 3 * 4)); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMalloc(&dR, 128 * 128 * 4)); CK(cudaMalloc(&dK, 128 * 64 * 4));
  std::vector<float> hP(128 * 3), hT(128 * 3), hD(128 * 128);
  std::vector<uint32_t> hR(128 * 128), hK(128 * 64);
  srand(7);
  const double cases[][2] = {{127, 127}, {100, 1}, {1, 100}, {60, 60}, {8, 8}, {1, 1}, {0.2, 0.2}, {0.01, 0.2}, {0.01, 0.01}, {1e-3, 1e-3}, {1e-3, 100}, {1e-5, 1e-5}};
  bool layout_shown = false;
  double worst_total = 0;
  for (auto& cs : cases) {
    double w_rel_val = 0, w_rel_mag = 0, w_abs_small = 0, w_vs32 = 0; long n_inf = 0, n_neg = 0, n_layout_bad = 0, n_hi_nonzero = 0;
    for (int rep = 0; rep < 20; ++rep) {
      for (int i = 0; i < 128; ++i) { double r = cs[0] * pow(urand(), 0.5) / sqrt(3.0); for (int c = 0; c < 3; ++c) hP[3 * i + c] = (float)(r * (2 * urand() - 1)); }
      for (int j = 0; j < 128; ++j) { double r = cs[1] * pow(urand(), 0.5) / sqrt(3.0); for (int c = 0; c < 3; ++c) hT[3 * j + c] = (float)(r * (2 * urand() - 1)); }
      if (rep & 1) for (int j = 0; j < 128; ++j) for (int c = 0; c < 3; ++c) hT[3 * j + c] = hP[3 * ((j * 7) & 127) + c] * (float)(1.0 + 0.02 * (urand() - 0.5));   // near pairs
      CK(cudaMemcpy(dP, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dT, hT.data(), hT.size() * 4, cudaMemcpyHostToDevice));
      err_kernel<<<1, 128>>>(dP, dT, dD, dR, dK); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hR.data(), dR, hR.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hK.data(), dK, hK.size() * 4, cudaMemcpyDeviceToHost));
      if (!layout_shown) {
        layout_shown = true;
        printf("row 5, columns 0..7: f32 acc | raw f16-acc words | packed words\n");
        for (int j = 0; j < 8; ++j) printf("  col %d: f32 %.6g   raw 0x%08x (lo half %.6g, hi half %.6g)\n", j, hD[5 * 128 + j], hR[5 * 128 + j],
                                          __half2float(__ushort_as_half((unsigned short)(hR[5 * 128 + j] & 0xffff))), __half2float(__ushort_as_half((unsigned short)(hR[5 * 128 + j] >> 16))));
        for (int j = 0; j < 4; ++j) printf("  packed word %d: 0x%08x (lo %.6g, hi %.6g)\n", j, hK[5 * 64 + j],
                                          __half2float(__ushort_as_half((unsigned short)(hK[5 * 64 + j] & 0xffff))), __half2float(__ushort_as_half((unsigned short)(hK[5 * 64 + j] >> 16))));
      }
      for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
        double d = 0, np = 0, nt = 0;
        for (int c = 0; c < 3; ++c) { double a = hP[3 * i + c], t = hT[3 * j + c]; d += (a - t) * (a - t); np += a * a; nt += t * t; }
        const uint32_t raw = hR[i * 128 + j];
        if (raw >> 16) ++n_hi_nonzero;
        const unsigned short hb = (unsigned short)(raw & 0xffff);
        const unsigned short pk = (unsigned short)((hK[i * 64 + j / 2] >> ((j & 1) * 16)) & 0xffff);
        if (pk != hb) ++n_layout_bad;
        const double v = (double)__half2float(__ushort_as_half(hb));
        if (std::isinf(v)) { ++n_inf; continue; }
        if (v < 0) ++n_neg;
        const double e = fabs(v - d), s = sqrt(np) + sqrt(nt);
        w_vs32 = fmax(w_vs32, fabs(v - (double)hD[i * 128 + j]) / fmax(fabs((double)hD[i * 128 + j]), 6.2e-5));
        // model: err <= u |d| + E (|P|+|T|)^2 (+ subnormal floor 2^-25); report the two coefficients separately by region
        if (s >= 0.25) { w_rel_mag = fmax(w_rel_mag, fmax(0.0, e - d * 4.8828125e-4) / (s * s)); w_rel_val = fmax(w_rel_val, fmax(0.0, e - 1.9e-6 * s * s - 3e-8) / fmax(d, 1e-30)); }
        else w_abs_small = fmax(w_abs_small, fmax(0.0, e - d * 4.8828125e-4));
      }
    }
    printf("radii (%g, %g): err - 2^-11 d over (|P|+|T|)^2 = 2^%.2f | (err - 2^-19 (|P|+|T|)^2) / d = 2^%.2f | small-magnitude abs residue 2^%.2f | vs f32 acc rel 2^%.2f | inf %ld neg %ld | pack mismatch %ld, raw hi-half nonzero %ld\n",
           cs[0], cs[1], w_rel_mag > 0 ? log2(w_rel_mag) : -99.0, w_rel_val > 0 ? log2(w_rel_val) : -99.0, w_abs_small > 0 ? log2(w_abs_small) : -99.0,
           w_vs32 > 0 ? log2(w_vs32) : -99.0, n_inf, n_neg, n_layout_bad, n_hi_nonzero);
    worst_total = fmax(worst_total, w_rel_val);
  }
  // ---- rates
  float* dout; long long* dclk; CK(cudaMalloc(&dout, 148 * 512 * 4)); CK(cudaMalloc(&dclk, 148 * 8));
  const int iters = 2000;
  const char* names[8] = {"f32 columns + FMNMX3", "packed f16 + 3-input half2 min", "packed f16 + 3-input s16x2 min", "packed loads only", "f32 loads only",
                          "packed f16 + 2-input HMNMX2", "packed f16, half VIMNMX3.S16x2 half HMNMX2", "packed f16, 3/4 VIMNMX3.S16x2 1/4 HMNMX2"};
  for (int mode = 0; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (mode) {
        case 0: rate_kernel<0><<<148, 512>>>(dout, dclk, iters); break;
        case 1: rate_kernel<1><<<148, 512>>>(dout, dclk, iters); break;
        case 2: rate_kernel<2><<<148, 512>>>(dout, dclk, iters); break;
        case 3: rate_kernel<3><<<148, 512>>>(dout, dclk, iters); break;
        case 5: rate_kernel<5><<<148, 512>>>(dout, dclk, iters); break;
        case 6: rate_kernel<6><<<148, 512>>>(dout, dclk, iters); break;
        case 7: rate_kernel<7><<<148, 512>>>(dout, dclk, iters); break;
        default: rate_kernel<4><<<148, 512>>>(dout, dclk, iters); break;
      }
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    }
    long long hc[148]; CK(cudaMemcpy(hc, dclk, sizeof(hc), cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)hc[i]; mean /= 148;
    // 16 warps = 4 buffers x 4 lane quarters: one iteration of all warps = 4 blocks of 128 x 128
    printf("epilogue rate, %s: %.1f clk per iteration per warp = %.1f clk per 128 x 128 block per SM\n",
           names[mode], mean / iters, mean / iters / 4.0);
  }
  return 0;
}
