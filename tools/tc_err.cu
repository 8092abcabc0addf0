// Error of the tensor-core hot value (chamfer_tc.cu: fp16 2-way split, tcgen05.mma kind::f16) against the exact
// squared distance, over random scaled tiles:   max |D - |P-T|^2| / (|P|+|T|)^2   and the absolute error for small
// magnitudes (fp16 subnormal low parts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../volumetric-primitives-net_b200/csrc -o tc_err tc_err.cu
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "chamfer_tc.cu"
void vpn_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
int vpn_check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -1; } return 0; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
using namespace vpn;

__global__ void __launch_bounds__(128) err_kernel(const float* __restrict__ Pp, const float* __restrict__ Tp, float* __restrict__ D) {
  __shared__ __align__(1024) unsigned char rows[kTcBlkBytes];
  __shared__ __align__(1024) unsigned char cols[2 * kTcBlkBytes];
  __shared__ __align__(8) u64 bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(tc_smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { tc_mbar_init(tc_smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  tc_make_operand(rows, tid, Pp[3 * tid], Pp[3 * tid + 1], Pp[3 * tid + 2], true);
  for (int j = tid; j < 256; j += 128) tc_make_operand(cols + (j >> 7) * kTcBlkBytes, j & 127, Tp[3 * j], Tp[3 * j + 1], Tp[3 * j + 2], false);
  tc_fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    if (tc_elect()) { tc_mma(tb, tc_desc(tc_smem_u32(rows)), tc_desc(tc_smem_u32(cols)), 0); tc_commit(tc_smem_u32(&bar)); }
    __syncwarp();
  }
  tc_mbar_wait(tc_smem_u32(&bar), 0);
  tc_fence_after();
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 256; c += 32) {
    float v[32];
    tc_ld32(tl + c, v); tc_wait_ld();
    for (int k = 0; k < 32; ++k) D[tid * 256 + c + k] = v[k];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tb) : "memory");
}

static double urand() { return (double)rand() / RAND_MAX; }
int main() {
  float *dP, *dT, *dD; CK(cudaMalloc(&dP, 128 * 3 * 4)); CK(cudaMalloc(&dT, 256 * 3 * 4)); CK(cudaMalloc(&dD, 128 * 256 * 4));
  std::vector<float> hP(128 * 3), hT(256 * 3), hD(128 * 256);
  srand(7);
  // (row radius, column radius) in scaled units; the kernel guarantees < 128
  const double cases[][2] = {{127, 127}, {100, 1}, {1, 100}, {8, 8}, {1, 1}, {0.2, 0.2}, {0.01, 0.2}, {0.01, 0.01}, {1e-3, 1e-3}, {1e-3, 100}, {100, 1e-3}, {1e-5, 1e-5}};
  double worst_rel = 0, worst_abs_small = 0;
  for (auto& cs : cases) {
    double wrel = 0, wabs = 0;
    for (int rep = 0; rep < 20; ++rep) {
      for (int i = 0; i < 128; ++i) { double r = cs[0] * pow(urand(), 0.5) / sqrt(3.0); for (int c = 0; c < 3; ++c) hP[3 * i + c] = (float)(r * (2 * urand() - 1)); }
      for (int j = 0; j < 256; ++j) { double r = cs[1] * pow(urand(), 0.5) / sqrt(3.0); for (int c = 0; c < 3; ++c) hT[3 * j + c] = (float)(r * (2 * urand() - 1)); }
      CK(cudaMemcpy(dP, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dT, hT.data(), hT.size() * 4, cudaMemcpyHostToDevice));
      err_kernel<<<1, 128>>>(dP, dT, dD); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      for (int i = 0; i < 128; ++i) for (int j = 0; j < 256; ++j) {
        double d = 0, np = 0, nt = 0;
        for (int c = 0; c < 3; ++c) { double a = hP[3 * i + c], t = hT[3 * j + c]; d += (a - t) * (a - t); np += a * a; nt += t * t; }
        const double e = fabs((double)hD[i * 256 + j] - d), s = sqrt(np) + sqrt(nt);
        if (s >= 0.25) wrel = fmax(wrel, e / (s * s)); else wabs = fmax(wabs, e);
      }
    }
    printf("radii (%g, %g): max rel err (|P|+|T| >= 1/4) = 2^%.2f   max abs err (|P|+|T| < 1/4) = 2^%.2f\n", cs[0], cs[1],
           wrel > 0 ? log2(wrel) : -99.0, wabs > 0 ? log2(wabs) : -99.0);
    worst_rel = fmax(worst_rel, wrel); worst_abs_small = fmax(worst_abs_small, wabs);
  }
  printf("worst: rel 2^%.2f (assumed 2^-19), abs-small 2^%.2f (assumed 2^-23)\n", log2(worst_rel), worst_abs_small > 0 ? log2(worst_abs_small) : -99.0);
  return 0;
}
