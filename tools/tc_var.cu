// Timing harness for chamfer_tc_kernel variants (VPN_TC_VARIANT: 0 product, 1 no reduction, 2 no TMEM reads, 3 no MMA):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DVPN_TC_VARIANT=1 -I../volumetric-primitives-net_b200/csrc -o tc_var1 tc_var.cu
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "chamfer_tc.cu"
void vpn_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
int vpn_check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -1; } return 0; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 32, P = 65536, M = 8192, NB = argc > 2 ? atoi(argv[2]) : 16;
  const int TM = NB * 128, ntiles = (P + TM - 1) / TM, nchunks = M / 128, nsplit = 1, cps = nchunks;
  std::vector<float> h1((size_t)B * P * 3), h2((size_t)B * M * 3);
  srand(1);
  for (int b = 0; b < B; ++b) for (int i = 0; i < P; ++i) { float cx = ((i / 4096) % 4) * 0.2f - 0.3f; for (int c = 0; c < 3; ++c) h1[((size_t)b * P + i) * 3 + c] = cx + 0.1f * rand() / RAND_MAX; }
  for (auto& x : h2) x = (float)rand() / RAND_MAX - 0.5f;
  float *p1, *p2, *rbest, *cbest; vpn::u64* rmask; unsigned* cmask; float2* tslack; int* fb; float* tmax;
  CK(cudaMalloc(&p1, h1.size() * 4)); CK(cudaMalloc(&p2, h2.size() * 4));
  CK(cudaMalloc(&rbest, (size_t)B * P * 4)); CK(cudaMalloc(&rmask, (size_t)B * P * 8));
  CK(cudaMalloc(&cbest, (size_t)B * ntiles * M * 4)); CK(cudaMalloc(&cmask, (size_t)B * ntiles * M * 4));
  CK(cudaMalloc(&tslack, (size_t)B * ntiles * 8)); CK(cudaMalloc(&fb, B * 4)); CK(cudaMalloc(&tmax, B * 4)); CK(cudaMemset(fb, 0, B * 4));
  CK(cudaMemcpy(p1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(p2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    if (vpn::chamfer_tc_launch(p1, p2, rbest, rmask, cbest, cmask, tslack, fb, tmax, B, P, M, NB, ntiles, nsplit, nchunks, cps, 0)) return 1;
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double stages = (double)B * ntiles * nchunks * NB / 148.0;      // 256-column stages per SM (both phases)
#ifndef VPN_TC_VARIANT
#define VPN_TC_VARIANT 0
#endif
    printf("variant %d rep %d: %.3f ms = %.1f clk per stage per SM\n", VPN_TC_VARIANT, rep, ms, ms * 1e-3 * 1.965e9 / stages);
  }
#ifdef VPN_TC_TRACE
  long long tr[32 * 8]; CK(cudaMemcpyFromSymbol(tr, vpn::g_tc_trace, sizeof(tr)));
  long long t0 = tr[0];
  printf("stage: issuer(empty seen, issued) | warp0(full seen, loads done, min done) | warp15(...)   [clk relative]\n");
  for (int i = 0; i < 32; ++i) { printf("%3d:", 200 + i); for (int k = 0; k < 8; ++k) printf(" %7lld", tr[i * 8 + k] ? tr[i * 8 + k] - t0 : -1); printf("\n"); }
#endif
  return 0;
}
