// Cycle accounting of chamfer_tc_kernel (CTA 0): build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DVPN_TC_PROF -I../volumetric-primitives-net_b200/csrc -o tc_prof tc_prof.cu
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
  float *p1, *p2, *rbest, *cbest; vpn::u64* rmask; unsigned* cmask; float2* tslack; int* fb;
  CK(cudaMalloc(&p1, h1.size() * 4)); CK(cudaMalloc(&p2, h2.size() * 4));
  CK(cudaMalloc(&rbest, (size_t)B * P * 4)); CK(cudaMalloc(&rmask, (size_t)B * P * 8));
  CK(cudaMalloc(&cbest, (size_t)B * ntiles * M * 4)); CK(cudaMalloc(&cmask, (size_t)B * ntiles * M * 4));
  CK(cudaMalloc(&tslack, (size_t)B * ntiles * 8)); CK(cudaMalloc(&fb, B * 4)); CK(cudaMemset(fb, 0, B * 4));
  CK(cudaMemcpy(p1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(p2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    long long zero[16] = {0}; CK(cudaMemcpyToSymbol(vpn::g_tc_prof, zero, sizeof(zero)));
    CK(cudaEventRecord(e0));
    if (vpn::chamfer_tc_launch(p1, p2, rbest, rmask, cbest, cmask, tslack, fb, B, P, M, NB, ntiles, nsplit, nchunks, cps, 0)) return 1;
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long pr[16]; CK(cudaMemcpyFromSymbol(pr, vpn::g_tc_prof, sizeof(pr)));
    const double nblk = (double)nchunks * NB;   // blocks; every block is visited in both phases
    printf("rep %d: %.3f ms, CTA0 total %lld clk = %.1f clk/block; setup %lld\n", rep, ms, pr[10], pr[10] / nblk, pr[9]);
    printf("  mma : wait colfull %.1f  wait empty %.1f  issue %.1f   (clk per block)\n", pr[0] / nblk, pr[1] / nblk, pr[2] / nblk);
    printf("  row : wait full %.1f  ld %.1f  min+fold %.1f\n", pr[3] / nblk, pr[4] / nblk, pr[5] / nblk);
    printf("  col : wait full %.1f  ld %.1f  min+fold %.1f\n", pr[6] / nblk, pr[7] / nblk, pr[8] / nblk);
  }
  return 0;
}
