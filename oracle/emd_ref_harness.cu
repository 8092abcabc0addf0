// TEST INFRASTRUCTURE - not product code.  Raw-pointer host harness around the REFERENCE's own EMD auction kernels.
//
// oracle/build_ref_emd.sh extracts the device code of /root/reference/modules/loss/emd/emd_cuda.cu (:9-226 - atomicMax,
// clear, calc_unass_*, Bid, GetMax, Assign, CalcDist - and :284-300 NmDistanceGradKernel) into a temporary header at
// build time and compiles this file against it into oracle/_ref/libemd_ref.so.  No reference source is stored in the
// repository; only the .so travels to the GPU box.  This file replaces the ATen host wrappers emd_cuda_forward /
// emd_cuda_backward (emd_cuda.cu:228-316) with the same launch sequence over plain device pointers; buffer
// initialisation follows emd_module.py:41-54 (assignment / assignment_inv = -1, everything else 0).
// Only tests/ may load the library (the -m gpu test that pins vpn_emd_fwd to the reference statistically).
#include <cuda_runtime.h>
#include <stdio.h>
#include EMD_REF_KERNELS

// All scratch lives in one caller-allocated device buffer of ref_emd_workspace_bytes(B, n) bytes.
struct RefWs { float *price, *bid_inc, *max_inc; int *ass_inv, *bid, *unass_idx, *max_idx, *unass_cnt, *unass_cnt_sum, *cnt_tmp; };
static RefWs carve(void* ws, int B, int n) {
  char* p = (char*)ws; RefWs w; size_t bn = (size_t)B * n * 4;
  w.price = (float*)p; p += bn; w.bid_inc = (float*)p; p += bn; w.max_inc = (float*)p; p += bn;
  w.ass_inv = (int*)p; p += bn; w.bid = (int*)p; p += bn; w.unass_idx = (int*)p; p += bn; w.max_idx = (int*)p; p += bn;
  w.unass_cnt = (int*)p; p += 512 * 4; w.unass_cnt_sum = (int*)p; p += 512 * 4; w.cnt_tmp = (int*)p;
  return w;
}
extern "C" size_t ref_emd_workspace_bytes(int B, int n) { return (size_t)B * n * 4 * 7 + 3 * 512 * 4; }

// returns 1 ok / 0 CUDA error / -1 bad shape, like emd_cuda_forward (emd_cuda.cu:236-249,276-281)
extern "C" int ref_emd_forward(float* xyz1, float* xyz2, float* dist, int* assignment, void* workspace, int B, int n,
                               float eps, int iters) {
  if (B > 512 || n % 1024 != 0) return -1;
  RefWs w = carve(workspace, B, n);
  cudaMemset(workspace, 0, ref_emd_workspace_bytes(B, n));
  cudaMemset(w.ass_inv, 0xff, (size_t)B * n * 4);
  cudaMemset(assignment, 0xff, (size_t)B * n * 4);
  cudaMemset(dist, 0, (size_t)B * n * 4);
  dim3 grid(B, n / 1024, 1);
  for (int i = 0; i < iters; i++) {
    clear<<<1, B>>>(B, w.cnt_tmp, w.unass_cnt);
    calc_unass_cnt<<<grid, 1024>>>(B, n, assignment, w.unass_cnt);
    calc_unass_cnt_sum<<<1, B>>>(B, w.unass_cnt, w.unass_cnt_sum);
    calc_unass_idx<<<grid, 1024>>>(B, n, assignment, w.unass_idx, w.unass_cnt, w.unass_cnt_sum, w.cnt_tmp);
    Bid<<<grid, 1024>>>(B, n, xyz1, xyz2, eps, assignment, w.ass_inv, w.price, w.bid, w.bid_inc, w.max_inc, w.unass_cnt,
                        w.unass_cnt_sum, w.unass_idx);
    GetMax<<<grid, 1024>>>(B, n, assignment, w.bid, w.bid_inc, w.max_inc, w.max_idx);
    Assign<<<grid, 1024>>>(B, n, assignment, w.ass_inv, w.price, w.bid, w.bid_inc, w.max_inc, w.max_idx, i == iters - 1);
  }
  CalcDist<<<grid, 1024>>>(B, n, xyz1, xyz2, dist, assignment);
  return cudaDeviceSynchronize() == cudaSuccess ? 1 : 0;
}

// grad_xyz1 must be zero-filled by the caller (emd_module.py:65: torch.zeros), the kernel accumulates with atomicAdd
extern "C" int ref_emd_backward(const float* xyz1, const float* xyz2, float* grad_xyz1, const float* grad_dist,
                                const int* assignment, int B, int n) {
  NmDistanceGradKernel<<<dim3(B, n / 1024, 1), 1024>>>(B, n, xyz1, xyz2, grad_dist, assignment, grad_xyz1);
  return cudaDeviceSynchronize() == cudaSuccess ? 1 : 0;
}
