// Area-weighted surface sampling of a batch of triangle meshes with shared topology.
//
// Replaces the per-sample Python loop over kaolin's TriangleMesh.sample at train_sphere.py:71-80 (and
// dataset/dataset.py:162-165): face areas -> discrete distribution -> face choice -> sqrt-u barycentric point
//     p = (1 - sqrt(u1)) v0 + sqrt(u1) (1 - u2) v1 + sqrt(u1) u2 v2.
// kaolin draws the face with torch.distributions.Categorical and the barycentric pair with two Uniform samples;
// here all three draws are explicit inputs u (B, n, 3) and the face is the inverse CDF of u0 (first face whose
// cumulative area share exceeds u0), so that the oracle can replay the same draws.  One launch builds every
// sample's CDF (one CTA per mesh, shared-memory scan), one launch places the points.
#include "common.cuh"

namespace vpn {

constexpr int kMsThreads = 256;

// 2 * area of face f of mesh `verts` (kaolin: sqrt(a + b + c) / 2 with a, b, c the squared cross-product components
// of the edge vectors (v0 - v1) and (v1 - v2))
__device__ __forceinline__ float face_area(const float* __restrict__ verts, const int* __restrict__ faces, int f) {
  const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
  // every operation rounded separately, as the chain of torch ops in kaolin does (no FMA contraction)
  const float x1 = __fsub_rn(verts[3 * i0], verts[3 * i1]), x2 = __fsub_rn(verts[3 * i0 + 1], verts[3 * i1 + 1]), x3 = __fsub_rn(verts[3 * i0 + 2], verts[3 * i1 + 2]);
  const float y1 = __fsub_rn(verts[3 * i1], verts[3 * i2]), y2 = __fsub_rn(verts[3 * i1 + 1], verts[3 * i2 + 1]), y3 = __fsub_rn(verts[3 * i1 + 2], verts[3 * i2 + 2]);
  const float a = __fsub_rn(__fmul_rn(x2, y3), __fmul_rn(x3, y2));
  const float bb = __fsub_rn(__fmul_rn(x3, y1), __fmul_rn(x1, y3));
  const float c = __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(x2, y1));
  return __fdiv_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(bb, bb)), __fmul_rn(c, c))), 2.0f);
}

// grid: x = mesh.  cdf (B, F): inclusive prefix sums of area / (sum + eps), in face order (sequential per chunk,
// fixed combination order: deterministic).
__global__ void __launch_bounds__(kMsThreads)
mesh_cdf_kernel(const float* __restrict__ verts, const int* __restrict__ faces, float* __restrict__ cdf,
                int V, int F, float eps) {
  __shared__ float part[kMsThreads];
  __shared__ float total_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* vb = verts + (size_t)b * V * 3;
  float* cb = cdf + (size_t)b * F;
  const int per = (F + kMsThreads - 1) / kMsThreads;
  const int f0 = tid * per, f1 = min(F, f0 + per);
  float run = 0.f;
  for (int f = f0; f < f1; ++f) { run = __fadd_rn(run, face_area(vb, faces, f)); cb[f] = run; }      // local inclusive sums
  part[tid] = run;
  __syncthreads();
  if (tid == 0) {
    float acc = 0.f;
    for (int i = 0; i < kMsThreads; ++i) { const float p = part[i]; part[i] = acc; acc = __fadd_rn(acc, p); }
    total_s = acc;
  }
  __syncthreads();
  const float base = part[tid], denom = __fadd_rn(total_s, eps);
  for (int f = f0; f < f1; ++f) cb[f] = __fdiv_rn(__fadd_rn(cb[f], base), denom);
}

// grid: x = blocks of points, y = mesh
__global__ void __launch_bounds__(kMsThreads)
mesh_sample_kernel(const float* __restrict__ verts, const int* __restrict__ faces, const float* __restrict__ cdf,
                   const float* __restrict__ u, float* __restrict__ points, int* __restrict__ face_idx,
                   int V, int F, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kMsThreads + threadIdx.x;
  if (i >= n) return;
  const float* ub = u + ((size_t)b * n + i) * 3;
  const float* cb = cdf + (size_t)b * F;
  const float u0 = ub[0];
  int lo = 0, hi = F - 1;                          // first face with cdf > u0 (last face if rounding leaves none)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cb[mid] > u0) hi = mid; else lo = mid + 1;
  }
  const float* vb = verts + (size_t)b * V * 3;
  const int i0 = faces[3 * lo], i1 = faces[3 * lo + 1], i2 = faces[3 * lo + 2];
  const float su = __fsqrt_rn(ub[1]), vv = ub[2];
  const float w0 = __fsub_rn(1.0f, su), w1 = __fmul_rn(su, __fsub_rn(1.0f, vv)), w2 = __fmul_rn(su, vv);
  float* p = points + ((size_t)b * n + i) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    p[c] = __fadd_rn(__fadd_rn(__fmul_rn(w0, vb[3 * i0 + c]), __fmul_rn(w1, vb[3 * i1 + c])), __fmul_rn(w2, vb[3 * i2 + c]));
  face_idx[(size_t)b * n + i] = lo;
}

// grad_verts (B, V, 3) must be zeroed by the caller; the face choice is piecewise constant (no gradient), the
// barycentric weights are constants.
__global__ void __launch_bounds__(kMsThreads)
mesh_sample_bwd_kernel(const int* __restrict__ faces, const float* __restrict__ u, const int* __restrict__ face_idx,
                       const float* __restrict__ grad_points, float* __restrict__ grad_verts, int V, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kMsThreads + threadIdx.x;
  if (i >= n) return;
  const float* ub = u + ((size_t)b * n + i) * 3;
  const int f = face_idx[(size_t)b * n + i];
  const float su = __fsqrt_rn(ub[1]), vv = ub[2];
  const float w[3] = {__fsub_rn(1.0f, su), __fmul_rn(su, __fsub_rn(1.0f, vv)), __fmul_rn(su, vv)};
  const float* g = grad_points + ((size_t)b * n + i) * 3;
  float* gv = grad_verts + (size_t)b * V * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int vi = faces[3 * f + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&gv[3 * vi + c], w[k] * g[c]);
  }
}

}  // namespace vpn

extern "C" int vpn_mesh_sample_fwd(const float* verts, const int* faces, const float* u, float* points, int* face_idx,
                                   float* cdf, int B, int V, int F, int n, void* stream) {
  if (B < 0 || V <= 0 || F <= 0 || n <= 0) { vpn_set_error("mesh sample fwd: bad shape B=%d V=%d F=%d n=%d", B, V, F, n); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!verts || !faces || !u || !points || !face_idx || !cdf) { vpn_set_error("mesh sample fwd: null pointer"); return VPN_ERR_ARG; }
  if (B > 65535) { vpn_set_error("mesh sample fwd: batch > 65535 unsupported"); return VPN_ERR_SHAPE; }
  cudaStream_t s = (cudaStream_t)stream;
  vpn::mesh_cdf_kernel<<<B, vpn::kMsThreads, 0, s>>>(verts, faces, cdf, V, F, 1e-10f);
  int rc = vpn_check_launch("mesh_cdf_kernel");
  if (rc) return rc;
  vpn::mesh_sample_kernel<<<dim3((n + vpn::kMsThreads - 1) / vpn::kMsThreads, B), vpn::kMsThreads, 0, s>>>(verts, faces, cdf, u, points, face_idx, V, F, n);
  return vpn_check_launch("mesh_sample_kernel");
}

extern "C" int vpn_mesh_sample_bwd(const int* faces, const float* u, const int* face_idx, const float* grad_points,
                                   float* grad_verts, int B, int V, int F, int n, void* stream) {
  if (B < 0 || V <= 0 || F <= 0 || n <= 0) { vpn_set_error("mesh sample bwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!faces || !u || !face_idx || !grad_points || !grad_verts) { vpn_set_error("mesh sample bwd: null pointer"); return VPN_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(grad_verts, 0, (size_t)B * V * 3 * sizeof(float), s) != cudaSuccess) { vpn_set_error("mesh sample bwd: memset failed"); return VPN_ERR_CUDA; }
  vpn::mesh_sample_bwd_kernel<<<dim3((n + vpn::kMsThreads - 1) / vpn::kMsThreads, B), vpn::kMsThreads, 0, s>>>(faces, u, face_idx, grad_points, grad_verts, V, n);
  return vpn_check_launch("mesh_sample_bwd_kernel");
}
