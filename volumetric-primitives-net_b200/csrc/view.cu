// Camera-frame changes of the reference, one pass over the points instead of 3-4 rotate_points calls:
//   view_to_obj_points  modules/transform/transform.py:21-47
//   obj_to_view_points  modules/transform/transform.py:50-73
// A setup kernel (one thread per sample) builds the chain of rotation matrices exactly the way the
// reference builds its quaternions; the apply kernel runs the chain per point (each rotation rounded
// separately, as the reference's successive bmm calls do) and scales by dist (or divides).
#include "common.cuh"

namespace vpn {

struct ViewChain {
  float m[3][9];
  float dist;
  int nrot;
};

__device__ __forceinline__ void quat_matrix(float ax, float ay, float az, float turns, float* m) {
  float q[4] = {ax, ay, az, turns};
  Pose p; make_pose(q, p);
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = p.r[i];
}

// mode 0: view_to_obj (needs angles), mode 1: obj_to_view
__global__ void view_setup_kernel(const float* __restrict__ dists, const float* __restrict__ elevs,
                                  const float* __restrict__ azims, const float* __restrict__ angles,
                                  ViewChain* __restrict__ chain, int B, int mode) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  ViewChain c;
  float e = __fdiv_rn(elevs[b], 360.0f), a = __fdiv_rn(azims[b], 360.0f);
  c.dist = dists[b];
  float me[9];
  quat_matrix(0.f, 0.f, -1.f, e, me);                 // rotation about -z by elev
  float yx = me[1], yy = me[4], yz = me[7];           // R * (0,1,0)
  if (mode == 0) {
    float ang = __fdiv_rn(-angles[b], 360.0f);
    quat_matrix(1.f, 0.f, 0.f, ang, c.m[0]);          // rotate_points_forward_x_axis(points, -angles)
    quat_matrix(yx, yy, yz, -a, c.m[1]);              // q2 = (y', -azim)
    quat_matrix(0.f, 0.f, -1.f, -e, c.m[2]);          // q3 = (-z, -elev)
    c.nrot = 3;
  } else {
    for (int i = 0; i < 9; ++i) c.m[0][i] = me[i];    // q = (-z, elev)
    quat_matrix(yx, yy, yz, a, c.m[1]);               // q = (y', azim)
    c.nrot = 2;
  }
  chain[b] = c;
}

__device__ __forceinline__ void mat_apply(const float* m, float* p) {
  float x = fmaf(m[2], p[2], fmaf(m[1], p[1], m[0] * p[0]));
  float y = fmaf(m[5], p[2], fmaf(m[4], p[1], m[3] * p[0]));
  float z = fmaf(m[8], p[2], fmaf(m[7], p[1], m[6] * p[0]));
  p[0] = x; p[1] = y; p[2] = z;
}
__device__ __forceinline__ void mat_apply_t(const float* m, float* p) {
  float x = fmaf(m[6], p[2], fmaf(m[3], p[1], m[0] * p[0]));
  float y = fmaf(m[7], p[2], fmaf(m[4], p[1], m[1] * p[0]));
  float z = fmaf(m[8], p[2], fmaf(m[5], p[1], m[2] * p[0]));
  p[0] = x; p[1] = y; p[2] = z;
}

// forward (transpose = 0): out = scale(M_k ... M_1 p); backward (transpose = 1): g_p = M_1^T ... M_k^T scale(g)
__global__ void view_apply_kernel(const float* __restrict__ in, float* __restrict__ out,
                                  const ViewChain* __restrict__ chain, int n, int mode, int transpose) {
  __shared__ ViewChain c;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) c = chain[b];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* src = in + 3 * ((size_t)b * n + i);
  float p[3] = {src[0], src[1], src[2]};
  if (!transpose) {
    for (int k = 0; k < c.nrot; ++k) mat_apply(c.m[k], p);
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = mode == 0 ? __fmul_rn(p[a], c.dist) : __fdiv_rn(p[a], c.dist);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = mode == 0 ? __fmul_rn(p[a], c.dist) : __fdiv_rn(p[a], c.dist);
    for (int k = c.nrot - 1; k >= 0; --k) mat_apply_t(c.m[k], p);
  }
  float* dst = out + 3 * ((size_t)b * n + i);
  dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
}

}  // namespace vpn

using namespace vpn;

extern "C" int vpn_view_workspace_bytes(int B, size_t* bytes) {
  if (B < 0 || !bytes) { vpn_set_error("view workspace: bad arguments"); return VPN_ERR_ARG; }
  *bytes = (size_t)B * sizeof(ViewChain);
  return VPN_OK;
}

// mode 0 = view_to_obj_points (angles required), 1 = obj_to_view_points (angles ignored).
// transpose 0 = forward; 1 = gradient w.r.t. the points given the upstream gradient in `in`.
extern "C" int vpn_view_points(int mode, int transpose, const float* in, const float* dists, const float* elevs,
                               const float* azims, const float* angles, float* out, void* workspace,
                               size_t workspace_bytes, int B, int n, void* stream) {
  if (B < 0 || n < 0 || mode < 0 || mode > 1) { vpn_set_error("view: bad shape/mode"); return VPN_ERR_SHAPE; }
  if (B == 0 || n == 0) return VPN_OK;
  if (B > 65535) { vpn_set_error("view: batch > 65535 unsupported"); return VPN_ERR_SHAPE; }
  if (!in || !out || !dists || !elevs || !azims || (mode == 0 && !angles) || !workspace) {
    vpn_set_error("view: null pointer"); return VPN_ERR_ARG;
  }
  if (workspace_bytes < (size_t)B * sizeof(ViewChain)) { vpn_set_error("view: workspace too small"); return VPN_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  ViewChain* chain = reinterpret_cast<ViewChain*>(workspace);
  view_setup_kernel<<<(B + 127) / 128, 128, 0, s>>>(dists, elevs, azims, angles, chain, B, mode);
  int rc = vpn_check_launch("view_setup_kernel");
  if (rc) return rc;
  dim3 grid((n + 255) / 256, B);
  view_apply_kernel<<<grid, 256, 0, s>>>(in, out, chain, n, mode, transpose);
  return vpn_check_launch("view_apply_kernel");
}
