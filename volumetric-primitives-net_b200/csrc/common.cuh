// Shared device helpers for the vpn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#define VPN_OK 0
#define VPN_ERR_CUDA (-1)
#define VPN_ERR_SHAPE (-2)
#define VPN_ERR_ARG (-3)
#define VPN_ERR_WORKSPACE (-4)

// rotate.py:4 / sphere.py:7 of the reference hard-code this fp32-rounded pi.
#define VPN_PI 3.1415927410125732f

void vpn_set_error(const char* fmt, ...);
int vpn_check_launch(const char* what);

namespace vpn {

typedef unsigned long long u64;

// Per-device "done once" flag.  cudaFuncSetAttribute is a property of (kernel, device), so a process that drives several
// GPUs must opt every one of them in to > 48 KB of dynamic shared memory; bit d = done on device ordinal d (ordinals
// >= 64 are simply set again on every launch).
struct DeviceOnce { std::atomic<unsigned long long> mask{0}; };
template <typename K>
inline cudaError_t set_dyn_smem(K kernel, int bytes, DeviceOnce& once) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
  if (bit && (once.mask.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && bit) once.mask.fetch_or(bit, std::memory_order_release);
  return e;
}
// SM count of the calling thread's current device (cached per ordinal; api.cu)
int device_sm_count();
// Tuning overrides set through vpn_set_tuning (tests and probes only; 0 = automatic choice)
int tuning_value(int key);
enum { kTuneTiledR = 0, kTuneTcNb = 1, kTuneEmdCluster = 2, kTuneTcPrune = 3, kTuneSerialRecovery = 4, kTuneArVariant = 5, kTuneArCtas = 6, kTuneArThreads = 7, kTuneArGridDiv = 8, kTunePrepNearCols = 9, kTunePrepRepsCols = 10, kTunePrepNearRows = 11, kTunePrepRepsRows = 12, kTunePrepProbe = 13, kTunePrepDeterministic = 14, kTuneTcHunits = 15, kTuneCount = 16 };

// Rigid pose of one primitive: row-major rotation R (from the reference's axis/turn-fraction
// "quaternion"), plus what the backward pass needs to chain dL/dR to dL/dq.
struct Pose {
  float r[9];
  float qn[4];     // normalised quaternion (x, y, z, w)
  float s, c;      // sin / cos of the half angle
  float len;       // norm of the un-normalised quaternion
};

// torch's `%` for floats (remainder with the sign of the divisor), divisor 1.
__device__ __forceinline__ float mod1(float w) {
  float m = fmodf(w, 1.0f);
  if (m != 0.0f && m < 0.0f) m += 1.0f;
  return m;
}

// modules/transform/rotate.py:59-72 (refine_quaternions) followed by :28-46 (get_rotation_matrices).
__device__ __forceinline__ void make_pose(const float* __restrict__ q, Pose& p) {
  float a = q[0], b = q[1], cc = q[2], w = q[3];
  // angles = ((w % 1) * 2 * PI) / 2
  float half = __fdiv_rn(__fmul_rn(__fmul_rn(mod1(w), 2.0f), VPN_PI), 2.0f);
  float s = sinf(half), c = cosf(half);
  float ux = a * s, uy = b * s, uz = cc * s, uw = c;
  float len = sqrtf(ux * ux + uy * uy + uz * uz + uw * uw);
  float x = ux / len, y = uy / len, z = uz / len, ww = uw / len;
  p.s = s; p.c = c; p.len = len;
  p.qn[0] = x; p.qn[1] = y; p.qn[2] = z; p.qn[3] = ww;
  float x2 = x * x, y2 = y * y, z2 = z * z, w2 = ww * ww;
  float xy = x * y, zw = z * ww, xz = x * z, yw = y * ww, yz = y * z, xw = x * ww;
  p.r[0] = x2 - y2 - z2 + w2;  p.r[1] = 2.f * (xy - zw);       p.r[2] = 2.f * (xz + yw);
  p.r[3] = 2.f * (xy + zw);    p.r[4] = -x2 + y2 - z2 + w2;    p.r[5] = 2.f * (yz - xw);
  p.r[6] = 2.f * (xz - yw);    p.r[7] = 2.f * (yz + xw);       p.r[8] = -x2 - y2 + z2 + w2;
}

// Chain dL/dR (G, row-major 3x3) back to dL/dq (raw axis a,b,c and turn fraction w).
__device__ __forceinline__ void pose_backward(const float* __restrict__ q, const Pose& p,
                                              const float* __restrict__ G, float* __restrict__ gq) {
  float x = p.qn[0], y = p.qn[1], z = p.qn[2], w = p.qn[3];
  float gx = 2.f * (x * G[0] + y * G[3] + z * G[6] + y * G[1] - x * G[4] + w * G[7] + z * G[2] - w * G[5] - x * G[8]);
  float gy = 2.f * (-y * G[0] + x * G[3] - w * G[6] + x * G[1] + y * G[4] + z * G[7] + w * G[2] + z * G[5] - y * G[8]);
  float gz = 2.f * (-z * G[0] + w * G[3] + x * G[6] - w * G[1] - z * G[4] + y * G[7] + x * G[2] + y * G[5] + z * G[8]);
  float gw = 2.f * (w * G[0] + z * G[3] - y * G[6] - z * G[1] + w * G[4] + x * G[7] + y * G[2] - x * G[5] + w * G[8]);
  // n = u / |u|
  float dot = x * gx + y * gy + z * gz + w * gw;
  float inv = 1.0f / p.len;
  float ux = (gx - x * dot) * inv, uy = (gy - y * dot) * inv, uz = (gz - z * dot) * inv, uw = (gw - w * dot) * inv;
  // u = (a s, b s, c s, cos)
  gq[0] = ux * p.s; gq[1] = uy * p.s; gq[2] = uz * p.s;
  float gth = (q[0] * ux + q[1] * uy + q[2] * uz) * p.c - uw * p.s;
  gq[3] = gth * VPN_PI;      // d(half angle)/dw = PI (the modulo has unit slope)
}

// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may be scheduled while its predecessor in
// the stream is still running; pdl_enter() first lets ITS successor do the same, then blocks until the predecessor has
// completed and its writes are visible.  It must come before the kernel's first global-memory access.  In a kernel
// launched the ordinary way both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 4- / 8-byte asynchronous global -> shared copies (cp.async.ca); commit / wait by the caller
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vpn
