// Gradient all-reduce over NVSwitch multicast (NVLS), the one collective of the data-parallel step (SURVEY.md
// section 8e: SUM of the network-parameter gradients, 22 875 848 fp32 for VPNetOneRes).
//
// The buffer is symmetric memory (same offset on every rank) with a multicast mapping.  Rank r owns the r-th
// slice: for each 16-byte element of its slice it issues ONE multimem.ld_reduce (the switch reads the element from
// every GPU and returns the sum) and ONE multimem.st (the switch writes the sum back to every GPU).  Per GPU that
// is S/N bytes in + S/N bytes out on its own links plus (N-1)/N * S delivered by the switch - half the traffic of
// a ring or two-shot all-reduce - in a single full-grid launch.  The caller brackets the launch with cross-rank
// barriers (inputs complete before, outputs visible after); the reference has no collective to compare against.
#include "common.cuh"

namespace vpn {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mc: multicast address of the buffer; nvec: float4 elements in the buffer; this rank reduces [lo, hi)
__global__ void __launch_bounds__(512)
allreduce_nvls_kernel(float4* __restrict__ mc, size_t lo, size_t hi) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent elements in flight per thread
  for (; i + 3 * stride < hi; i += 4 * stride) {
    const float4 a = multimem_ld_reduce_add(mc + i), b = multimem_ld_reduce_add(mc + i + stride);
    const float4 c = multimem_ld_reduce_add(mc + i + 2 * stride), d = multimem_ld_reduce_add(mc + i + 3 * stride);
    multimem_st(mc + i, a); multimem_st(mc + i + stride, b); multimem_st(mc + i + 2 * stride, c); multimem_st(mc + i + 3 * stride, d);
  }
  for (; i < hi; i += stride) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
  // one system-scope fence per CTA (cumulative over the CTA's stores after the barrier); a fence per thread costs
  // ~0.2 ms here
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

}  // namespace vpn

// multicast_ptr: multicast virtual address of a symmetric fp32 buffer of `numel` elements (numel % 4 == 0, 16-byte
// aligned); rank / world: this process.  In-place SUM over the ranks.  The caller must place a cross-rank barrier on
// the stream before (every rank's input is complete) and after (every rank's writes have landed) this call.
extern "C" int vpn_allreduce_nvls(void* multicast_ptr, size_t numel, int rank, int world, void* stream) {
  if (!multicast_ptr || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) || numel % 4 != 0 || world < 1 || rank < 0 || rank >= world) {
    vpn_set_error("allreduce nvls: bad arguments"); return VPN_ERR_ARG;
  }
  const size_t nvec = numel / 4;
  const size_t per = (nvec + world - 1) / world;
  const size_t lo = (size_t)rank * per < nvec ? (size_t)rank * per : nvec;
  const size_t hi = lo + per < nvec ? lo + per : nvec;
  if (hi <= lo) return VPN_OK;
  const int threads = 512;
  size_t blocks = (hi - lo + (size_t)threads * 4 - 1) / ((size_t)threads * 4);
  const size_t cap = (size_t)vpn::device_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  vpn::allreduce_nvls_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(multicast_ptr), lo, hi);
  return vpn_check_launch("allreduce_nvls_kernel");
}
