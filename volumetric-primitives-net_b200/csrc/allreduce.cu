// Gradient all-reduce over NVSwitch multicast (NVLS), the one collective of the data-parallel step (SURVEY.md
// section 8e: SUM of the network-parameter gradients, 22 875 848 fp32 for VPNetOneRes).
//
// The buffer is symmetric memory (same offset on every rank) with a multicast mapping.  Rank r owns the r-th
// slice: for each 16-byte element of its slice it issues ONE multimem.ld_reduce (the switch reads the element from
// every GPU and returns the sum) and ONE multimem.st (the switch writes the sum back to every GPU).  Per GPU that
// is S/N bytes in + S/N bytes out on its own links plus (N-1)/N * S delivered by the switch - half the traffic of
// a ring or two-shot all-reduce - in a single full-grid launch.
//
// Two entry points:
//   vpn_allreduce_nvls       the bare kernel; the caller brackets it with cross-rank barriers on the stream;
//   vpn_allreduce_nvls_sync  ONE self-synchronising launch: the cross-rank barriers are inside the kernel (arrive =
//                            one multimem.red.add on a counter that lives in every GPU's copy of the buffer tail, wait =
//                            spin on the local copy), the epoch is kept in device memory, so the launch takes no
//                            per-call host state and can be captured in a CUDA graph and replayed.
// The reference is single process and has no collective to compare against.
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
// counter += v in EVERY GPU's copy, release at system scope (everything this thread observed / fenced happens before)
__device__ __forceinline__ void multimem_red_add_release(unsigned* mc, unsigned v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" :: "l"(mc), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait (a peer that never arrives must not hang the GPU): after kSpinTimeoutNs the error word is set, this and
// every later launch skip their waits, and the host reports the failure (vpn_allreduce_nvls_sync_status).
constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned target, unsigned* err) {
  if (*reinterpret_cast<volatile unsigned*>(err)) return;
  unsigned long long t0 = 0;
  for (unsigned it = 1; (int)(ld_acquire_sys(p) - target) < 0; ++it) {
    __nanosleep(20);
    if ((it & 1023u) == 0) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSpinTimeoutNs) { atomicExch(err, 1u); return; }
    }
  }
}

// this rank reduces float4 elements [lo, hi) of the multicast buffer mc
__device__ __forceinline__ void reduce_slice(float4* __restrict__ mc, size_t lo, size_t hi) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent elements in flight per thread
  for (; i + 3 * stride < hi; i += 4 * stride) {
    const float4 a = multimem_ld_reduce_add(mc + i), b = multimem_ld_reduce_add(mc + i + stride);
    const float4 c = multimem_ld_reduce_add(mc + i + 2 * stride), d = multimem_ld_reduce_add(mc + i + 3 * stride);
    multimem_st(mc + i, a); multimem_st(mc + i + stride, b); multimem_st(mc + i + 2 * stride, c); multimem_st(mc + i + 3 * stride, d);
  }
  for (; i < hi; i += stride) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
}

// Pipelined form of the same slice reduction: every thread walks its elements with the store of batch i issued AFTER the
// load of batch i + 1, so that the switch carries reduced data towards this GPU and broadcast data away from it at the
// same time.  (In reduce_slice a thread issues all its loads and then all its stores; with one batch per thread the
// whole GPU first only reads - link egress busy serving the peers' reads, ingress nearly idle - and then only writes.)
template <int U>
__device__ __forceinline__ void reduce_slice_pipelined(float4* __restrict__ mc, size_t lo, size_t hi) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 cur[U]; size_t at[U];
  int n = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) { at[u] = i + (size_t)u * stride; if (at[u] < hi) { cur[u] = multimem_ld_reduce_add(mc + at[u]); n = u + 1; } }
  i += (size_t)U * stride;
  while (n > 0) {
    float4 nxt[U]; size_t nat[U];
    int nn = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) { nat[u] = i + (size_t)u * stride; if (nat[u] < hi) { nxt[u] = multimem_ld_reduce_add(mc + nat[u]); nn = u + 1; } }
#pragma unroll
    for (int u = 0; u < U; ++u) if (u < n) multimem_st(mc + at[u], cur[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) { cur[u] = nxt[u]; at[u] = nat[u]; }
    n = nn;
    i += (size_t)U * stride;
  }
}

__global__ void __launch_bounds__(512)
allreduce_nvls_kernel(float4* __restrict__ mc, size_t lo, size_t hi) {
  reduce_slice(mc, lo, hi);
  // one system-scope fence per CTA (cumulative over the CTA's stores after the barrier); a fence per thread costs
  // ~0.2 ms here
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

// Flag words in the buffer tail (u32 index; each on its own 128-byte line).  kArrive / kDone are written through the
// multicast address (every GPU's copy counts every rank) and read locally; kEpoch / kCtas are private to the GPU.
enum { kArrive = 0, kDone = 32, kEpoch = 64, kCtas = 96, kError = 97, kFlagWords = 128 };

// mc / local: multicast and local (unicast) address of the same symmetric buffer; flags start at float index `flag_off`.
// Start: every rank's input is complete (stream order on each rank, then arrive + wait).  End: the LAST CTA of the
// rank - after every CTA has fenced its stores at system scope - announces the rank done and waits for all ranks, so
// that the kernel's completion on a GPU means every rank's sums have landed in that GPU's copy.
// V: 0 = all loads then all stores per thread, 1 / 2 / 4 = pipelined with that many elements per batch
template <int V>
__global__ void __launch_bounds__(512, (V == 4 ? 2 : 4))
allreduce_nvls_sync_kernel(float4* __restrict__ mc, float* __restrict__ local, size_t flag_off, size_t lo, size_t hi, int world) {
  unsigned* fl_local = reinterpret_cast<unsigned*>(local + flag_off);
  unsigned* fl_mc = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(mc) + flag_off);
  __shared__ unsigned s_target;
  if (threadIdx.x == 0) {
    // kEpoch is only advanced by this GPU's last CTA of the previous launch (stream ordered): plain read
    const unsigned target = (*reinterpret_cast<volatile unsigned*>(fl_local + kEpoch) + 1u) * (unsigned)world;
    if (blockIdx.x == 0) multimem_red_add_release(fl_mc + kArrive, 1u);
    spin_until(fl_local + kArrive, target, fl_local + kError);
    s_target = target;
  }
  __syncthreads();
  if (V == 0) reduce_slice(mc, lo, hi); else reduce_slice_pipelined<(V == 0 ? 1 : V)>(mc, lo, hi);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                     // this CTA's multimem stores are ordered before the count
    const unsigned prev = atomicAdd(fl_local + kCtas, 1u);
    if (prev == gridDim.x - 1) {                                // last CTA of this rank
      fl_local[kCtas] = 0;
      __threadfence_system();                                   // cumulativity: the other CTAs' fenced stores come first
      multimem_red_add_release(fl_mc + kDone, 1u);
      spin_until(fl_local + kDone, s_target, fl_local + kError);
      fl_local[kEpoch] = s_target / (unsigned)world;
      __threadfence();
    }
  }
}

static int slice_of(size_t numel, int rank, int world, size_t* lo, size_t* hi) {
  const size_t nvec = numel / 4;
  const size_t per = (nvec + world - 1) / world;
  *lo = (size_t)rank * per < nvec ? (size_t)rank * per : nvec;
  *hi = *lo + per < nvec ? *lo + per : nvec;
  return 0;
}

static unsigned grid_for(size_t nvec_slice, int threads, int ctas_per_sm = 4) {
  size_t blocks = (nvec_slice + (size_t)threads * 4 - 1) / ((size_t)threads * 4);
  const size_t cap = (size_t)device_sm_count() * ctas_per_sm;  // <= 4 x 512 threads per SM: all CTAs co-resident
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace vpn

// multicast_ptr: multicast virtual address of a symmetric fp32 buffer of `numel` elements (numel % 4 == 0, 16-byte
// aligned); rank / world: this process.  In-place SUM over the ranks.  The caller must place a cross-rank barrier on
// the stream before (every rank's input is complete) and after (every rank's writes have landed) this call.
extern "C" int vpn_allreduce_nvls(void* multicast_ptr, size_t numel, int rank, int world, void* stream) {
  if (!multicast_ptr || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) || numel % 4 != 0 || world < 1 || rank < 0 || rank >= world) {
    vpn_set_error("allreduce nvls: bad arguments"); return VPN_ERR_ARG;
  }
  size_t lo, hi;
  vpn::slice_of(numel, rank, world, &lo, &hi);
  if (hi <= lo) return VPN_OK;
  const int threads = 512;
  vpn::allreduce_nvls_kernel<<<vpn::grid_for(hi - lo, threads), threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(multicast_ptr), lo, hi);
  return vpn_check_launch("allreduce_nvls_kernel");
}

// Index (in floats, relative to the flag area) of the word that is non-zero after a cross-rank wait timed out.
extern "C" int vpn_allreduce_nvls_error_word(void) { return vpn::kError; }

extern "C" int vpn_allreduce_nvls_flag_floats(size_t* floats) {
  if (!floats) { vpn_set_error("allreduce nvls: null pointer"); return VPN_ERR_ARG; }
  *floats = vpn::kFlagWords;
  return VPN_OK;
}

// Self-synchronising variant.  The symmetric buffer holds numel payload floats (numel % 4 == 0) followed by
// vpn_allreduce_nvls_flag_floats() flag words, which must be ZERO on every rank before the first call (and all ranks
// must have observed that, e.g. one host-side barrier at set-up); multicast_ptr / local_ptr address the same buffer.
// Every rank must issue the same sequence of calls.  No host state: capturable in a CUDA graph.
extern "C" int vpn_allreduce_nvls_sync(void* multicast_ptr, void* local_ptr, size_t numel, int rank, int world, void* stream) {
  if (!multicast_ptr || !local_ptr || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) || numel % 4 != 0 || world < 1 || rank < 0 ||
      rank >= world) {
    vpn_set_error("allreduce nvls sync: bad arguments"); return VPN_ERR_ARG;
  }
  size_t lo, hi;
  vpn::slice_of(numel, rank, world, &lo, &hi);
  // Launch shape.  Measured on 8 B200s, 91.5 MB (tools/allreduce_sweep.py -> profiles/r02u_allreduce_sweep_8gpu.txt): the
  // switch is saturated by ~19 000 threads with one load and one store in flight each; more threads only add contention
  // (4 CTAs x 512 threads per SM, loads-then-stores: 0.267 ms; 1 CTA x 512: 0.237; pipelined: 0.222; pipelined, 256
  // threads on every second SM: 0.2095 ms = 765 GB/s bus bandwidth; NCCL 2.28.9: 0.355 ms).  Probe knobs
  // (vpn_set_tuning): "ar_variant" 0 = loads then stores, 1 / 2 / 4 = pipelined with that batch (default 1, passed as 8
  // for variant 0); "ar_threads" 128 | 256 | 512; "ar_ctas" CTAs per SM; "ar_grid_div" use 1 / div of the SMs.
  int threads = vpn::tuning_value(vpn::kTuneArThreads);
  if (threads != 128 && threads != 512) threads = 256;
  int variant = vpn::tuning_value(vpn::kTuneArVariant);
  if (variant == 0) variant = 1; else if (variant == 8) variant = 0;
  int cps = vpn::tuning_value(vpn::kTuneArCtas);
  if (cps < 1 || cps > 4) cps = 1;
  if (variant == 4 && cps > 2) cps = 2;                        // that variant is built for two resident CTAs per SM
  unsigned grid = vpn::grid_for(hi > lo ? hi - lo : 0, threads, cps);
  int div = vpn::tuning_value(vpn::kTuneArGridDiv);
  if (div < 1) div = 2;
  if (cps == 1) grid = (grid + div - 1) / div;
  float4* mc = reinterpret_cast<float4*>(multicast_ptr); float* lp = reinterpret_cast<float*>(local_ptr);
  cudaStream_t s = (cudaStream_t)stream;
  switch (variant) {
    case 1:  vpn::allreduce_nvls_sync_kernel<1><<<grid, threads, 0, s>>>(mc, lp, numel, lo, hi, world); break;
    case 2:  vpn::allreduce_nvls_sync_kernel<2><<<grid, threads, 0, s>>>(mc, lp, numel, lo, hi, world); break;
    case 4:  vpn::allreduce_nvls_sync_kernel<4><<<grid, threads, 0, s>>>(mc, lp, numel, lo, hi, world); break;
    default: vpn::allreduce_nvls_sync_kernel<0><<<grid, threads, 0, s>>>(mc, lp, numel, lo, hi, world); break;
  }
  return vpn_check_launch("allreduce_nvls_sync_kernel");
}
