// Earth mover's distance approximation by the auction algorithm (SURVEY.md section 8f "next" #1).
//
// Replaces modules/loss/emd (emd_module.py:29-79 -> emd.cpp:6-23 -> emd_cuda.cu:227-316): the reference runs
// 7 launches per iteration (351 for the training setting eps = 0.005, iters = 50, train.py:188-195), keeps all
// state in global memory and lets racing writes decide ties.  Here ONE launch runs the whole auction of a sample:
// one CTA per sample, the auction state (prices, both assignment maps, bids, the objects' coordinates) lives in
// shared memory for n <= 4096 (44 n bytes; larger clouds keep it in a caller-provided workspace), phases are
// separated by __syncthreads, and ties are resolved deterministically (lowest index wins), so the result is
// reproducible run to run.
//   per iteration   compact the unassigned bidders
//                   Bid    : one warp per bidder, lanes stride over the objects: value = 3 - |x2 - x1| - price,
//                            best (first maximum) / second best, warp merge; increment = best - better + eps;
//                            atomicMax of the object's largest increment          (emd_cuda.cu:95-179)
//                   GetMax : the lowest bidder within 1e-6 of the object's maximum wins   (:181-194)
//                   Assign : winner takes the object, evicts the previous owner, price += increment (:196-216);
//                            in the last iteration every unassigned bidder takes its object unconditionally
//   finally         dist = |x1 - x2[assignment]|^2                                  (:218-226)
// Arithmetic: fp32, every operation rounded separately (the reference mixes a double literal and nvcc's FMA
// contraction; its results are an approximation and not reproducible bit for bit in any case).
#include "common.cuh"

namespace vpn {

constexpr int kEmdThreads = 1024;
constexpr int kEmdSmemMaxN = 4096;
constexpr int kEmdArrays = 11;                 // 4-byte words of state per point

struct EmdState {
  float *x, *y, *z, *price, *inc;
  int *assign, *assign_inv, *bid, *max_inc, *winner, *un;
};
__device__ __forceinline__ EmdState emd_carve(unsigned char* p, int n) {
  EmdState s;
  float* f = reinterpret_cast<float*>(p);
  s.x = f; s.y = f + n; s.z = f + 2 * (size_t)n; s.price = f + 3 * (size_t)n; s.inc = f + 4 * (size_t)n;
  int* i = reinterpret_cast<int*>(f + 5 * (size_t)n);
  s.assign = i; s.assign_inv = i + n; s.bid = i + 2 * (size_t)n; s.max_inc = i + 3 * (size_t)n;
  s.winner = i + 4 * (size_t)n; s.un = i + 5 * (size_t)n;
  return s;
}

struct Top2 { float best, better; int i; };
// multiset top-2 merge; equal best values: the lower index keeps the object, the other becomes the runner-up
__device__ __forceinline__ Top2 top2_merge(const Top2& a, const Top2& b) {
  Top2 r;
  const bool a_wins = (a.best > b.best) || (a.best == b.best && a.i < b.i);
  if (a_wins) { r.best = a.best; r.i = a.i; r.better = fmaxf(a.better, b.best); }
  else        { r.best = b.best; r.i = b.i; r.better = fmaxf(b.better, a.best); }
  return r;
}

// grid: x = sample.  SMEM: state in dynamic shared memory, else in ws + sample * 44 n bytes.
template <bool SMEM>
__global__ void __launch_bounds__(kEmdThreads, 1)
emd_auction_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, float* __restrict__ dist,
                   int* __restrict__ assignment, unsigned char* __restrict__ ws, int n, float eps, int iters) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ int s_cnt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const EmdState st = emd_carve(SMEM ? s_dyn : ws + (size_t)b * kEmdArrays * 4 * n, n);
  const float* X1 = xyz1 + (size_t)b * n * 3;
  const float* X2 = xyz2 + (size_t)b * n * 3;
  const int neg_big = __float_as_int(-1e9f);
  for (int k = tid; k < n; k += kEmdThreads) {
    st.x[k] = X2[3 * k]; st.y[k] = X2[3 * k + 1]; st.z[k] = X2[3 * k + 2];
    st.price[k] = 0.f; st.assign[k] = -1; st.assign_inv[k] = -1; st.max_inc[k] = neg_big; st.winner[k] = 0x7fffffff;
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const bool last = (it == iters - 1);
    // ---- unassigned bidders (any order: the phases below do not depend on it)
    for (int j0 = 0; j0 < n; j0 += kEmdThreads) {
      const int j = j0 + tid;
      const bool un = j < n && st.assign[j] == -1;
      const unsigned bal = __ballot_sync(0xffffffffu, un);
      int base = 0;
      if (lane == 0 && bal) base = atomicAdd(&s_cnt, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (un) st.un[base + __popc(bal & ((1u << lane) - 1u))] = j;
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) break;                                     // everything assigned: further iterations change nothing
    // ---- Bid
    for (int u = warp; u < cnt; u += kEmdThreads / 32) {
      const int j = st.un[u];
      const float x1 = X1[3 * j], y1 = X1[3 * j + 1], z1 = X1[3 * j + 2];
      Top2 t; t.best = -1e9f; t.better = -1e9f; t.i = 0x7fffffff;
      for (int k = lane; k < n; k += 32) {
        const float dx = __fsub_rn(st.x[k], x1), dy = __fsub_rn(st.y[k], y1), dz = __fsub_rn(st.z[k], z1);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float v = __fsub_rn(__fsub_rn(3.0f, __fsqrt_rn(d2)), st.price[k]);
        if (v > t.best) { t.better = t.best; t.best = v; t.i = k; }
        else if (v > t.better) t.better = v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Top2 other;
        other.best = __shfl_xor_sync(0xffffffffu, t.best, o);
        other.better = __shfl_xor_sync(0xffffffffu, t.better, o);
        other.i = __shfl_xor_sync(0xffffffffu, t.i, o);
        t = top2_merge(t, other);
      }
      if (lane == 0) {
        const float inc = __fadd_rn(__fsub_rn(t.best, t.better), eps);
        st.bid[j] = t.i; st.inc[j] = inc;
        atomicMax(&st.max_inc[t.i], __float_as_int(inc));   // increments are >= eps >= 0: integer order = float order
      }
    }
    __syncthreads();
    // ---- GetMax: lowest bidder within 1e-6 of the object's largest increment
    for (int u = tid; u < cnt; u += kEmdThreads) {
      const int j = st.un[u], t = st.bid[j];
      const double inc = (double)st.inc[j], mx = (double)__int_as_float(st.max_inc[t]);
      if (inc - 1e-6 <= mx && mx <= inc + 1e-6) atomicMin(&st.winner[t], j);
    }
    __syncthreads();
    // ---- Assign
    for (int u = tid; u < cnt; u += kEmdThreads) {
      const int j = st.un[u], t = st.bid[j];
      if (last || st.winner[t] == j) {
        const int old = st.assign_inv[t];
        if (!last && old != -1) st.assign[old] = -1;
        st.assign_inv[t] = j;
        st.assign[j] = t;
        st.price[t] = __fadd_rn(st.price[t], st.inc[j]);
        st.max_inc[t] = neg_big;
      }
    }
    __syncthreads();
    // reset the winners of the objects that were bid on (after every Assign thread has read them)
    for (int u = tid; u < cnt; u += kEmdThreads) st.winner[st.bid[st.un[u]]] = 0x7fffffff;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
  }
  for (int j = tid; j < n; j += kEmdThreads) {
    const int k = st.assign[j];
    float d2 = 0.f;
    if (k >= 0) {
      const float dx = __fsub_rn(X1[3 * j], st.x[k]), dy = __fsub_rn(X1[3 * j + 1], st.y[k]), dz = __fsub_rn(X1[3 * j + 2], st.z[k]);
      d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    dist[(size_t)b * n + j] = d2;
    assignment[(size_t)b * n + j] = k;
  }
}

// NmDistanceGradKernel (emd_cuda.cu:283-300): grad_xyz1 = 2 g (x1 - x2[assignment])
__global__ void emd_bwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                               const int* __restrict__ assignment, const float* __restrict__ gdist,
                               float* __restrict__ gxyz1, size_t total, int n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / n;
  const int k = assignment[i];
  const float g = gdist[i] * 2.0f;
  const float* a = xyz1 + 3 * i;
  const float* t = xyz2 + 3 * (b * n + (k < 0 ? 0 : k));
  gxyz1[3 * i] = g * (a[0] - t[0]); gxyz1[3 * i + 1] = g * (a[1] - t[1]); gxyz1[3 * i + 2] = g * (a[2] - t[2]);
}

}  // namespace vpn

using namespace vpn;

extern "C" int vpn_emd_workspace_bytes(int B, int n, size_t* bytes) {
  if (B < 0 || n <= 0 || !bytes) { vpn_set_error("emd workspace: bad arguments"); return VPN_ERR_ARG; }
  *bytes = n <= kEmdSmemMaxN ? 0 : (size_t)B * kEmdArrays * 4 * n;
  return VPN_OK;
}

extern "C" int vpn_emd_fwd(const float* xyz1, const float* xyz2, float* dist, int* assignment, void* workspace,
                           size_t workspace_bytes, int B, int n, float eps, int iters, void* stream) {
  if (B < 0 || n <= 0) { vpn_set_error("emd fwd: bad shape B=%d n=%d", B, n); return VPN_ERR_SHAPE; }
  if (iters < 1 || !(eps >= 0.f)) { vpn_set_error("emd fwd: need iters >= 1 and eps >= 0"); return VPN_ERR_ARG; }
  if (B == 0) return VPN_OK;
  if (!xyz1 || !xyz2 || !dist || !assignment) { vpn_set_error("emd fwd: null pointer"); return VPN_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= kEmdSmemMaxN) {
    static int attr_for = 0;
    const size_t smem = (size_t)kEmdArrays * 4 * n;
    if (attr_for < (int)smem) {
      if (cudaFuncSetAttribute(emd_auction_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmdArrays * 4 * kEmdSmemMaxN) != cudaSuccess) {
        vpn_set_error("emd fwd: smem attribute"); return VPN_ERR_CUDA;
      }
      attr_for = kEmdArrays * 4 * kEmdSmemMaxN;
    }
    emd_auction_kernel<true><<<B, kEmdThreads, smem, s>>>(xyz1, xyz2, dist, assignment, nullptr, n, eps, iters);
  } else {
    const size_t need = (size_t)B * kEmdArrays * 4 * n;
    if (!workspace || workspace_bytes < need) { vpn_set_error("emd fwd: workspace too small (%zu < %zu)", workspace_bytes, need); return VPN_ERR_WORKSPACE; }
    emd_auction_kernel<false><<<B, kEmdThreads, 0, s>>>(xyz1, xyz2, dist, assignment, reinterpret_cast<unsigned char*>(workspace), n, eps, iters);
  }
  return vpn_check_launch("emd_auction_kernel");
}

extern "C" int vpn_emd_bwd(const float* xyz1, const float* xyz2, const int* assignment, const float* grad_dist,
                           float* grad_xyz1, int B, int n, void* stream) {
  if (B < 0 || n <= 0) { vpn_set_error("emd bwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!xyz1 || !xyz2 || !assignment || !grad_dist || !grad_xyz1) { vpn_set_error("emd bwd: null pointer"); return VPN_ERR_ARG; }
  const size_t total = (size_t)B * n;
  emd_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, assignment, grad_dist, grad_xyz1, total, n);
  return vpn_check_launch("emd_bwd_kernel");
}
