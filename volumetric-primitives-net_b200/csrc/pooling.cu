// GCN vertex-feature pooling for train_gcn.py: image bounds + perceptual feature pooling.
//
// Replaces modules/network/gcn.py:84-164:
//   get_bound_of_images        (:90-133)  per sample a Python loop over columns and rows, each iteration reading device
//                                          scalars (`if x_any[i] and bounds[b, 0] == 0`: ~4 host syncs per pixel column)
//   perceptual_feature_pooling (:135-164) per-sample Python loop for the y/z range, then one grid_sample per feature
//                                          map + cat + permute (the (B, sum C, N) intermediate is written and re-read)
// Here: one launch for all bounds, one for the ranges, and per feature map one launch that stages a chunk of channel
// planes in shared memory (every feature element is read once, coalesced) and writes the pooled features directly in
// the (B, N, sum C) layout with 128-byte rows.  HBM-bound on the output: B*N*sumC*4 bytes (train_gcn.py at B = 64:
// 64 x 2048 x 960 x 4 = 503 MB) + the feature maps once.
#include "common.cuh"

namespace vpn {

constexpr int kPoolThreads = 256;
constexpr int kPoolSmemBudget = 96 * 1024;      // backward, per CTA: two CTAs per SM
constexpr int kPoolFwdSmemBudget = 44 * 1024;   // forward, per CTA: five CTAs (40 warps) per SM - the kernel is latency bound

// ---- image bounds ---------------------------------------------------------------------------------------------
// grid: x = sample.  dynamic smem: int flags[W + H].  One CTA of 1024 threads per image, four pixels per thread in
// flight (the kernel is latency bound: B CTAs, 4 C H W bytes each).
constexpr int kBoundsThreads = 1024;
__global__ void __launch_bounds__(kBoundsThreads)
image_bounds_kernel(const float* __restrict__ imgs, float* __restrict__ bounds, int C, int H, int W, float thr) {
  extern __shared__ int pool_flags[];
  __shared__ int lohi[4];
  int* col_any = pool_flags;
  int* row_any = pool_flags + W;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < W + H; i += kBoundsThreads) pool_flags[i] = 0;
  if (tid < 4) lohi[tid] = (tid & 1) ? -1 : 0x7fffffff;
  __syncthreads();
  const float* img = imgs + (size_t)b * C * H * W;
  const int HW = H * W;
  for (int p0 = tid; p0 < HW; p0 += 4 * kBoundsThreads) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {                                              // img.sum(0), channels in order
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kBoundsThreads;
        if (p < HW) s[u] = __fadd_rn(s[u], img[(size_t)c * HW + p]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * kBoundsThreads;
      if (p < HW && s[u] > thr) { col_any[p % W] = 1; row_any[p / W] = 1; }
    }
  }
  __syncthreads();
  // lower bound: first occupied index that is not 0 (gcn.py:107,119: `bounds == 0` doubles as "unset");
  // upper bound: last occupied index
  for (int i = tid; i < W; i += kBoundsThreads) if (col_any[i]) { if (i > 0) atomicMin(&lohi[0], i); atomicMax(&lohi[1], i); }
  for (int i = tid; i < H; i += kBoundsThreads) if (row_any[i]) { if (i > 0) atomicMin(&lohi[2], i); atomicMax(&lohi[3], i); }
  __syncthreads();
  if (tid < 4) {
    const int size = tid < 2 ? W : H;
    const int v = lohi[tid];
    const float raw = (tid & 1) ? (v < 0 ? (float)size : (float)v) : (v == 0x7fffffff ? 0.f : (float)v);
    bounds[4 * b + tid] = __fsub_rn(__fmul_rn(__fdiv_rn(raw, (float)size), 2.0f), 1.0f);      // x / w * 2 - 1
  }
}

// ---- per-sample y / z range of the vertices -----------------------------------------------------------------------
// range[b] = (min z, max z, min y, max y), arg[b] = the (first) vertex attaining each.  grid: x = sample
struct MinMaxIdx { float v; int i; };
__device__ __forceinline__ void take_min(MinMaxIdx& a, float v, int i) { if (v < a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }
__device__ __forceinline__ void take_max(MinMaxIdx& a, float v, int i) { if (v > a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }

__global__ void __launch_bounds__(kPoolThreads)
points_yz_range_kernel(const float* __restrict__ pts, float* __restrict__ range, int* __restrict__ arg, int N) {
  __shared__ float sv[4][kPoolThreads / 32];
  __shared__ int si[4][kPoolThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);
  MinMaxIdx m[4] = {{inf, 0x7fffffff}, {-inf, 0x7fffffff}, {inf, 0x7fffffff}, {-inf, 0x7fffffff}};
  const float* P = pts + (size_t)b * N * 3;
  for (int n = tid; n < N; n += kPoolThreads) {
    const float y = P[3 * (size_t)n + 1], z = P[3 * (size_t)n + 2];
    take_min(m[0], z, n); take_max(m[1], z, n); take_min(m[2], y, n); take_max(m[3], y, n);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, m[k].v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, m[k].i, o);
      if (k & 1) take_max(m[k], ov, oi); else take_min(m[k], ov, oi);
    }
    if (lane == 0) { sv[k][warp] = m[k].v; si[k][warp] = m[k].i; }
  }
  __syncthreads();
  if (tid < 4) {
    MinMaxIdx r = {sv[tid][0], si[tid][0]};
    for (int w = 1; w < kPoolThreads / 32; ++w) { if (tid & 1) take_max(r, sv[tid][w], si[tid][w]); else take_min(r, sv[tid][w], si[tid][w]); }
    range[4 * b + tid] = r.v;
    arg[4 * b + tid] = r.i == 0x7fffffff ? 0 : r.i;
  }
}

// ---- bilinear taps of one vertex in one feature map (grid_sample, bilinear, zeros padding, align_corners=True) ------
struct Taps {
  int off[4];       // y * W + x of nw, ne, sw, se (clamped into the plane)
  float w[4];       // tap weights; 0 for taps outside the plane
  float dx[4];      // d weight / d ix  (0 for taps outside the plane)
  float dy[4];      // d weight / d iy
};

__device__ __forceinline__ void grid_of_vertex(const float* __restrict__ p, const float* __restrict__ bnd,
                                               const float* __restrict__ rng, float& gx, float& gy) {
  // gcn.py:153-154: grid x from z, grid y from y, flipped, rescaled into the image bounds
  const float sz = __fdiv_rn(__fsub_rn(p[2], rng[0]), __fsub_rn(rng[1], rng[0]));
  const float sy = __fdiv_rn(__fsub_rn(p[1], rng[2]), __fsub_rn(rng[3], rng[2]));
  gx = __fadd_rn(bnd[0], __fmul_rn(__fsub_rn(1.0f, sz), __fsub_rn(bnd[1], bnd[0])));
  gy = __fadd_rn(bnd[2], __fmul_rn(__fsub_rn(1.0f, sy), __fsub_rn(bnd[3], bnd[2])));
}

// invalid taps (outside the plane) get weight 0 and offset `pad`: a word the staged planes keep at zero (index H*W), so the
// forward loop needs no predicate; the un-staged paths pass pad = 0 and skip zero weights instead.
__device__ __forceinline__ void make_taps(float gx, float gy, int H, int W, Taps& t, int pad = 0) {
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  const float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
  const float xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
  const float wx[4] = {x1 - ix, ix - x0, x1 - ix, ix - x0}, wy[4] = {y1 - iy, y1 - iy, iy - y0, iy - y0};
  const float sx[4] = {-1.f, 1.f, -1.f, 1.f}, sy[4] = {-1.f, -1.f, 1.f, 1.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // NaN coordinates (degenerate range: max == min) fail every comparison: all taps invalid, output 0 like grid_sample
    const bool ok = xs[k] >= 0.f && xs[k] <= (float)(W - 1) && ys[k] >= 0.f && ys[k] <= (float)(H - 1);
    const int xi = ok ? (int)xs[k] : 0, yi = ok ? (int)ys[k] : 0;
    t.off[k] = ok ? yi * W + xi : pad;
    t.w[k] = ok ? wx[k] * wy[k] : 0.f;
    t.dx[k] = ok ? sx[k] * wy[k] : 0.f;
    t.dy[k] = ok ? sy[k] * wx[k] : 0.f;
  }
}

// Forward.  grid: x = channel chunk (CH channels), y = sample, z = vertex slice.  dynamic smem: CH planes of `stride`
// floats (staged == 1).  A warp takes 32 vertices at a time: lane j prepares the taps of vertex j, then the warp walks
// the 32 vertices with the lanes spread over channels.  CH < 32 (large planes): G = 32 / CH vertices per pass.
// CH = 32 KC (small planes: many channels fit): one vertex per pass, every lane owns KC channels, so the tap
// broadcast (8 shuffles) is paid once per KC x 32 outputs; a pass writes KC rows of 128 bytes.
__global__ void __launch_bounds__(kPoolThreads)
feature_pool_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ pts, const float* __restrict__ bounds,
                        const float* __restrict__ range, float* __restrict__ out,
                        int C, int H, int W, int N, int Ctot, int coff, int CH, int stride, int staged, int per_slice) {
  extern __shared__ float pool_planes[];
  const int b = blockIdx.y, c0 = blockIdx.x * CH, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = H * W;
  const int nch = min(CH, C - c0);
  const float* fb = feat + ((size_t)b * C + c0) * HW;
  if (staged) {
    for (int i = tid; i < nch * HW; i += kPoolThreads) pool_planes[(i / HW) * stride + (i % HW)] = fb[i];
    for (int i = tid; i < CH; i += kPoolThreads) pool_planes[i * stride + HW] = 0.f;       // the pad word of every plane
    __syncthreads();
  }
  const float* planes = staged ? pool_planes : fb;
  const int pstride = staged ? stride : HW;
  const float* bnd = bounds + 4 * b;
  const float* rng = range + 4 * b;
  const int LCH = min(CH, 32), KC = (CH + 31) / 32;
  const int G = 32 / LCH, sub = lane / LCH, c = lane % LCH;
  const int n_lo = blockIdx.z * per_slice, n_hi = min(N, n_lo + per_slice);
  for (int base = n_lo + warp * 32; base < n_hi; base += (kPoolThreads / 32) * 32) {
    Taps mine;
    {
      const int n = min(base + lane, N - 1);
      float gx, gy;
      grid_of_vertex(pts + ((size_t)b * N + n) * 3, bnd, rng, gx, gy);
      make_taps(gx, gy, H, W, mine, staged ? HW : 0);
    }
    for (int j0 = 0; j0 < 32; j0 += G) {
      const int j = j0 + sub;
      int off[4]; float w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { off[k] = __shfl_sync(0xffffffffu, mine.off[k], j); w[k] = __shfl_sync(0xffffffffu, mine.w[k], j); }
      const int n = base + j;
      if (n >= n_hi) continue;                     // after the shuffles: every lane took part in them
      float* orow = out + ((size_t)b * N + n) * Ctot + coff + c0;
      if (staged) {
        // every tap is a shared-memory word (invalid ones the zero pad): no predicates, two channels in flight
        const float* pl = planes + (size_t)c * pstride;
        const size_t step = (size_t)32 * pstride;
        int kc = 0;
        for (; kc + 1 < KC; kc += 2, pl += 2 * step) {
          const float* p2 = pl + step;
          const float a = fmaf(pl[off[3]], w[3], fmaf(pl[off[2]], w[2], fmaf(pl[off[1]], w[1], pl[off[0]] * w[0])));
          const float d = fmaf(p2[off[3]], w[3], fmaf(p2[off[2]], w[2], fmaf(p2[off[1]], w[1], p2[off[0]] * w[0])));
          const int cc = c + 32 * kc;
          if (cc < nch) orow[cc] = a;
          if (cc + 32 < nch) orow[cc + 32] = d;
        }
        if (kc < KC) {
          const int cc = c + 32 * kc;
          if (cc < nch) orow[cc] = fmaf(pl[off[3]], w[3], fmaf(pl[off[2]], w[2], fmaf(pl[off[1]], w[1], pl[off[0]] * w[0])));
        }
      } else {
        for (int kc = 0; kc < KC; ++kc) {
          const int cc = c + 32 * kc;
          if (cc < nch) {
            const float* pl = planes + (size_t)cc * pstride;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (w[k] != 0.f) acc = fmaf(pl[off[k]], w[k], acc);
            orow[cc] = acc;
          }
        }
      }
    }
  }
}

// Backward.  Same decomposition; dynamic smem: CH feature planes + CH gradient planes.  The gradient planes are
// accumulated with shared-memory atomics and stored once (every (sample, channel) plane belongs to exactly one CTA, so
// grad_feat needs no zero fill and no global atomics); the grid gradient of a vertex is reduced over the CTA's channels
// with shuffles and added to grad_grid (B, N, 2) with one global atomic pair per vertex and CTA.
__global__ void __launch_bounds__(kPoolThreads)
feature_pool_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ pts, const float* __restrict__ bounds,
                        const float* __restrict__ range, const float* __restrict__ gout, float* __restrict__ gfeat,
                        float* __restrict__ ggrid, int C, int H, int W, int N, int Ctot, int coff, int CH, int stride,
                        int staged) {
  extern __shared__ float pool_planes[];
  const int b = blockIdx.y, c0 = blockIdx.x * CH, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = H * W;
  const int nch = min(CH, C - c0);
  const float* fb = feat + ((size_t)b * C + c0) * HW;
  float* gb = gfeat + ((size_t)b * C + c0) * HW;
  // staged == 0 (a plane larger than shared memory): features read through L2, gradients added to the zero-filled
  // grad_feat with global atomics
  const float* fpl = staged ? pool_planes : fb;
  float* gpl = staged ? pool_planes + (size_t)CH * stride : gb;
  if (staged) {
    for (int i = tid; i < nch * HW; i += kPoolThreads) {
      const int o = (i / HW) * stride + (i % HW);
      pool_planes[o] = fb[i]; gpl[o] = 0.f;
    }
    __syncthreads();
  } else {
    stride = HW;
  }
  const float* bnd = bounds + 4 * b;
  const float* rng = range + 4 * b;
  const int G = 32 / CH, sub = lane / CH, c = lane % CH;
  const float mx = 0.5f * (float)(W - 1), my = 0.5f * (float)(H - 1);       // d ix / d grid x (align_corners=True)
  for (int base = warp * 32; base < N; base += (kPoolThreads / 32) * 32) {
    Taps mine;
    {
      const int n = min(base + lane, N - 1);
      float gx, gy;
      grid_of_vertex(pts + ((size_t)b * N + n) * 3, bnd, rng, gx, gy);
      make_taps(gx, gy, H, W, mine);
    }
    for (int j0 = 0; j0 < 32; j0 += G) {
      const int j = j0 + sub;
      const int n = base + j;
      const bool live = n < N && c < nch;
      const float g = live ? gout[((size_t)b * N + n) * Ctot + coff + c0 + c] : 0.f;
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int off = __shfl_sync(0xffffffffu, mine.off[k], j);
        const float w = __shfl_sync(0xffffffffu, mine.w[k], j);
        const float dx = __shfl_sync(0xffffffffu, mine.dx[k], j);
        const float dy = __shfl_sync(0xffffffffu, mine.dy[k], j);
        if (live && (w != 0.f || dx != 0.f || dy != 0.f)) {
          const size_t o = (size_t)c * stride + off;
          if (w != 0.f) atomicAdd(&gpl[o], g * w);
          const float f = fpl[o] * g;
          gix = fmaf(f, dx, gix); giy = fmaf(f, dy, giy);
        }
      }
      // sum over the CH channels of this vertex (lanes of one sub-group are contiguous and CH is a power of two)
      for (int o = CH >> 1; o > 0; o >>= 1) { gix += __shfl_xor_sync(0xffffffffu, gix, o); giy += __shfl_xor_sync(0xffffffffu, giy, o); }
      if (n < N && c == 0) {
        atomicAdd(&ggrid[((size_t)b * N + n) * 2], gix * mx);
        atomicAdd(&ggrid[((size_t)b * N + n) * 2 + 1], giy * my);
      }
    }
  }
  if (!staged) return;
  __syncthreads();
  for (int i = tid; i < nch * HW; i += kPoolThreads) gb[i] = gpl[(i / HW) * stride + (i % HW)];
}

// ---- backward, round 2: vertices sorted by bilinear cell, register accumulation per cell --------------------------
// The first backward (feature_pool_bwd_kernel above, still behind vpn_feature_pool_bwd) spent its time on per-vertex tap broadcasts (16 shuffles), shared-memory float
// atomics (compare-and-swap loops) and a shuffle reduction per (vertex, 32-channel chunk): ~90 instructions per
// (vertex, chunk), 1.85 ms at train_gcn's shape.  Here:
//   feature_pool_sort_kernel   one CTA per sample: the four taps of every vertex and a counting sort of the vertices by
//                              the bilinear cell (x0, y0) they fall in.  All vertices of a cell share their four texels.
//   feature_pool_bwd2_kernel   a warp walks a contiguous run of the sorted list with the lanes over channels (32 KCS per
//                              CTA): per vertex and chunk ONE coalesced 128-byte load of grad_out and 14 FMAs - the
//                              texel gradients of the current cell accumulate in registers, the cell's features stay in
//                              registers - and one shuffle reduction per vertex and CTA for the grid gradient.  When the
//                              cell changes, the four accumulators per channel are flushed with global float atomics
//                              (single RED instructions) into the zero-filled grad_feat.
struct __align__(16) PoolRec { int off[4]; float w[4]; float dx[4]; float dy[4]; };

// grid: x = sample; kPoolThreads threads; dynamic smem: (H + 1) (W + 1) + 1 ints
__global__ void __launch_bounds__(kPoolThreads)
feature_pool_sort_kernel(const float* __restrict__ pts, const float* __restrict__ bounds, const float* __restrict__ range,
                         PoolRec* __restrict__ rec, int* __restrict__ key_out, int* __restrict__ vid_out, int H, int W, int N) {
  extern __shared__ int pool_hist[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ncell = (H + 1) * (W + 1);
  for (int i = tid; i <= ncell; i += kPoolThreads) pool_hist[i] = 0;
  __syncthreads();
  const float* bnd = bounds + 4 * b;
  const float* rng = range + 4 * b;
  int valid_mask = 0;
  auto cell_of = [&](int n, Taps& t) {
    float gx, gy;
    grid_of_vertex(pts + ((size_t)b * N + n) * 3, bnd, rng, gx, gy);
    make_taps(gx, gy, H, W, t);
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
    // cell (x0, y0) = floor of the pixel coordinates, shifted by one so that the border cells (x0 = -1) get key >= 0;
    // coordinates outside [-1, W) / NaN share cell 0 - all their taps are invalid anyway
    const float x0 = floorf(ix), y0 = floorf(iy);
    const bool in = x0 >= -1.f && x0 <= (float)(W - 1) && y0 >= -1.f && y0 <= (float)(H - 1);
    // which of the four taps (nw, ne, sw, se) lie inside the plane: a property of the cell, kept in the key's top bits
    // (a tap's coefficients can all vanish for ONE vertex - one that sits exactly on a texel - while the tap is valid)
    valid_mask = 0;
    if (in) {
      const bool xa = x0 >= 0.f, xb = x0 + 1.0f <= (float)(W - 1), ya = y0 >= 0.f, yb = y0 + 1.0f <= (float)(H - 1);
      valid_mask = (xa && ya ? 1 : 0) | (xb && ya ? 2 : 0) | (xa && yb ? 4 : 0) | (xb && yb ? 8 : 0);
    }
    return in ? ((int)y0 + 1) * (W + 1) + ((int)x0 + 1) : 0;
  };
  for (int n = tid; n < N; n += kPoolThreads) { Taps t; atomicAdd(&pool_hist[cell_of(n, t) + 1], 1); }
  __syncthreads();
  if (tid == 0) { int run = 0; for (int i = 1; i <= ncell; ++i) { run += pool_hist[i]; pool_hist[i] = run; } }      // hist[c] = start of cell c
  __syncthreads();
  for (int n = tid; n < N; n += kPoolThreads) {
    Taps t;
    const int key = cell_of(n, t);
    const int slot = atomicAdd(&pool_hist[key], 1);                                // order inside a cell: whichever thread comes first
    const size_t o = (size_t)b * N + slot;
    PoolRec r;
#pragma unroll
    for (int k = 0; k < 4; ++k) { r.off[k] = t.off[k]; r.w[k] = t.w[k]; r.dx[k] = t.dx[k]; r.dy[k] = t.dy[k]; }
    rec[o] = r; key_out[o] = key | (valid_mask << 24); vid_out[o] = n;
  }
}

// grid: x = run of the sorted vertex list, y = channel slab (32 KCS channels), z = sample
template <int KCS>
__global__ void __launch_bounds__(kPoolThreads)
feature_pool_bwd2_kernel(const float* __restrict__ feat, const float* __restrict__ gout, const PoolRec* __restrict__ rec,
                         const int* __restrict__ keys, const int* __restrict__ vids, float* __restrict__ gfeat,
                         float* __restrict__ ggrid, int C, int H, int W, int N, int Ctot, int coff, int per_cta) {
  const int b = blockIdx.z, c0 = blockIdx.y * 32 * KCS, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W;
  const int per_warp = (per_cta + kPoolThreads / 32 - 1) / (kPoolThreads / 32);
  const int p0 = blockIdx.x * per_cta + warp * per_warp;
  const int p1 = min(min(p0 + per_warp, (int)(blockIdx.x + 1) * per_cta), N);
  const float mx = 0.5f * (float)(W - 1), my = 0.5f * (float)(H - 1);             // d ix / d grid x (align_corners=True)
  bool chan[KCS];
  const float* fch[KCS]; float* gch[KCS];
#pragma unroll
  for (int kc = 0; kc < KCS; ++kc) {
    const int c = c0 + kc * 32 + lane;
    chan[kc] = c < C;
    const size_t plane = ((size_t)b * C + (chan[kc] ? c : 0)) * HW;
    fch[kc] = feat + plane; gch[kc] = gfeat + plane;
  }
  float acc[KCS][4], f[KCS][4];
  int cur_key = -1, cur_off[4] = {0, 0, 0, 0};
  bool cur_ok[4] = {false, false, false, false};
  auto flush = [&]() {
#pragma unroll
    for (int kc = 0; kc < KCS; ++kc)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (chan[kc] && cur_ok[k] && acc[kc][k] != 0.f) atomicAdd(gch[kc] + cur_off[k], acc[kc][k]);
  };
  // two-deep software pipeline: a vertex's record (rec, key, vertex id) is loaded two iterations ahead, its grad_out
  // words - whose address needs the vertex id - one iteration ahead; a warp walks its run serially, so without this every
  // vertex paid an L2 round trip plus an HBM round trip
  PoolRec r0, r1; int key0 = 0, key1 = 0, n0 = 0, n1 = 0;
  float g0[KCS];
  auto load_rec = [&](int p, PoolRec& r, int& key, int& n) {
    const size_t o = (size_t)b * N + p;
    r = rec[o]; key = keys[o]; n = vids[o];                                         // same address on every lane: broadcast
  };
  auto load_g = [&](int n, float (&g)[KCS]) {
    const float* grow = gout + ((size_t)b * N + n) * Ctot + coff + c0 + lane;
#pragma unroll
    for (int kc = 0; kc < KCS; ++kc) g[kc] = chan[kc] ? grow[kc * 32] : 0.f;
  };
  if (p0 < p1) { load_rec(p0, r0, key0, n0); load_g(n0, g0); }
  if (p0 + 1 < p1) load_rec(p0 + 1, r1, key1, n1);
  for (int p = p0; p < p1; ++p) {
    const PoolRec r = r0;
    const int key = key0, n = n0;
    float g[KCS];
#pragma unroll
    for (int kc = 0; kc < KCS; ++kc) g[kc] = g0[kc];
    if (p + 1 < p1) { r0 = r1; key0 = key1; n0 = n1; load_g(n0, g0); }
    if (p + 2 < p1) load_rec(p + 2, r1, key1, n1);
    if (key != cur_key) {
      if (cur_key >= 0) flush();
      cur_key = key;
#pragma unroll
      for (int k = 0; k < 4; ++k) { cur_off[k] = r.off[k]; cur_ok[k] = (key >> (24 + k)) & 1; }      // validity: per cell, in the key
#pragma unroll
      for (int kc = 0; kc < KCS; ++kc)
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[kc][k] = 0.f; f[kc][k] = (chan[kc] && cur_ok[k]) ? __ldg(fch[kc] + cur_off[k]) : 0.f; }
    }
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int kc = 0; kc < KCS; ++kc) {
      float sx = 0.f, sy = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sx = fmaf(f[kc][k], r.dx[k], sx); sy = fmaf(f[kc][k], r.dy[k], sy);
        acc[kc][k] = fmaf(g[kc], r.w[k], acc[kc][k]);
      }
      gix = fmaf(g[kc], sx, gix); giy = fmaf(g[kc], sy, giy);
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) { gix += __shfl_xor_sync(0xffffffffu, gix, o2); giy += __shfl_xor_sync(0xffffffffu, giy, o2); }
    if (lane == 0) {
      atomicAdd(&ggrid[((size_t)b * N + n) * 2], gix * mx);
      atomicAdd(&ggrid[((size_t)b * N + n) * 2 + 1], giy * my);
    }
  }
  if (cur_key >= 0) flush();
}

// grad_grid (B, N, 2) -> grad_points (B, N, 3), including the terms that reach the arg-min / arg-max vertices through
// the range (autograd of gcn.py:146-154: points.max(0) / .min(0) are differentiable).  grid: x = sample
__global__ void __launch_bounds__(kPoolThreads)
feature_pool_points_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ bounds, const float* __restrict__ range,
                               const int* __restrict__ arg, const float* __restrict__ ggrid, float* __restrict__ gpts, int N) {
  __shared__ float red[4][kPoolThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* bnd = bounds + 4 * b;
  const float* rng = range + 4 * b;
  const float Dz = bnd[1] - bnd[0], Dy = bnd[3] - bnd[2];
  const float Rz = rng[1] - rng[0], Ry = rng[3] - rng[2];
  // grid = b0 + (1 - num / den) D:   d/dnum = -D / den,   d/dden = D num / den^2
  float acc[4] = {0.f, 0.f, 0.f, 0.f};      // z: sum d/dnum, sum d/dden;  y: the same
  for (int n = tid; n < N; n += kPoolThreads) {
    const size_t o = (size_t)b * N + n;
    const float gx = ggrid[2 * o], gy = ggrid[2 * o + 1];
    const float numz = pts[3 * o + 2] - rng[0], numy = pts[3 * o + 1] - rng[2];
    const float dnz = gx * (-Dz / Rz), dny = gy * (-Dy / Ry);
    gpts[3 * o] = 0.f; gpts[3 * o + 1] = dny; gpts[3 * o + 2] = dnz;
    acc[0] += dnz; acc[1] += gx * Dz * numz / (Rz * Rz);
    acc[2] += dny; acc[3] += gy * Dy * numy / (Ry * Ry);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float s = warp_sum(acc[k]); if (lane == 0) red[k][warp] = s; }
  __syncthreads();
  if (tid == 0) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < kPoolThreads / 32; ++w) for (int k = 0; k < 4; ++k) s[k] += red[k][w];
    // num = p - min, den = max - min:  d/dmin = -(sum d/dnum) - (sum d/dden),  d/dmax = sum d/dden
    float* g = gpts + (size_t)b * N * 3;
    g[3 * (size_t)arg[4 * b + 0] + 2] += -s[0] - s[1];
    g[3 * (size_t)arg[4 * b + 1] + 2] += s[1];
    g[3 * (size_t)arg[4 * b + 2] + 1] += -s[2] - s[3];
    g[3 * (size_t)arg[4 * b + 3] + 1] += s[3];
  }
}

// channels per CTA (power of two <= 32) such that `planes_per_channel` padded planes fit the budget; 0 = does not fit
static int pool_chunk_budget(int C, int HW, int planes_per_channel, int* stride, int budget) {
  *stride = (HW + 1) | 1;                            // >= one pad word; odd: the 32 channel lanes hit 32 different banks
  int ch = 32;
  while (ch > 1 && (ch >> 1) >= C) ch >>= 1;         // no wider than the map needs
  while (ch >= 1 && (size_t)ch * planes_per_channel * (*stride) * 4 > (size_t)budget) ch >>= 1;
  return ch;
}
static int pool_chunk(int C, int HW, int planes_per_channel, int* stride) {
  return pool_chunk_budget(C, HW, planes_per_channel, stride, kPoolSmemBudget);
}

// forward: as above, then widened to 32 KC channels (KC <= 16) while the planes still fit
static int pool_chunk_fwd(int C, int HW, int* stride) {
  int ch = pool_chunk_budget(C, HW, 1, stride, kPoolFwdSmemBudget);
  if (ch == 32) while (ch < C && ch < 512 && (size_t)2 * ch * (*stride) * 4 <= (size_t)kPoolFwdSmemBudget) ch *= 2;
  return ch;
}

}  // namespace vpn

using namespace vpn;

// imgs (B,C,H,W) -> bounds (B,4) = [x lo, x hi, y lo, y hi] in [-1, 1]   (gcn.py:90-133)
extern "C" int vpn_image_bounds(const float* imgs, float* bounds, int B, int C, int H, int W, float threshold, void* stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) { vpn_set_error("image bounds: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!imgs || !bounds) { vpn_set_error("image bounds: null pointer"); return VPN_ERR_ARG; }
  if ((size_t)(W + H) * 4 > 48 * 1024) { vpn_set_error("image bounds: W + H too large"); return VPN_ERR_SHAPE; }
  image_bounds_kernel<<<B, kBoundsThreads, (size_t)(W + H) * 4, (cudaStream_t)stream>>>(imgs, bounds, C, H, W, threshold);
  return vpn_check_launch("image_bounds_kernel");
}

// pts (B,N,3) -> range (B,4) = [min z, max z, min y, max y], arg (B,4) = first vertex attaining each   (gcn.py:146-150)
extern "C" int vpn_points_yz_range(const float* pts, float* range, int* arg, int B, int N, void* stream) {
  if (B < 0 || N <= 0) { vpn_set_error("yz range: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!pts || !range || !arg) { vpn_set_error("yz range: null pointer"); return VPN_ERR_ARG; }
  points_yz_range_kernel<<<B, kPoolThreads, 0, (cudaStream_t)stream>>>(pts, range, arg, N);
  return vpn_check_launch("points_yz_range_kernel");
}

// One feature map feat (B,C,H,W) pooled at the vertices into out[:, :, coff : coff + C] of out (B,N,Ctot)   (gcn.py:153-163)
extern "C" int vpn_feature_pool_fwd(const float* feat, const float* pts, const float* bounds, const float* range, float* out,
                                    int B, int C, int H, int W, int N, int Ctot, int coff, void* stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || coff < 0 || coff + C > Ctot) { vpn_set_error("feature pool fwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!feat || !pts || !bounds || !range || !out) { vpn_set_error("feature pool fwd: null pointer"); return VPN_ERR_ARG; }
  int stride = 0;
  int ch = pool_chunk_fwd(C, H * W, &stride);
  const int staged = ch >= 1;
  if (!staged) ch = 32;                                            // plane larger than shared memory: taps read through L2
  const size_t smem = staged ? (size_t)ch * stride * 4 : 0;
  static DeviceOnce once;
  {
    cudaError_t e = set_dyn_smem(feature_pool_fwd_kernel, kPoolFwdSmemBudget, once);
    if (e != cudaSuccess) { vpn_set_error("feature pool fwd: smem attribute: %s", cudaGetErrorString(e)); return VPN_ERR_CUDA; }
  }
  // enough CTAs for two waves of five per SM: split the vertices when (channel chunks x samples) alone are too few
  const int chunks = (C + ch - 1) / ch;
  int slices = (10 * device_sm_count() + chunks * B - 1) / (chunks * B);
  const int max_slices = (N + kPoolThreads - 1) / kPoolThreads;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  const int per_slice = (((N + slices - 1) / slices) + 31) / 32 * 32;
  dim3 grid(chunks, B, (N + per_slice - 1) / per_slice);
  feature_pool_fwd_kernel<<<grid, kPoolThreads, smem, (cudaStream_t)stream>>>(feat, pts, bounds, range, out, C, H, W, N, Ctot, coff,
                                                                             ch, stride, staged, per_slice);
  return vpn_check_launch("feature_pool_fwd_kernel");
}

// Gradient of one feature map's slice: grad_feat (B,C,H,W) is fully written; grad_grid (B,N,2) is ACCUMULATED (the caller
// zeroes it before the first map and runs vpn_feature_pool_points_bwd after the last).
extern "C" int vpn_feature_pool_bwd(const float* feat, const float* pts, const float* bounds, const float* range,
                                    const float* grad_out, float* grad_feat, float* grad_grid,
                                    int B, int C, int H, int W, int N, int Ctot, int coff, void* stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || coff < 0 || coff + C > Ctot) { vpn_set_error("feature pool bwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!feat || !pts || !bounds || !range || !grad_out || !grad_feat || !grad_grid) { vpn_set_error("feature pool bwd: null pointer"); return VPN_ERR_ARG; }
  int stride = 0;
  int ch = pool_chunk(C, H * W, 2, &stride);
  const int staged = ch >= 1;
  if (!staged) {
    ch = 32;
    if (cudaMemsetAsync(grad_feat, 0, (size_t)B * C * H * W * 4, (cudaStream_t)stream) != cudaSuccess) {
      vpn_set_error("feature pool bwd: memset failed"); return VPN_ERR_CUDA;
    }
  }
  static DeviceOnce once;
  {
    cudaError_t e = set_dyn_smem(feature_pool_bwd_kernel, kPoolSmemBudget, once);
    if (e != cudaSuccess) { vpn_set_error("feature pool bwd: smem attribute: %s", cudaGetErrorString(e)); return VPN_ERR_CUDA; }
  }
  dim3 grid((C + ch - 1) / ch, B);
  feature_pool_bwd_kernel<<<grid, kPoolThreads, staged ? (size_t)2 * ch * stride * 4 : 0, (cudaStream_t)stream>>>(
      feat, pts, bounds, range, grad_out, grad_feat, grad_grid, C, H, W, N, Ctot, coff, ch, stride, staged);
  return vpn_check_launch("feature_pool_bwd_kernel");
}

extern "C" int vpn_feature_pool_bwd_workspace_bytes(int B, int N, size_t* bytes) {
  if (B < 0 || N <= 0 || !bytes) { vpn_set_error("feature pool bwd workspace: bad arguments"); return VPN_ERR_ARG; }
  *bytes = (size_t)B * N * (sizeof(PoolRec) + 8) + 256;
  return VPN_OK;
}

// Same contract as vpn_feature_pool_bwd (grad_feat fully written, grad_grid ACCUMULATED), through the cell-sorted kernels.
// workspace: vpn_feature_pool_bwd_workspace_bytes(B, N) bytes of scratch (reused for every map).
extern "C" int vpn_feature_pool_bwd_sorted(const float* feat, const float* pts, const float* bounds, const float* range,
                                           const float* grad_out, float* grad_feat, float* grad_grid, void* workspace,
                                           size_t workspace_bytes, int B, int C, int H, int W, int N, int Ctot, int coff,
                                           void* stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || coff < 0 || coff + C > Ctot) { vpn_set_error("feature pool bwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (B > 65535) { vpn_set_error("feature pool bwd: batch > 65535 unsupported"); return VPN_ERR_SHAPE; }
  if (!feat || !pts || !bounds || !range || !grad_out || !grad_feat || !grad_grid || !workspace) { vpn_set_error("feature pool bwd: null pointer"); return VPN_ERR_ARG; }
  const size_t need = (size_t)B * N * (sizeof(PoolRec) + 8) + 256;
  const size_t hist_bytes = ((size_t)(H + 1) * (W + 1) + 1) * sizeof(int);
  if (workspace_bytes < need) { vpn_set_error("feature pool bwd: workspace too small"); return VPN_ERR_WORKSPACE; }
  if (hist_bytes > 200 * 1024) { vpn_set_error("feature pool bwd: feature map too large for the cell histogram"); return VPN_ERR_SHAPE; }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  PoolRec* rec = reinterpret_cast<PoolRec*>(ws);
  int* keys = reinterpret_cast<int*>(ws + (((size_t)B * N * sizeof(PoolRec) + 255) & ~(size_t)255));
  int* vids = keys + (size_t)B * N;
  if (cudaMemsetAsync(grad_feat, 0, (size_t)B * C * H * W * 4, s) != cudaSuccess) { vpn_set_error("feature pool bwd: memset failed"); return VPN_ERR_CUDA; }
  static DeviceOnce once;
  if (hist_bytes > 48 * 1024 && set_dyn_smem(feature_pool_sort_kernel, 200 * 1024, once) != cudaSuccess) {
    vpn_set_error("feature pool bwd: smem attribute"); return VPN_ERR_CUDA;
  }
  feature_pool_sort_kernel<<<B, kPoolThreads, hist_bytes, s>>>(pts, bounds, range, rec, keys, vids, H, W, N);
  int rc = vpn_check_launch("feature_pool_sort_kernel");
  if (rc) return rc;
  // 32 KCS channels per CTA; enough runs of the sorted list for ~4 CTAs per SM, at least 64 vertices per CTA
  const int kcs = C > 64 ? 4 : (C > 32 ? 2 : 1);
  const int slabs = (C + 32 * kcs - 1) / (32 * kcs);
  int runs = (4 * device_sm_count() + slabs * B - 1) / (slabs * B);
  const int max_runs = (N + 63) / 64;
  if (runs > max_runs) runs = max_runs;
  if (runs < 1) runs = 1;
  const int per_cta = (N + runs - 1) / runs;
  dim3 grid((N + per_cta - 1) / per_cta, slabs, B);
  switch (kcs) {
    case 4:  feature_pool_bwd2_kernel<4><<<grid, kPoolThreads, 0, s>>>(feat, grad_out, rec, keys, vids, grad_feat, grad_grid, C, H, W, N, Ctot, coff, per_cta); break;
    case 2:  feature_pool_bwd2_kernel<2><<<grid, kPoolThreads, 0, s>>>(feat, grad_out, rec, keys, vids, grad_feat, grad_grid, C, H, W, N, Ctot, coff, per_cta); break;
    default: feature_pool_bwd2_kernel<1><<<grid, kPoolThreads, 0, s>>>(feat, grad_out, rec, keys, vids, grad_feat, grad_grid, C, H, W, N, Ctot, coff, per_cta); break;
  }
  return vpn_check_launch("feature_pool_bwd2_kernel");
}

extern "C" int vpn_feature_pool_points_bwd(const float* pts, const float* bounds, const float* range, const int* arg,
                                           const float* grad_grid, float* grad_pts, int B, int N, void* stream) {
  if (B < 0 || N <= 0) { vpn_set_error("feature pool points bwd: bad shape"); return VPN_ERR_SHAPE; }
  if (B == 0) return VPN_OK;
  if (!pts || !bounds || !range || !arg || !grad_grid || !grad_pts) { vpn_set_error("feature pool points bwd: null pointer"); return VPN_ERR_ARG; }
  feature_pool_points_bwd_kernel<<<B, kPoolThreads, 0, (cudaStream_t)stream>>>(pts, bounds, range, arg, grad_grid, grad_pts, N);
  return vpn_check_launch("feature_pool_points_bwd_kernel");
}
