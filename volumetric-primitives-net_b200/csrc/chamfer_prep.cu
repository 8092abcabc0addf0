// Spatial preparation for the tensor-core Chamfer filter (chamfer_tc.cu): what lets it SKIP distance blocks.
//
// The filter evaluates the (P x M) distance matrix of modules/loss/chamfer_distance.py:14-23 in blocks of 128 rows x
// 128 columns.  A block cannot hold any row's (column's) nearest neighbour when the two point sets are further apart
// than a distance some point of the block is already known to achieve.  Two small kernels provide the ingredients:
//
//   chamfer_sort_targets_kernel   puts BOTH clouds in spatial order, one launch.
//       CTA b < B: sample b's targets by Morton cell (16 cells per axis of the sample's bounding box, 12 bits) with a
//       single-pass counting sort (prep_cell_sort); writes the sorted copies p2s / p2v, the permutation perm (sorted
//       position -> original index), the axis-aligned box of every 128-column chunk, and the largest |coordinate| (the
//       filter's scale; NaN if any coordinate is).  After sorting a chunk is a compact patch of the target shape.
//       The CTAs after them sort the PREDICTED points: they arrive primitive-major (train.py:119 torch.cat(dim=1)) but
//       in the random order of the surface samples, so 128 consecutive rows cover whole faces of a primitive.  Every
//       segment of 4096 consecutive rows (one primitive at the training sizes) is sorted on its own, inside the
//       segment's box: a block of 128 sorted rows is a compact patch, its box small, and far more blocks prune (ideal
//       share of live blocks on the C2 clouds: 27 % unsorted, 18 % sorted; measured 32 % -> 22 %).  A segment whose
//       natural order is already the more compact one (mesh vertices: one small primitive per block) keeps it.
//       Output: the sorted copy p1s, rperm (sorted row -> original row) and the box of every 128-row block.
//   chamfer_prune_bounds_kernel   per 128-row block: T_r = max over its rows of an UPPER bound of the row's
//       nearest-target distance (distance to 4 representatives of each of the 8 chunks whose boxes are nearest to the
//       block's box); per 128-column chunk: U_c likewise over 32 row blocks x 4 rows.
//
// chamfer_tc_kernel skips the block (row block r, chunk c) for the row direction when gap(box_r, box_c)^2 > T_r and for
// the column direction when gap^2 > U_c (both with a 1e-5 relative margin): every pair in the block is then further
// apart than a distance row i (column j) certainly achieves elsewhere, so neither its arg-min nor a tie can be there.
// Results are unchanged bit for bit (the exact recovery kernels still decide); first-index ties survive both
// permutations because the recovery keys carry the ORIGINAL target / row index, and the row results are written back
// through rperm.
#include "common.cuh"

namespace vpn {

constexpr int kSortThreads = 1024;
constexpr int kSortMaxM = 16384;          // targets per sample the sort handles: 16 items per thread, ent[] of 64 KB in shared memory
constexpr int kBlk = 128;                 // rows per block = columns per chunk

__device__ __forceinline__ float prep_inf() { return __int_as_float(0x7f800000); }
__device__ __forceinline__ unsigned spread6(unsigned x) {       // 6 bits -> every third bit
  x = (x | (x << 8)) & 0x0000300Fu;
  x = (x | (x << 4)) & 0x000030C3u;
  x = (x | (x << 2)) & 0x00009249u;
  return x;
}
__device__ __forceinline__ u64 prep_pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void prep_upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 prep_sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 prep_mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 prep_fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// warp-wide float min / max in one instruction (redux.sync.f32 is sm_100a; NaN inputs are dropped, like fminf / fmaxf)
__device__ __forceinline__ float warp_min_f(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float warp_max_f(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }

// Box of the CTA's points from per-thread (lo, hi): warp redux, then warp 0 over the 32 warps.  Result in red[0][0..5];
// amu (optional, bits of the largest |coordinate|, NaN above everything) likewise in red[0][6].  Two barriers.
__device__ __forceinline__ void prep_block_box(float (&lo)[3], float (&hi)[3], unsigned amu, float (*red)[7]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) { lo[k] = warp_min_f(lo[k]); hi[k] = warp_max_f(hi[k]); }
  amu = __reduce_max_sync(0xffffffffu, amu);
  if (lane == 0) { for (int k = 0; k < 3; ++k) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; } red[warp][6] = __uint_as_float(amu); }
  __syncthreads();
  if (warp == 0) {
    float l[3], h[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { l[k] = warp_min_f(red[lane][k]); h[k] = warp_max_f(red[lane][3 + k]); }
    const unsigned a = __reduce_max_sync(0xffffffffu, __float_as_uint(red[lane][6]));
    if (lane == 0) { for (int k = 0; k < 3; ++k) { red[0][k] = l[k]; red[0][3 + k] = h[k]; } red[0][6] = __uint_as_float(a); }
  }
  __syncthreads();
}

// Morton cell of a point: 16 cells per axis of the cloud's box (q0 = lower corner, qs = 15.999 / extent), 12 bits
constexpr int kCells = 4096;
constexpr int kCellCap = 64;              // deterministic mode: cells up to this size are insertion-sorted, larger ones heap-sorted
__device__ __forceinline__ unsigned prep_cell(float x, float y, float z, const float (&q0)[3], const float (&qs)[3]) {
  const float fx = (x - q0[0]) * qs[0], fy = (y - q0[1]) * qs[1], fz = (z - q0[2]) * qs[2];
  const unsigned qx = (fx >= 0.f) ? (unsigned)fminf(fx, 15.f) : 0u, qy = (fy >= 0.f) ? (unsigned)fminf(fy, 15.f) : 0u,
                 qz = (fz >= 0.f) ? (unsigned)fminf(fz, 15.f) : 0u;                                   // NaN -> 0
  return spread6(qx) | (spread6(qy) << 1) | (spread6(qz) << 2);
}

// Counting sort of the CTA's n items (item i = tid + 1024 k, k < K, belongs to this thread; its cell is half k & 1 of
// pk[k >> 1]) into ent[] (sorted position -> item), ONE pass over the 4096 cells: count with shared-memory reductions,
// prefix-sum the cells, then every item takes the next free position of its cell (atomic cursor).  The order INSIDE a
// cell is the order the atomics were served in; with `deterministic` set (a test / debugging knob) one thread per cell
// then puts the cell's items in index order (insertion sort up to kCellCap items, heapsort for a crowded cell), so that
// the permutation is reproducible - which sorted position an item gets never changes a result (the recovery kernels key
// on original indices), only which block of 128 a few border points fall into.  The whole CTA (1024 threads) calls it.
// (History: a bitonic network, 62 us for 8192 keys on one SM; 6-bit radix passes ranked with ballots, 15 us per pass;
// 4-bit radix passes with shuffle scans, 10 us per pass - an 18-bit code needed 3 to 5 such passes.  With 128 points per
// block the order inside a cell of 1/16 of the box does not matter, so 12 code bits and one pass are enough.)
template <int K>
__device__ __forceinline__ void prep_cell_sort(const unsigned (&pk)[(K + 1) / 2], int n, unsigned* cnt, unsigned* start, unsigned* ent,
                                               unsigned* wsum, int deterministic) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  reinterpret_cast<uint4*>(cnt)[tid] = make_uint4(0u, 0u, 0u, 0u);        // kCells == 4 x 1024
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (tid + k * kSortThreads < n) atomicAdd(&cnt[(pk[k >> 1] >> ((k & 1) * 16)) & 0xffffu], 1u);
  }
  __syncthreads();
  {
    const uint4 v = reinterpret_cast<const uint4*>(cnt)[tid];
    const unsigned sum = v.x + v.y + v.z + v.w;
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = wsum[lane];
      unsigned winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
      wsum[lane] = winc - w;
    }
    __syncthreads();
    const unsigned base = wsum[warp] + inc - sum;
    reinterpret_cast<uint4*>(start)[tid] = make_uint4(base, base + v.x, base + v.x + v.y, base + v.x + v.y + v.z);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i = tid + k * kSortThreads;
    if (i < n) ent[atomicAdd(&start[(pk[k >> 1] >> ((k & 1) * 16)) & 0xffffu], 1u)] = (unsigned)i;
  }
  __syncthreads();
  if (deterministic) {                                             // start[c] is now the END of cell c
#pragma unroll 1
    for (int c = 4 * tid; c < 4 * tid + 4; ++c) {
      const unsigned sz = cnt[c];
      if (sz < 2u) continue;
      unsigned* e = ent + (start[c] - sz);
      if (sz <= (unsigned)kCellCap) {
        for (unsigned a = 1; a < sz; ++a) {                        // insertion sort by item index
          const unsigned v = e[a];
          unsigned j = a;
          while (j > 0u && e[j - 1] > v) { e[j] = e[j - 1]; --j; }
          e[j] = v;
        }
      } else {                                                     // a crowded cell (coincident points): heapsort, O(n log n)
        auto sift = [&](unsigned root, unsigned end) {
          for (;;) {
            unsigned child = 2u * root + 1u;
            if (child >= end) break;
            if (child + 1u < end && e[child] < e[child + 1u]) ++child;
            if (e[root] >= e[child]) break;
            const unsigned t = e[root]; e[root] = e[child]; e[child] = t;
            root = child;
          }
        };
        for (unsigned st = sz / 2u; st-- > 0u;) sift(st, sz);
        for (unsigned end = sz - 1u; end > 0u; --end) {
          const unsigned t = e[0]; e[0] = e[end]; e[end] = t;
          sift(0u, end);
        }
      }
    }
    __syncthreads();
  }
}

// probe (NULL unless vpn_set_tuning("prep_probe", 1)): clock64 deltas of the phases of CTA 0 (target sort, statistics
// words 2-6) and CTA B (first row segment, words 7-11)
struct PrepProbe {
  unsigned long long* out; long long t; int slot;
  __device__ __forceinline__ void begin(unsigned long long* probe, bool mine, int first) {
    out = (probe != nullptr && mine && threadIdx.x == 0) ? probe : nullptr; slot = first; t = out ? clock64() : 0;
  }
  __device__ __forceinline__ void mark() {
    if (out) { const long long n = clock64(); atomicAdd(&out[slot++], (unsigned long long)(n - t)); t = n; }
  }
};

// sort of one sample's targets: the sorted copies p2s / p2v, the permutation perm, the chunk boxes, tmax
template <int K>
__device__ __forceinline__ void prep_sort_targets(const float* __restrict__ T, float* __restrict__ Ts, float4* __restrict__ Tv,
                                                  int* __restrict__ pm, float* __restrict__ cb, float* __restrict__ tmax_b,
                                                  int M, int do_sort, int deterministic, int nchunks, unsigned* cnt, unsigned* start,
                                                  unsigned* ent, unsigned* wsum, float (*red)[7], PrepProbe& pp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- bounding box (finite values only: fminf / fmaxf drop NaN) and the NaN-sticky largest |coordinate| (as bits: a NaN
  // compares above every number)
  float lo[3] = {prep_inf(), prep_inf(), prep_inf()}, hi[3] = {-prep_inf(), -prep_inf(), -prep_inf()};
  unsigned amu = 0u;
  for (int i0 = tid; i0 < M; i0 += 4 * kSortThreads) {              // four items per round: twelve loads in flight
    float v[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kSortThreads;
#pragma unroll
      for (int k = 0; k < 3; ++k) v[u][k] = (i < M) ? T[3 * (size_t)i + k] : __int_as_float(0x7fc00000);     // NaN: dropped below
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * kSortThreads < M) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          lo[k] = fminf(lo[k], v[u][k]); hi[k] = fmaxf(hi[k], v[u][k]);
          amu = max(amu, __float_as_uint(fabsf(v[u][k])));
        }
      }
    }
  }
  prep_block_box(lo, hi, amu, red);
  if (tid == 0) *tmax_b = red[0][6];
  pp.mark();                                                        // target: box
  if (do_sort) {
    float q0[3], qs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q0[k] = red[0][k];
      const float ext = red[0][3 + k] - red[0][k];
      qs[k] = (ext > 0.f && ext < 1e30f) ? 15.999f / ext : 0.f;
    }
    unsigned pk[K / 2];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int i = tid + k * kSortThreads;
      const unsigned c = (i < M) ? prep_cell(T[3 * (size_t)i], T[3 * (size_t)i + 1], T[3 * (size_t)i + 2], q0, qs) : 0u;
      if (k & 1) pk[k >> 1] |= c << 16; else pk[k >> 1] = c;
    }
    pp.mark();                                                      // target: cells
    prep_cell_sort<K>(pk, M, cnt, start, ent, wsum, deterministic);
    pp.mark();                                                      // target: sort
  } else { pp.mark(); pp.mark(); }
  for (int i0 = tid; i0 < nchunks * kBlk; i0 += 4 * kSortThreads) {       // four items per round: the gathers in flight together
    int src[4]; float x[4], y[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * kSortThreads; src[u] = (i < M) ? (do_sort ? (int)ent[i] : i) : -1; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x[u] = y[u] = z[u] = __int_as_float(0x7fc00000);
      if (src[u] >= 0) { x[u] = T[3 * (size_t)src[u]]; y[u] = T[3 * (size_t)src[u] + 1]; z[u] = T[3 * (size_t)src[u] + 2]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kSortThreads;
      if (i < nchunks * kBlk) {
        if (src[u] >= 0) { pm[i] = src[u]; Ts[3 * (size_t)i] = x[u]; Ts[3 * (size_t)i + 1] = y[u]; Ts[3 * (size_t)i + 2] = z[u]; }
        // what the row recovery stages: one 16-byte copy per target (NaN padding up to whole chunks)
        Tv[i] = make_float4(x[u], y[u], z[u], __int_as_float(src[u] >= 0 ? src[u] : 0x7fffffff));
      }
    }
  }
  __syncthreads();                                                 // the CTA's own global writes are visible to it
  pp.mark();                                                        // target: sorted copies
  // ---- chunk boxes: one warp per chunk, 4 columns per lane
  for (int c = warp; c < nchunks; c += kSortThreads / 32) {
    float l[3] = {prep_inf(), prep_inf(), prep_inf()}, h[3] = {-prep_inf(), -prep_inf(), -prep_inf()};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int col = c * kBlk + u * 32 + lane;
      if (col < M) {
        const float4 q = Tv[col];
        l[0] = fminf(l[0], q.x); l[1] = fminf(l[1], q.y); l[2] = fminf(l[2], q.z);
        h[0] = fmaxf(h[0], q.x); h[1] = fmaxf(h[1], q.y); h[2] = fmaxf(h[2], q.z);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { l[k] = warp_min_f(l[k]); h[k] = warp_max_f(h[k]); }
    if (lane == 0) {
      float* o = cb + (size_t)c * 8;
      o[0] = l[0]; o[1] = l[1]; o[2] = l[2]; o[3] = h[0]; o[4] = h[1]; o[5] = h[2]; o[6] = 0.f; o[7] = 0.f;
    }
  }
  __syncthreads();
  pp.mark();                                                        // target: chunk boxes
}

// Box layout: 8 floats per block / chunk: lo.x lo.y lo.z hi.x hi.y hi.z pad pad.  An empty box is (+inf, -inf).
// grid: x = sample [0, B), then (sample, row segment) [B, B + B nseg).  dynamic smem: u32 cnt[4096], start[4096], ent[max(M,
// 4096)], then (row segments) float sx / sy / sz[4096].  sort_targets = 0 when the sample is too large to sort: identity order.
constexpr int kRowSeg = 4096;             // rows per independently sorted segment
constexpr int kRowK = kRowSeg / kSortThreads;
__global__ void __launch_bounds__(kSortThreads)
chamfer_sort_targets_kernel(const float* __restrict__ p2, float* __restrict__ p2s, float4* __restrict__ p2v, int* __restrict__ perm,
                            float* __restrict__ cbox, float* __restrict__ tmax, int M, int sort_targets, int nchunks,
                            const float* __restrict__ p1, float* __restrict__ p1s, int* __restrict__ rperm,
                            float* __restrict__ rbox, int P, int nrb, int B, int deterministic,
                            unsigned long long* __restrict__ probe) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  unsigned* cnt = reinterpret_cast<unsigned*>(prep_smem);
  unsigned* start = cnt + kCells;
  unsigned* ent = start + kCells;
  __shared__ float red[32][7];
  __shared__ unsigned wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PrepProbe pp;
  pp.begin(probe, (int)blockIdx.x == 0 || (int)blockIdx.x == B, (int)blockIdx.x == 0 ? 2 : 7);
  if ((int)blockIdx.x >= B) {
    // ---- CTAs past the first B: one segment of kRowSeg predicted points each - Morton sort inside the segment's box, the
    // sorted copy, the permutation and the boxes of its 128-row blocks.  Independent of the target sort; sharing its
    // launch lets the two run side by side.
    const int nseg = (P + kRowSeg - 1) / kRowSeg;
    const int sid = (int)blockIdx.x - B, b = sid / nseg, s0 = (sid % nseg) * kRowSeg;
    const int n = min(kRowSeg, P - s0);
    const float* A = p1 + ((size_t)b * P + s0) * 3;
    // the segment's points are staged in shared memory: the gathers of the sorted copy and of the boxes read them there
    float* sx = reinterpret_cast<float*>(ent + kRowSeg); float* sy = sx + kRowSeg; float* sz = sy + kRowSeg;
    float x[kRowK], y[kRowK], z[kRowK];
    float lo[3] = {prep_inf(), prep_inf(), prep_inf()}, hi[3] = {-prep_inf(), -prep_inf(), -prep_inf()};
#pragma unroll
    for (int k = 0; k < kRowK; ++k) {
      const int i = tid + k * kSortThreads;
      x[k] = y[k] = z[k] = 0.f;
      if (i < n) { x[k] = A[3 * (size_t)i]; y[k] = A[3 * (size_t)i + 1]; z[k] = A[3 * (size_t)i + 2]; }
    }
#pragma unroll
    for (int k = 0; k < kRowK; ++k) {
      const int i = tid + k * kSortThreads;
      if (i < n) {
        sx[i] = x[k]; sy[i] = y[k]; sz[i] = z[k];
        lo[0] = fminf(lo[0], x[k]); lo[1] = fminf(lo[1], y[k]); lo[2] = fminf(lo[2], z[k]);           // NaN dropped
        hi[0] = fmaxf(hi[0], x[k]); hi[1] = fmaxf(hi[1], y[k]); hi[2] = fmaxf(hi[2], z[k]);
      }
    }
    prep_block_box(lo, hi, 0u, red);
    pp.mark();                                                      // row: staged + box
    float q0[3], qs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q0[k] = red[0][k];
      const float ext = red[0][3 + k] - red[0][k];
      qs[k] = (ext > 0.f && ext < 1e30f) ? 15.999f / ext : 0.f;
    }
    unsigned pk[kRowK / 2];
#pragma unroll
    for (int k = 0; k < kRowK; ++k) {
      const unsigned c = (tid + k * kSortThreads < n) ? prep_cell(x[k], y[k], z[k], q0, qs) : 0u;
      if (k & 1) pk[k >> 1] |= c << 16; else pk[k >> 1] = c;
    }
    pp.mark();                                                      // row: cells
    prep_cell_sort<kRowK>(pk, n, cnt, start, ent, wsum, deterministic);
    pp.mark();                                                      // row: sort
    // boxes of the segment's 128-row blocks in both orders: warp w = block w, 4 rows per lane.  The order whose blocks are
    // more compact (sum of squared box diagonals) is kept: surface samples of a primitive gain a lot from the sort, mesh
    // vertices that already arrive one small primitive per block (train_gcn.py) would lose.
    float bl[2][3], bh[2][3], diag[2] = {0.f, 0.f};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { bl[v][k] = prep_inf(); bh[v][k] = -prep_inf(); }
    }
    if (warp * kBlk < n) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = warp * kBlk + u * 32 + lane;
        if (i < n) {
          const int src = (int)ent[i];
          bl[0][0] = fminf(bl[0][0], sx[i]); bl[0][1] = fminf(bl[0][1], sy[i]); bl[0][2] = fminf(bl[0][2], sz[i]);
          bh[0][0] = fmaxf(bh[0][0], sx[i]); bh[0][1] = fmaxf(bh[0][1], sy[i]); bh[0][2] = fmaxf(bh[0][2], sz[i]);
          bl[1][0] = fminf(bl[1][0], sx[src]); bl[1][1] = fminf(bl[1][1], sy[src]); bl[1][2] = fminf(bl[1][2], sz[src]);
          bh[1][0] = fmaxf(bh[1][0], sx[src]); bh[1][1] = fmaxf(bh[1][1], sy[src]); bh[1][2] = fmaxf(bh[1][2], sz[src]);
        }
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          bl[v][k] = warp_min_f(bl[v][k]); bh[v][k] = warp_max_f(bh[v][k]);
          const float e = bh[v][k] - bl[v][k];
          if (e > 0.f) diag[v] = fmaf(e, e, diag[v]);
        }
      }
    }
    if (lane == 0) { red[warp][0] = diag[0]; red[warp][1] = diag[1]; }       // red[0][0..5] were consumed before the sort's barriers
    __syncthreads();
    float dn = 0.f, ds = 0.f;
    if (warp == 0) {
      dn = red[lane][0]; ds = red[lane][1];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { dn += __shfl_xor_sync(0xffffffffu, dn, o); ds += __shfl_xor_sync(0xffffffffu, ds, o); }
      if (lane == 0) wsum[0] = (ds < dn) ? 1u : 0u;                          // NaN sums: natural order
    }
    __syncthreads();
    const bool use_sorted = wsum[0] != 0u;                                   // CTA-uniform
    pp.mark();                                                      // row: block boxes, both orders
    float* As = p1s + ((size_t)b * P + s0) * 3;
    int* rp = rperm + (size_t)b * P + s0;
#pragma unroll
    for (int k = 0; k < kRowK; ++k) {
      const int i = tid + k * kSortThreads;
      if (i < n) rp[i] = s0 + (use_sorted ? (int)ent[i] : i);
    }
#pragma unroll
    for (int k = 0; k < 3 * kRowK; ++k) {                            // consecutive threads write consecutive floats
      const int j = tid + k * kSortThreads;
      if (j < 3 * n) {
        const int r = j / 3, c = j - 3 * r;
        const int src = use_sorted ? (int)ent[r] : r;
        As[j] = (c == 0 ? sx : (c == 1 ? sy : sz))[src];
      }
    }
    if (warp * kBlk < n && lane == 0) {
      float* o = rbox + ((size_t)b * nrb + s0 / kBlk + warp) * 8;
#pragma unroll
      for (int k = 0; k < 3; ++k) { o[k] = use_sorted ? bl[1][k] : bl[0][k]; o[3 + k] = use_sorted ? bh[1][k] : bh[0][k]; }
      o[6] = 0.f; o[7] = 0.f;
    }
    __syncthreads();
    pp.mark();                                                      // row: sorted copy, permutation, boxes
    return;
  }
  const int b = blockIdx.x;
  const float* T = p2 + (size_t)b * M * 3;
  float* Ts = p2s + (size_t)b * M * 3;
  float4* Tv = p2v + (size_t)b * nchunks * kBlk;
  int* pm = perm + (size_t)b * M;
  float* cb = cbox + (size_t)b * nchunks * 8;
  if (M <= 8 * kSortThreads) prep_sort_targets<8>(T, Ts, Tv, pm, cb, tmax + b, M, sort_targets, deterministic, nchunks, cnt, start, ent, wsum, red, pp);
  else prep_sort_targets<16>(T, Ts, Tv, pm, cb, tmax + b, M, sort_targets, deterministic, nchunks, cnt, start, ent, wsum, red, pp);
}

// ---- upper bounds of the nearest-neighbour distances ---------------------------------------------------------------
// For a block of 128 points of one cloud: pick the kNear blocks of the OTHER cloud whose boxes are closest to this
// block's box, take kReps points of each as representatives, and give every point of the block the smallest exact
// distance to a representative - an upper bound of its nearest-neighbour distance, tight when the neighbour lies in
// one of the near blocks (the usual case) and still valid when it does not.  The block's bound is the max over its points.
constexpr int kNearRows = 8, kRepsRows = 4;      // a row block looks at 8 chunks x 4 columns
constexpr int kNearCols = 32, kRepsCols = 4;     // a chunk looks at 32 row blocks x 4 rows (its nearest row is one of tens of thousands:
                                                 // 16 x 2 left 46 % of the column-direction blocks live on the C2 clouds, 32 x 4 leaves 42 %; exact: 25-30 %)
constexpr int kMaxNear = 32, kMaxReps = 128;
constexpr int kGapCap = 1024;                    // boxes of the other cloud considered per block (strided subset beyond that)
constexpr int kBoundWarps = 4;                   // blocks of 128 points per CTA: one per warp, no block-level barrier

__device__ __forceinline__ float box_gap2(const float* __restrict__ a, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float g = fmaxf(0.f, fmaxf(a[k] - b[3 + k], b[k] - a[3 + k]));
    s = fmaf(g, g, s);
  }
  return s;
}

// grid: x = groups of kBoundWarps blocks (chunks first, then row blocks), y = sample; one WARP per block of 128 points,
// 4 points per lane (the first version gave a block a whole CTA: 58 % of its stall samples were the three other warps
// waiting at barriers for warp 0's selection rounds).
//   blocks [0, nchunks)             chunk c      : cub[b][c]   = max_j min_rep d2(col j, rep)
//   blocks [nchunks, nchunks + nrb) row block rb : rthr[b][rb] = max_i min_rep d2(row i, rep)
__global__ void __launch_bounds__(kBoundWarps * 32, 8)
chamfer_prune_bounds_kernel(const float* __restrict__ p1, const float* __restrict__ p2s,
                            const float* __restrict__ rbox, const float* __restrict__ cbox,
                            float* __restrict__ rthr, float* __restrict__ cub, int P, int M, int nrb, int nchunks,
                            int near_rows, int reps_rows, int near_cols, int reps_cols, int gap_cap) {
  // (quantised gap bits | candidate index), 0xffffffff = taken: gap_cap entries per warp, dynamic - sized by the larger
  // block count of the two clouds (<= kGapCap), so that the C2 shape keeps 8 CTAs per SM instead of 6
  extern __shared__ unsigned s_gap[];
  __shared__ float4 s_reps[kBoundWarps][2 * kMaxReps];   // per representative (x, x, y, y) (z, z, -, -): operands of the packed f32x2 ops
  __shared__ int s_sel[kBoundWarps][kMaxNear];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int id = blockIdx.x * kBoundWarps + warp;
  if (id >= nrb + nchunks) return;                                           // warp-uniform
  unsigned* gap = s_gap + warp * gap_cap; float4* reps = s_reps[warp]; int* sel = s_sel[warp];
  const bool is_row = id >= nchunks;                                         // the chunks (4 x the work of a row block) are dispatched first
  const int blk = is_row ? id - nchunks : id;
  const int n_mine = is_row ? P : M, n_other = is_row ? M : P;
  const float* mine = (is_row ? p1 + (size_t)b * P * 3 : p2s + (size_t)b * M * 3);
  const float* other = (is_row ? p2s + (size_t)b * M * 3 : p1 + (size_t)b * P * 3);
  const int nob = is_row ? nchunks : nrb;                                    // blocks of the other cloud
  const float* obox = (is_row ? cbox + (size_t)b * nchunks * 8 : rbox + (size_t)b * nrb * 8);
  const int near = is_row ? near_rows : near_cols, per = is_row ? reps_rows : reps_cols;
  const float* mb = is_row ? rbox + ((size_t)b * nrb + blk) * 8 : cbox + ((size_t)b * nchunks + blk) * 8;
  float mybox[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) mybox[k] = mb[k];
  float x[4], y[4], z[4]; bool valid[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = blk * kBlk + u * 32 + lane;
    valid[u] = idx < n_mine;
    x[u] = y[u] = z[u] = 0.f;
    if (valid[u]) { x[u] = mine[3 * (size_t)idx]; y[u] = mine[3 * (size_t)idx + 1]; z[u] = mine[3 * (size_t)idx + 2]; }
  }
  const int stride = (nob + kGapCap - 1) / kGapCap, ncand = (nob + stride - 1) / stride;
  // key = gap^2 with its low 10 mantissa bits replaced by the candidate index: unsigned order = (gap to ~2^-13 relative,
  // index).  Which near boxes are chosen only affects how tight the bound is, never its validity.
  unsigned lmin = 0xffffffffu;                                               // smallest key among this lane's candidates
  for (int c = lane; c < ncand; c += 32) {
    const float4* ob = reinterpret_cast<const float4*>(obox + (size_t)c * stride * 8);      // 32-byte box records: two 16-byte loads
    const float4 o0 = ob[0], o1 = ob[1];
    const float other_box[6] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y};
    const float g = box_gap2(mybox, other_box);
    const unsigned k = ((g == g) ? (__float_as_uint(g) & ~0x3ffu) : 0x7f800000u) | (unsigned)c;   // NaN boxes sort last
    gap[c] = k;
    lmin = min(lmin, k);
  }
  // `near` rounds of arg-min with removal: one warp reduction per round; only the winner's lane rescans its candidates
  const int nsel = min(near, ncand);
  for (int s = 0; s < nsel; ++s) {
    const unsigned w = __reduce_min_sync(0xffffffffu, lmin);
    const int bi = (w == 0xffffffffu) ? -1 : (int)(w & 0x3ffu);
    if (lane == 0) sel[s] = bi < 0 ? 0 : bi;
    if (bi >= 0 && (bi & 31) == lane) {
      gap[bi] = 0xffffffffu;
      lmin = 0xffffffffu;
      for (int c = lane; c < ncand; c += 32) lmin = min(lmin, gap[c]);
    }
  }
  __syncwarp();
  const int nrep = nsel * per;
  for (int i = lane; i < nrep; i += 32) {
    const int ob = sel[i / per] * stride;
    int o = ob * kBlk + (i % per) * (kBlk / per);
    o = min(o, n_other - 1);                                                 // a clamped duplicate is still a real point of the cloud
    const float qx = other[3 * (size_t)o], qy = other[3 * (size_t)o + 1], qz = other[3 * (size_t)o + 2];
    reps[2 * i] = make_float4(qx, qx, qy, qy); reps[2 * i + 1] = make_float4(qz, qz, 0.f, 0.f);
  }
  __syncwarp();
  // Two points per packed op (sub / mul / fma .f32x2: 6 ops per two distances where the scalar form needs 16).  The fused
  // multiply-adds round differently from the reference's separate operations, by ~1e-7 relative: the bound only has to
  // hold within the 1e-5 margin the plan kernel applies.
  const u64 x01 = prep_pk(x[0], x[1]), x23 = prep_pk(x[2], x[3]), y01 = prep_pk(y[0], y[1]), y23 = prep_pk(y[2], y[3]),
            z01 = prep_pk(z[0], z[1]), z23 = prep_pk(z[2], z[3]);
  float ub[4] = {prep_inf(), prep_inf(), prep_inf(), prep_inf()};
  for (int r = 0; r < nrep; ++r) {
    const ulonglong2 qa = *reinterpret_cast<const ulonglong2*>(&reps[2 * r]);       // (x, x), (y, y)
    const u64 qz = *reinterpret_cast<const u64*>(&reps[2 * r + 1]);                 // (z, z)
    const u64 dx0 = prep_sub2(x01, qa.x), dy0 = prep_sub2(y01, qa.y), dz0 = prep_sub2(z01, qz);
    const u64 dx1 = prep_sub2(x23, qa.x), dy1 = prep_sub2(y23, qa.y), dz1 = prep_sub2(z23, qz);
    const u64 s0 = prep_fma2(dz0, dz0, prep_fma2(dy0, dy0, prep_mul2(dx0, dx0)));
    const u64 s1 = prep_fma2(dz1, dz1, prep_fma2(dy1, dy1, prep_mul2(dx1, dx1)));
    float a, c, e, f;
    prep_upk(s0, a, c); prep_upk(s1, e, f);
    ub[0] = fminf(ub[0], a); ub[1] = fminf(ub[1], c); ub[2] = fminf(ub[2], e); ub[3] = fminf(ub[3], f);   // NaN distances are dropped: +inf -> nothing pruned
  }
  float m = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) if (valid[u]) m = fmaxf(m, ub[u]);
  m = warp_max_f(m);
  if (lane == 0) { if (is_row) rthr[(size_t)b * nrb + blk] = m; else cub[(size_t)b * nchunks + blk] = m; }
}

int chamfer_row_segment() { return kRowSeg; }

size_t chamfer_sort_smem_bytes(int M) {
  if (M > kSortMaxM) return 0;
  return (size_t)kCells * 8 + (((size_t)M * 4 + 15) & ~(size_t)15);         // cnt, start, ent[M]
}

// p1s / rperm: the predicted points sorted inside segments of kRowSeg rows and the map sorted row -> original row; the
// bounds, and every later kernel of the forward, work on the sorted copy.
int chamfer_prep_launch(const float* p1, const float* p2, float* p2s, float4* p2v, int* perm, float* cbox, float* rbox, float* rthr,
                        float* cub, float* tmax, float* p1s, int* rperm, unsigned long long* stats, int B, int P, int M, cudaStream_t s) {
  const int nchunks = (M + kBlk - 1) / kBlk, nrb = (P + kBlk - 1) / kBlk;
  const size_t smem_t = chamfer_sort_smem_bytes(M);
  const size_t smem_r = (size_t)kCells * 8 + (size_t)kRowSeg * 4 + (size_t)kRowSeg * 12;      // cnt, start, ent + the staged points
  const size_t smem = smem_t > smem_r ? smem_t : smem_r;
  static DeviceOnce once;
  if (set_dyn_smem(chamfer_sort_targets_kernel, kCells * 8 + kSortMaxM * 4 + kRowSeg * 12, once) != cudaSuccess) {
    vpn_set_error("chamfer prep: smem attribute"); return VPN_ERR_CUDA;
  }
  const long long seg_ctas = (long long)((P + kRowSeg - 1) / kRowSeg) * B;
  if (B + seg_ctas > 0x7fffffffLL) { vpn_set_error("chamfer prep: too many row segments"); return VPN_ERR_SHAPE; }
  chamfer_sort_targets_kernel<<<(unsigned)(B + seg_ctas), kSortThreads, smem, s>>>(p2, p2s, p2v, perm, cbox, tmax, M, smem_t > 0 ? 1 : 0, nchunks,
                                                                                   p1, p1s, rperm, rbox, P, nrb, B,
                                                                                   tuning_value(kTunePrepDeterministic) == 1 ? 1 : 0,
                                                                                   tuning_value(kTunePrepProbe) == 1 ? stats : nullptr);
  int rc = vpn_check_launch("chamfer_sort_targets_kernel");
  if (rc) return rc;
  // vpn_set_tuning("prep_near_rows" / "prep_reps_rows" / "prep_near_cols" / "prep_reps_cols"): probes only (near <= 32, reps 1 / 2 / 4)
  auto pick = [](int key, int dflt, int cap) { const int v = tuning_value(key); return (v >= 1 && v <= cap && (key == kTunePrepNearRows || key == kTunePrepNearCols || v == 1 || v == 2 || v == 4)) ? v : dflt; };
  const int most = nrb > nchunks ? nrb : nchunks;
  const int gap_cap = most >= kGapCap ? kGapCap : ((most + 31) & ~31);      // candidates per block are min(blocks of the other cloud, kGapCap)
  chamfer_prune_bounds_kernel<<<dim3((nrb + nchunks + kBoundWarps - 1) / kBoundWarps, B), kBoundWarps * 32,
                                (size_t)kBoundWarps * gap_cap * sizeof(unsigned), s>>>(
      p1s, p2s, rbox, cbox, rthr, cub, P, M, nrb, nchunks, pick(kTunePrepNearRows, kNearRows, kMaxNear), pick(kTunePrepRepsRows, kRepsRows, 4),
      pick(kTunePrepNearCols, kNearCols, kMaxNear), pick(kTunePrepRepsCols, kRepsCols, 4), gap_cap);
  return vpn_check_launch("chamfer_prune_bounds_kernel");
}

}  // namespace vpn
