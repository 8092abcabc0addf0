// Spatial preparation for the tensor-core Chamfer filter (chamfer_tc.cu): what lets it SKIP distance blocks.
//
// The filter evaluates the (P x M) distance matrix of modules/loss/chamfer_distance.py:14-23 in stages of 128 rows x
// 256 columns.  A stage cannot hold any row's (column's) nearest neighbour when the two point sets are further apart
// than a distance some point of the block is already known to achieve.  Two small kernels provide the ingredients:
//
//   chamfer_sort_targets_kernel   one CTA per sample: Morton-sorts the sample's targets (18-bit code, 64 cells per axis
//       of the sample's bounding box, index in the low 14 key bits: a deterministic permutation), writes the sorted copy p2s, the
//       permutation perm (sorted position -> original index), the axis-aligned box of every 128-column chunk, and the
//       largest |coordinate| (the filter's scale).  After sorting a chunk is a compact patch of the target shape.
//       Predicted points need no sorting: they arrive primitive-major (train.py:119 torch.cat(dim=1)), so 128
//       consecutive rows are one patch of one primitive.
//       The CTAs past the first B of the same launch compute the box of every 128-row block of the predicted cloud.
//   chamfer_prune_bounds_kernel   per 128-row block: T_r = max over its rows of an UPPER bound of the row's
//       nearest-target distance (exact distance, the reference's arithmetic, to 4 representatives of each of the 8 chunks
//       whose boxes are nearest to the block's box); per 128-column chunk: U_c likewise over 16 row blocks x 2 rows.
//
// chamfer_tc_kernel skips the stage (row block r, chunk c) for the row direction when gap(box_r, box_c)^2 > T_r and for
// the column direction when gap^2 > U_c (both with a 1e-5 relative margin): every pair in the stage is then further
// apart than a distance row i (column j) certainly achieves elsewhere, so neither its arg-min nor a tie can be there.
// Results are unchanged bit for bit (the exact recovery kernels still decide); first-index ties survive the
// permutation because recovery keys carry the ORIGINAL target index.
#include "common.cuh"

namespace vpn {

constexpr int kSortThreads = 1024;
constexpr int kSortMaxM = 16384;          // 14 index bits in the 32-bit sort key; two key buffers of 64 KB in shared memory
constexpr int kSortMinKeys = 1024;        // key count granularity: 32 warps x 32 lanes
constexpr int kBlk = 128;                 // rows per block = columns per chunk

__device__ __forceinline__ float prep_inf() { return __int_as_float(0x7f800000); }
__device__ __forceinline__ unsigned spread6(unsigned x) {       // 6 bits -> every third bit
  x = (x | (x << 8)) & 0x0000300Fu;
  x = (x | (x << 4)) & 0x000030C3u;
  x = (x | (x << 2)) & 0x00009249u;
  return x;
}
// lanes of the warp whose 6-bit digit equals this lane's (match.any costs hundreds of cycles on sm_100; six ballots do not)
__device__ __forceinline__ unsigned peers6(unsigned d) {
  unsigned m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
    m &= ((d >> b) & 1u) ? bal : ~bal;
  }
  return m;
}
__device__ __forceinline__ float prep_d2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Box layout: 8 floats per block / chunk: lo.x lo.y lo.z hi.x hi.y hi.z pad pad.  An empty box is (+inf, -inf).
// grid: x = sample.  dynamic smem: u32 keys[2][npad], npad = M rounded up to 1024 (npad = 0 when the sample is too large to
// sort: identity order).
__global__ void __launch_bounds__(kSortThreads)
chamfer_sort_targets_kernel(const float* __restrict__ p2, float* __restrict__ p2s, float4* __restrict__ p2v, int* __restrict__ perm,
                            float* __restrict__ cbox, float* __restrict__ tmax, int M, int npad, int nchunks,
                            const float* __restrict__ p1, float* __restrict__ rbox, int P, int nrb, int B) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  if ((int)blockIdx.x >= B) {
    // ---- CTAs past the first B: boxes of the 128-row blocks of the predicted cloud, one warp per block, 4 rows per lane.
    // Independent of the sort; sharing its launch lets the two run side by side (the sort occupies B SMs for tens of
    // microseconds; as a kernel of its own the boxes waited for it in the stream).
    const int lane = threadIdx.x & 31;
    const long long blk = ((long long)blockIdx.x - B) * (kSortThreads / 32) + (threadIdx.x >> 5);
    if (blk >= (long long)nrb * B) return;
    const int b = (int)(blk / nrb), rb = (int)(blk % nrb);
    float l[3] = {prep_inf(), prep_inf(), prep_inf()}, h[3] = {-prep_inf(), -prep_inf(), -prep_inf()};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = rb * kBlk + u * 32 + lane;
      if (row < P) {
        const float* a = p1 + 3 * ((size_t)b * P + row);
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = a[k]; l[k] = fminf(l[k], v); h[k] = fmaxf(h[k], v); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o));
        h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o));
      }
    }
    if (lane == 0) {
      float* o = rbox + ((size_t)b * nrb + rb) * 8;
      o[0] = l[0]; o[1] = l[1]; o[2] = l[2]; o[3] = h[0]; o[4] = h[1]; o[5] = h[2]; o[6] = 0.f; o[7] = 0.f;
    }
    return;
  }
  unsigned* keys = reinterpret_cast<unsigned*>(prep_smem);
  __shared__ float red[32][7];
  __shared__ unsigned hist[64 * 33];
  __shared__ int wsum[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* T = p2 + (size_t)b * M * 3;
  float* Ts = p2s + (size_t)b * M * 3;
  int* pm = perm + (size_t)b * M;
  // ---- bounding box (finite values only: fminf / fmaxf drop NaN) and the NaN-sticky largest |coordinate|
  float lo[3] = {prep_inf(), prep_inf(), prep_inf()}, hi[3] = {-prep_inf(), -prep_inf(), -prep_inf()}, am = 0.f;
  for (int i = tid; i < M; i += kSortThreads) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float v = T[3 * (size_t)i + k];
      lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v);
      const float a = fabsf(v);
      am = (a <= am) ? am : a;                                    // NaN sticks
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    const float x = __shfl_xor_sync(0xffffffffu, am, o);
    am = (x <= am) ? am : x;
  }
  if (lane == 0) { for (int k = 0; k < 3; ++k) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; } red[warp][6] = am; }
  __syncthreads();
  if (warp == 0) {
    for (int k = 0; k < 3; ++k) { lo[k] = red[lane][k]; hi[k] = red[lane][3 + k]; }
    am = red[lane][6];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
        hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      }
      const float x = __shfl_xor_sync(0xffffffffu, am, o);
      am = (x <= am) ? am : x;
    }
    if (lane == 0) { for (int k = 0; k < 3; ++k) { red[0][k] = lo[k]; red[0][3 + k] = hi[k]; } tmax[b] = am; }
  }
  __syncthreads();
  const unsigned* sorted = keys;
  if (npad > 0) {
    float q0[3], qs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q0[k] = red[0][k];
      const float ext = red[0][3 + k] - red[0][k];
      qs[k] = (ext > 0.f && ext < 1e30f) ? 63.999f / ext : 0.f;
    }
    for (int i = tid; i < npad; i += kSortThreads) {
      unsigned key = 0xffffffffu;
      if (i < M) {
        unsigned code = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float f = (T[3 * (size_t)i + k] - q0[k]) * qs[k];
          const unsigned q = (f >= 0.f) ? (unsigned)fminf(f, 63.f) : 0u;            // NaN -> 0
          code |= spread6(q) << k;
        }
        key = (code << 14) | (unsigned)i;
      }
      keys[i] = key;
    }
    __syncthreads();
    // LSD radix sort of the 18 code bits, three stable passes of 6 bits.  Keys start in index order and every pass is
    // stable, so equal codes stay in index order: the same deterministic permutation as sorting the full 32-bit keys
    // (the bitonic network this replaces: 91 substeps, 62 us for 8192 keys on one SM).  Warp w owns the contiguous
    // run [w seg, (w + 1) seg) of the pass's input; a round ranks 32 keys with six ballots (rank = peers of the same
    // digit in lower lanes), hist[digit][warp] counted in a first walk and scanned digit-major gives every (digit,
    // warp) its output offset.  Padding keys (0xffffffff) have digit 63 in every pass and come last in the input: they
    // stay behind every real key.
    const int seg = npad >> 5;
    unsigned* kin = keys; unsigned* kout = keys + npad;
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = 14 + 6 * pass;
      for (int i = tid; i < 64 * 33; i += kSortThreads) hist[i] = 0u;
      __syncthreads();
      for (int r = 0; r < seg; r += 32) {
        const unsigned d = (kin[warp * seg + r + lane] >> shift) & 63u;
        const unsigned peers = peers6(d);
        if ((peers & ((1u << lane) - 1u)) == 0u) hist[d * 33 + warp] += (unsigned)__popc(peers);
        __syncwarp();
      }
      __syncthreads();
      {                                                              // exclusive scan over (digit, warp), two entries per thread
        const int l0 = 2 * tid, l1 = 2 * tid + 1;
        const int p0 = (l0 >> 5) * 33 + (l0 & 31), p1i = (l1 >> 5) * 33 + (l1 & 31);
        const unsigned a = hist[p0], c = hist[p1i];
        int inc = (int)(a + c);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
          const int w = wsum[lane];
          int winc = w;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
          wsum[lane] = winc - w;
        }
        __syncthreads();
        const unsigned off = (unsigned)(wsum[warp] + inc) - (a + c);
        hist[p0] = off; hist[p1i] = off + a;
      }
      __syncthreads();
      for (int r = 0; r < seg; r += 32) {
        const unsigned key = kin[warp * seg + r + lane];
        const unsigned d = (key >> shift) & 63u;
        const unsigned peers = peers6(d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const unsigned base = hist[d * 33 + warp];
        kout[base + rank] = key;
        __syncwarp();
        if (rank == 0) hist[d * 33 + warp] = base + (unsigned)__popc(peers);
        __syncwarp();
      }
      __syncthreads();
      unsigned* t = kin; kin = kout; kout = t;
    }
    sorted = kin;
  }
  float4* Tv = p2v + (size_t)b * nchunks * kBlk;
  for (int i = tid; i < nchunks * kBlk; i += kSortThreads) {
    if (i < M) {
      const int src = npad > 0 ? (int)(sorted[i] & 0x3fffu) : i;
      const float x = T[3 * (size_t)src], y = T[3 * (size_t)src + 1], z = T[3 * (size_t)src + 2];
      pm[i] = src;
      Ts[3 * (size_t)i] = x; Ts[3 * (size_t)i + 1] = y; Ts[3 * (size_t)i + 2] = z;
      Tv[i] = make_float4(x, y, z, __int_as_float(src));            // what the row recovery stages: one 16-byte copy per target
    } else {
      const float qnan = __int_as_float(0x7fc00000);
      Tv[i] = make_float4(qnan, qnan, qnan, __int_as_float(0x7fffffff));
    }
  }
  __syncthreads();                                                 // the CTA's own global writes are visible to it
  // ---- chunk boxes: one warp per chunk, 4 columns per lane
  for (int c = warp; c < nchunks; c += kSortThreads / 32) {
    float l[3] = {prep_inf(), prep_inf(), prep_inf()}, h[3] = {-prep_inf(), -prep_inf(), -prep_inf()};
    for (int u = 0; u < 4; ++u) {
      const int col = c * kBlk + u * 32 + lane;
      if (col < M) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = Ts[3 * (size_t)col + k]; l[k] = fminf(l[k], v); h[k] = fmaxf(h[k], v); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o));
        h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o));
      }
    }
    if (lane == 0) {
      float* o = cbox + ((size_t)b * nchunks + c) * 8;
      o[0] = l[0]; o[1] = l[1]; o[2] = l[2]; o[3] = h[0]; o[4] = h[1]; o[5] = h[2]; o[6] = 0.f; o[7] = 0.f;
    }
  }
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
__global__ void __launch_bounds__(kBoundWarps * 32)
chamfer_prune_bounds_kernel(const float* __restrict__ p1, const float* __restrict__ p2s,
                            const float* __restrict__ rbox, const float* __restrict__ cbox,
                            float* __restrict__ rthr, float* __restrict__ cub, int P, int M, int nrb, int nchunks) {
  __shared__ unsigned s_gap[kBoundWarps][kGapCap];      // (quantised gap bits | candidate index), 0xffffffff = taken
  __shared__ float4 s_reps[kBoundWarps][kMaxReps];
  __shared__ int s_sel[kBoundWarps][kMaxNear];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int id = blockIdx.x * kBoundWarps + warp;
  if (id >= nrb + nchunks) return;                                           // warp-uniform
  unsigned* gap = s_gap[warp]; float4* reps = s_reps[warp]; int* sel = s_sel[warp];
  const bool is_row = id >= nchunks;                                         // the chunks (4 x the work of a row block) are dispatched first
  const int blk = is_row ? id - nchunks : id;
  const int n_mine = is_row ? P : M, n_other = is_row ? M : P;
  const float* mine = (is_row ? p1 + (size_t)b * P * 3 : p2s + (size_t)b * M * 3);
  const float* other = (is_row ? p2s + (size_t)b * M * 3 : p1 + (size_t)b * P * 3);
  const int nob = is_row ? nchunks : nrb;                                    // blocks of the other cloud
  const float* obox = (is_row ? cbox + (size_t)b * nchunks * 8 : rbox + (size_t)b * nrb * 8);
  const int near = is_row ? kNearRows : kNearCols, per = is_row ? kRepsRows : kRepsCols;
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
    const float g = box_gap2(mybox, obox + (size_t)c * stride * 8);
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
    reps[i] = make_float4(other[3 * (size_t)o], other[3 * (size_t)o + 1], other[3 * (size_t)o + 2], 0.f);
  }
  __syncwarp();
  float ub[4] = {prep_inf(), prep_inf(), prep_inf(), prep_inf()};
  for (int r = 0; r < nrep; ++r) {
    const float4 q = reps[r];
#pragma unroll
    for (int u = 0; u < 4; ++u) ub[u] = fminf(ub[u], prep_d2(x[u], y[u], z[u], q.x, q.y, q.z));   // NaN distances are dropped: +inf -> nothing pruned
  }
  float m = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) if (valid[u]) m = fmaxf(m, ub[u]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) { if (is_row) rthr[(size_t)b * nrb + blk] = m; else cub[(size_t)b * nchunks + blk] = m; }
}

size_t chamfer_sort_smem_bytes(int M) {
  if (M > kSortMaxM) return 0;
  const int npad = (M + kSortMinKeys - 1) / kSortMinKeys * kSortMinKeys;     // whole rounds of 32 keys for each of the 32 warps
  return (size_t)npad * 8;
}

int chamfer_prep_launch(const float* p1, const float* p2, float* p2s, float4* p2v, int* perm, float* cbox, float* rbox, float* rthr,
                        float* cub, float* tmax, int B, int P, int M, cudaStream_t s) {
  const int nchunks = (M + kBlk - 1) / kBlk, nrb = (P + kBlk - 1) / kBlk;
  const size_t smem = chamfer_sort_smem_bytes(M);
  static DeviceOnce once;
  if (set_dyn_smem(chamfer_sort_targets_kernel, kSortMaxM * 8, once) != cudaSuccess) {
    vpn_set_error("chamfer prep: smem attribute"); return VPN_ERR_CUDA;
  }
  const long long box_ctas = ((long long)nrb * B + kSortThreads / 32 - 1) / (kSortThreads / 32);
  if (B + box_ctas > 0x7fffffffLL) { vpn_set_error("chamfer prep: too many row blocks"); return VPN_ERR_SHAPE; }
  chamfer_sort_targets_kernel<<<(unsigned)(B + box_ctas), kSortThreads, smem, s>>>(p2, p2s, p2v, perm, cbox, tmax, M, (int)(smem / 8), nchunks,
                                                                                   p1, rbox, P, nrb, B);
  int rc = vpn_check_launch("chamfer_sort_targets_kernel");
  if (rc) return rc;
  chamfer_prune_bounds_kernel<<<dim3((nrb + nchunks + kBoundWarps - 1) / kBoundWarps, B), kBoundWarps * 32, 0, s>>>(
      p1, p2s, rbox, cbox, rthr, cub, P, M, nrb, nchunks);
  return vpn_check_launch("chamfer_prune_bounds_kernel");
}

}  // namespace vpn
