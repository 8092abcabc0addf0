// Chamfer nearest-neighbour search, tiled FP32 path for sm_100a.
//
// One pass over the (P x M) distance matrix serves BOTH directions of
// modules/loss/chamfer_distance.py:14-23 (row minima over the targets, column minima over the
// predicted points), with O(P + M) memory instead of the reference's dense (B,P,M,3) temporary.
//
// Structure
//   main kernel   : CTA = 256 threads; each thread owns R rows of cloud 1 held as packed f32x2
//                   register pairs (FADD2/FMUL2/FFMA2, two pairs per issue slot); columns of cloud 2
//                   are staged through shared memory in chunks of kCW and broadcast to all lanes.
//                   The hot loop is branch free: it only tracks VALUES - per-row running minima
//                   (FMNMX3) and per-column warp minima (FMNMX3 + CREDUX.MIN.F32).  At the end of each
//                   chunk it folds them into a tiny candidate record: best value + bit mask of the
//                   chunks (rows) / warps (columns) that can still hold the arg-min.
//   recovery      : one thread per row / per column re-evaluates only the candidate chunks with the
//                   reference's exact arithmetic and picks (min sqrt(d), first index) - torch.min's
//                   tie rule - so the arg-min is bit-exact whatever arithmetic the hot loop used.
//
// Arithmetic modes of the hot loop (recovery is always exact):
//   MODE_EXACT : d = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)), the reference's own rounding sequence
//                (8 flop per pair on the FP32 pipe, no fusion) - candidate slack covers only the
//                sqrt rounding classes.
//   MODE_DIFF  : same differences, FMA-accumulated squares (6 pipe ops per pair).  Relative error
//                vs the reference <= ~6 ulp, absorbed by a 2^-18 relative candidate slack.
//
// NOTE on ptxas: mul.rn.f32x2 followed by add.rn.f32x2 IS contracted into FFMA2 by ptxas 12.9 even
// with --fmad=false (the scalar .rn forms are not).  The exact mode therefore performs its two
// additions as fma(x, 1.0, y) with the 1.0 passed as a kernel argument, which rounds exactly like
// an add and cannot be contracted.
#include "common.cuh"

namespace vpn {

enum { MODE_EXACT = 0, MODE_DIFF = 1 };

constexpr int kTThreads = 256;
constexpr int kTWarps = kTThreads / 32;
constexpr int kCW = 128;                  // columns per chunk (candidate granularity for rows)
constexpr int kMaxChunksPerSplit = 64;    // bits in the row candidate mask

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float warp_min_f32(float a) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(a)); return r; }

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

// Upper bound of the set of hot-loop values that may still belong to the arg-min when the smallest
// hot-loop value seen is x (sqrt rounding class of the reference + hot-loop arithmetic error).
template <int MODE>
__device__ __forceinline__ float thr_of(float x) {
  const float rel = (MODE == MODE_EXACT) ? 9.5367431640625e-07f /* 2^-20 */ : 3.814697265625e-06f /* 2^-18 */;
  return __fadd_ru(__fmaf_ru(x, rel, x), 1e-36f);
}

// Squared distances of one packed row pair against one broadcast column.
template <int MODE>
__device__ __forceinline__ u64 pair_d2(u64 px, u64 py, u64 pz, float cx, float cy, float cz, float one) {
  u64 dx = sub2(px, pk(cx, cx)), dy = sub2(py, pk(cy, cy)), dz = sub2(pz, pk(cz, cz));
  if (MODE == MODE_EXACT) {
    u64 xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
    u64 o2 = pk(one, one);
    return fma2(fma2(xx, o2, yy), o2, zz);          // fl(fl(xx + yy) + zz), never contracted
  } else {
    return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
  }
}

__device__ __forceinline__ float exact_d2s(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <int R>
struct TiledSmem {
  float4 tile[2][kCW];
  float colw[2][kTWarps][kCW];
  float rs_best[R][kTThreads];
  float rs_thr[R][kTThreads];
  u64 rs_mask[R][kTThreads];
};

// grid: x = row tile, y = column split, z = sample
template <int R, int MODE>
__global__ void __launch_bounds__(kTThreads, (R <= 8 ? 2 : 1))
chamfer_tiled_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                     float* __restrict__ rbest, u64* __restrict__ rmask,
                     float* __restrict__ cbest, unsigned* __restrict__ cmask,
                     int P, int M, int nchunks, int cps, float one) {
  constexpr int TM = kTThreads * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TiledSmem<R>& sm = *reinterpret_cast<TiledSmem<R>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_i = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int ntiles = gridDim.x, nsplit = gridDim.y;
  const float* A = p1 + (size_t)b * P * 3;
  const float* T = p2 + (size_t)b * M * 3;

  u64 px[R / 2], py[R / 2], pz[R / 2];
  float rm[R];
#pragma unroll
  for (int rp = 0; rp < R / 2; ++rp) {
    int i0 = min(tile_i * TM + (2 * rp) * kTThreads + tid, P - 1);
    int i1 = min(tile_i * TM + (2 * rp + 1) * kTThreads + tid, P - 1);
    px[rp] = pk(A[3 * (size_t)i0], A[3 * (size_t)i1]);
    py[rp] = pk(A[3 * (size_t)i0 + 1], A[3 * (size_t)i1 + 1]);
    pz[rp] = pk(A[3 * (size_t)i0 + 2], A[3 * (size_t)i1 + 2]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rm[r] = inf_f();
    sm.rs_best[r][tid] = inf_f(); sm.rs_thr[r][tid] = inf_f(); sm.rs_mask[r][tid] = 0ull;
  }
  const int c_first = split * cps;
  const int c_last = min(nchunks, c_first + cps);
  if (tid < kCW) {
    int col = min(c_first * kCW + tid, M - 1);
    sm.tile[0][tid] = make_float4(T[3 * (size_t)col], T[3 * (size_t)col + 1], T[3 * (size_t)col + 2], 0.f);
  }
  __syncthreads();

  for (int c = c_first; c < c_last; ++c) {
    const int buf = (c - c_first) & 1;
    // prefetch the next chunk into registers while this one is being consumed
    float nx = 0.f, ny = 0.f, nz = 0.f;
    const bool has_next = (c + 1 < c_last) && (tid < kCW);
    if (has_next) {
      int col = min((c + 1) * kCW + tid, M - 1);
      nx = T[3 * (size_t)col]; ny = T[3 * (size_t)col + 1]; nz = T[3 * (size_t)col + 2];
    }
    const float4* tl = sm.tile[buf];
    float* cw = sm.colw[buf][warp];
#pragma unroll 2
    for (int k = 0; k < kCW; k += 2) {
      const float4 c0 = tl[k], c1 = tl[k + 1];
      float cm0 = inf_f(), cm1 = inf_f();
#pragma unroll
      for (int rp = 0; rp < R / 2; ++rp) {
        u64 d0 = pair_d2<MODE>(px[rp], py[rp], pz[rp], c0.x, c0.y, c0.z, one);
        u64 d1 = pair_d2<MODE>(px[rp], py[rp], pz[rp], c1.x, c1.y, c1.z, one);
        float a0, b0, a1, b1;
        upk(d0, a0, b0); upk(d1, a1, b1);
        rm[2 * rp] = min3f(rm[2 * rp], a0, a1);
        rm[2 * rp + 1] = min3f(rm[2 * rp + 1], b0, b1);
        cm0 = min3f(cm0, a0, b0);
        cm1 = min3f(cm1, a1, b1);
      }
      float w0 = warp_min_f32(cm0), w1 = warp_min_f32(cm1);
      if (lane == 0) *reinterpret_cast<float2*>(cw + k) = make_float2(w0, w1);
    }
    // fold this chunk's row minima into the per-row candidate record
    const u64 bit = 1ull << (c - c_first);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float m = rm[r];
      rm[r] = inf_f();
      if (m <= sm.rs_thr[r][tid]) {
        const float best = sm.rs_best[r][tid];
        const float tm = thr_of<MODE>(m);
        u64 mask = (tm < best) ? 0ull : sm.rs_mask[r][tid];
        sm.rs_mask[r][tid] = mask | bit;
        if (m < best) { sm.rs_best[r][tid] = m; sm.rs_thr[r][tid] = tm; }
      }
    }
    if (has_next) sm.tile[buf ^ 1][tid] = make_float4(nx, ny, nz, 0.f);
    __syncthreads();
    // per-column candidate record of this tile: best warp minimum + mask of warps within slack
    if (tid < kCW) {
      const int col = c * kCW + tid;
      if (col < M) {
        float w[kTWarps];
        float best = inf_f();
#pragma unroll
        for (int i = 0; i < kTWarps; ++i) { w[i] = sm.colw[buf][i][tid]; best = fminf(best, w[i]); }
        const float t = thr_of<MODE>(best);
        unsigned mask = 0;
#pragma unroll
        for (int i = 0; i < kTWarps; ++i) mask |= (w[i] <= t) ? (1u << i) : 0u;
        size_t o = ((size_t)b * ntiles + tile_i) * M + col;
        cbest[o] = best; cmask[o] = mask;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int row = tile_i * TM + r * kTThreads + tid;
    if (row < P) {
      size_t o = ((size_t)b * nsplit + split) * P + row;
      rbest[o] = sm.rs_best[r][tid]; rmask[o] = sm.rs_mask[r][tid];
    }
  }
}

// Exact (min sqrt(d), first index) update shared by both recovery kernels.
struct ExactBest {
  float d, hi, v; int i;
  __device__ __forceinline__ void init() { d = inf_f(); hi = inf_f(); v = inf_f(); i = 0x7fffffff; }
  __device__ __forceinline__ void offer(float dd, int idx) {
    if (dd <= hi) {
      float vv = sqrtf(dd);
      if (vv < v || (vv == v && idx < i)) { v = vv; i = idx; }
      if (dd < d) { d = dd; hi = thr_of<MODE_EXACT>(dd); }
    }
  }
};

template <int MODE>
__global__ void __launch_bounds__(128)
chamfer_recover_rows_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                            const float* __restrict__ rbest, const u64* __restrict__ rmask,
                            float* __restrict__ min1, int* __restrict__ idx1,
                            int P, int M, int nsplit, int cps) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= P) return;
  const float* T = p2 + (size_t)b * M * 3;
  const float* a = p1 + 3 * ((size_t)b * P + row);
  const float ax = a[0], ay = a[1], az = a[2];
  float g = inf_f();
  for (int s = 0; s < nsplit; ++s) g = fminf(g, rbest[((size_t)b * nsplit + s) * P + row]);
  const float gthr = thr_of<MODE>(g);
  ExactBest eb; eb.init();
  for (int s = 0; s < nsplit; ++s) {
    size_t o = ((size_t)b * nsplit + s) * P + row;
    if (!(rbest[o] <= gthr)) continue;
    u64 mask = rmask[o];
    while (mask) {
      int cb = __ffsll((long long)mask) - 1;
      mask &= mask - 1;
      int c0 = (s * cps + cb) * kCW;
      int c1 = min(M, c0 + kCW);
      for (int col = c0; col < c1; ++col) {
        const float* t = T + 3 * (size_t)col;
        eb.offer(exact_d2s(ax, ay, az, __ldg(t), __ldg(t + 1), __ldg(t + 2)), col);
      }
    }
  }
  min1[(size_t)b * P + row] = eb.v;
  idx1[(size_t)b * P + row] = eb.i;
}

template <int MODE>
__global__ void __launch_bounds__(128)
chamfer_recover_cols_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                            const float* __restrict__ cbest, const unsigned* __restrict__ cmask,
                            float* __restrict__ min2, int* __restrict__ idx2,
                            int P, int M, int ntiles, int R) {
  const int b = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= M) return;
  const float* A = p1 + (size_t)b * P * 3;
  const float* t = p2 + 3 * ((size_t)b * M + col);
  const float tx = t[0], ty = t[1], tz = t[2];
  const int TM = kTThreads * R;
  float g = inf_f();
  for (int ti = 0; ti < ntiles; ++ti) g = fminf(g, cbest[((size_t)b * ntiles + ti) * M + col]);
  const float gthr = thr_of<MODE>(g);
  ExactBest eb; eb.init();
  for (int ti = 0; ti < ntiles; ++ti) {
    size_t o = ((size_t)b * ntiles + ti) * M + col;
    if (!(cbest[o] <= gthr)) continue;
    unsigned mask = cmask[o];
    while (mask) {
      int w = __ffs((int)mask) - 1;
      mask &= mask - 1;
      for (int r = 0; r < R; ++r) {
        int r0 = ti * TM + r * kTThreads + w * 32;
        int r1 = min(P, r0 + 32);
        for (int row = r0; row < r1; ++row) {
          const float* a = A + 3 * (size_t)row;
          eb.offer(exact_d2s(__ldg(a), __ldg(a + 1), __ldg(a + 2), tx, ty, tz), row);
        }
      }
    }
  }
  min2[(size_t)b * M + col] = eb.v;
  idx2[(size_t)b * M + col] = eb.i;
}

struct TiledPlan { int R, ntiles, nchunks, nsplit, cps; };

static bool make_plan(int B, int P, int M, int sm_count, TiledPlan& pl) {
  if (P < 1024 || M < kCW || B < 1) return false;
  pl.nchunks = (M + kCW - 1) / kCW;
  const int rs[3] = {16, 8, 4};
  int pick = 4;
  for (int k = 0; k < 3; ++k) {
    int R = rs[k], tm = kTThreads * R;
    long long nt = (P + tm - 1) / tm;
    double waste = (double)(nt * tm) / P;
    long long slots = (long long)sm_count * (R <= 8 ? 2 : 1);
    if (waste <= 1.26 && nt * B * pl.nchunks >= 2 * slots) { pick = R; break; }
  }
  pl.R = pick;
  int tm = kTThreads * pl.R;
  pl.ntiles = (P + tm - 1) / tm;
  long long base = (long long)pl.ntiles * B;
  long long slots = (long long)sm_count * (pl.R <= 8 ? 2 : 1);
  long long s = (4 * slots + base - 1) / base;
  long long smin = (pl.nchunks + kMaxChunksPerSplit - 1) / kMaxChunksPerSplit;
  if (s < smin) s = smin;
  if (s > pl.nchunks) s = pl.nchunks;
  if (s < 1) s = 1;
  pl.cps = (int)((pl.nchunks + s - 1) / s);
  pl.nsplit = (pl.nchunks + pl.cps - 1) / pl.cps;
  return pl.nsplit <= 65535;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TiledWs { size_t rbest, rmask, cbest, cmask, total; };
static TiledWs ws_layout(int B, int P, int M, const TiledPlan& pl) {
  TiledWs w; size_t o = 0;
  w.rbest = o; o += al256((size_t)B * pl.nsplit * P * 4);
  w.rmask = o; o += al256((size_t)B * pl.nsplit * P * 8);
  w.cbest = o; o += al256((size_t)B * pl.ntiles * M * 4);
  w.cmask = o; o += al256((size_t)B * pl.ntiles * M * 4);
  w.total = o;
  return w;
}

static int g_plan_sms = 148;

int chamfer_tiled_supported(int B, int P, int M) {
  TiledPlan pl;
  return make_plan(B, P, M, g_plan_sms, pl) ? 1 : 0;
}

size_t chamfer_tiled_workspace_bytes(int B, int P, int M) {
  TiledPlan pl;
  if (!make_plan(B, P, M, g_plan_sms, pl)) return 0;
  return ws_layout(B, P, M, pl).total;
}

template <int R, int MODE>
static int launch_main(const float* p1, const float* p2, char* ws, const TiledWs& wl, const TiledPlan& pl,
                       int B, int P, int M, cudaStream_t s) {
  static bool attr_set = false;
  size_t smem = sizeof(TiledSmem<R>);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chamfer_tiled_kernel<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { vpn_set_error("chamfer tiled: smem attribute: %s", cudaGetErrorString(e)); return VPN_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid(pl.ntiles, pl.nsplit, B);
  chamfer_tiled_kernel<R, MODE><<<grid, kTThreads, smem, s>>>(
      p1, p2, reinterpret_cast<float*>(ws + wl.rbest), reinterpret_cast<u64*>(ws + wl.rmask),
      reinterpret_cast<float*>(ws + wl.cbest), reinterpret_cast<unsigned*>(ws + wl.cmask),
      P, M, pl.nchunks, pl.cps, 1.0f);
  return vpn_check_launch("chamfer_tiled_kernel");
}

template <int MODE>
static int run_mode(const float* p1, const float* p2, float* min1, int* idx1, float* min2, int* idx2,
                    int B, int P, int M, char* ws, const TiledWs& wl, const TiledPlan& pl, cudaStream_t s) {
  int rc;
  switch (pl.R) {
    case 16: rc = launch_main<16, MODE>(p1, p2, ws, wl, pl, B, P, M, s); break;
    case 8:  rc = launch_main<8, MODE>(p1, p2, ws, wl, pl, B, P, M, s); break;
    default: rc = launch_main<4, MODE>(p1, p2, ws, wl, pl, B, P, M, s); break;
  }
  if (rc) return rc;
  chamfer_recover_rows_kernel<MODE><<<dim3((P + 127) / 128, B), 128, 0, s>>>(
      p1, p2, reinterpret_cast<const float*>(ws + wl.rbest), reinterpret_cast<const u64*>(ws + wl.rmask),
      min1, idx1, P, M, pl.nsplit, pl.cps);
  rc = vpn_check_launch("chamfer_recover_rows_kernel");
  if (rc) return rc;
  chamfer_recover_cols_kernel<MODE><<<dim3((M + 127) / 128, B), 128, 0, s>>>(
      p1, p2, reinterpret_cast<const float*>(ws + wl.cbest), reinterpret_cast<const unsigned*>(ws + wl.cmask),
      min2, idx2, P, M, pl.ntiles, pl.R);
  return vpn_check_launch("chamfer_recover_cols_kernel");
}

// mode: -1 auto, 0 exact, 1 diff
int chamfer_tiled_fwd(const float* p1, const float* p2, float* min1, int* idx1, float* min2, int* idx2,
                      int B, int P, int M, void* ws, size_t ws_bytes, int sm_count, int mode, cudaStream_t s) {
  (void)sm_count;
  TiledPlan pl;
  if (!make_plan(B, P, M, g_plan_sms, pl)) { vpn_set_error("chamfer tiled: unsupported shape"); return VPN_ERR_SHAPE; }
  TiledWs wl = ws_layout(B, P, M, pl);
  if (ws_bytes < wl.total) { vpn_set_error("chamfer tiled: workspace too small (%zu < %zu)", ws_bytes, wl.total); return VPN_ERR_WORKSPACE; }
  if (mode < 0) mode = MODE_DIFF;
  char* w = reinterpret_cast<char*>(ws);
  if (mode == MODE_EXACT) return run_mode<MODE_EXACT>(p1, p2, min1, idx1, min2, idx2, B, P, M, w, wl, pl, s);
  return run_mode<MODE_DIFF>(p1, p2, min1, idx1, min2, idx2, B, P, M, w, wl, pl, s);
}

}  // namespace vpn
