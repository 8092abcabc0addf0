// Chamfer nearest-neighbour search: tensor-core candidate filter for sm_100a (tcgen05 + TMEM).
//
// Same contract as chamfer_tiled_kernel (chamfer_tiled.cu): ONE pass over the (P x M) squared-distance matrix
// of modules/loss/chamfer_distance.py:14-23 leaves, for every row (predicted point) and every column (target),
// the smallest hot value and a bit mask of the places that can still hold the arg-min; the exact recovery
// kernels then redo only those places with the reference's own arithmetic, so min and arg-min stay bit-exact.
//
// What moves to the tensor core: the distance itself.  With coordinates centred on the row tile's centroid,
//     |p - t|^2 = |p|^2 + |t|^2 - 2 p.t = R_i . C_j,     a 16-term dot product of
//     R_i = [ahx ahx alx  ahy ahy aly  ahz ahz alz  sh sl 1 1  alx aly alz]      a = -2p = ah + al (tf32 split),
//     C_j = [thx tlx thx  thy tly thy  thz tlz thz  1  1  ch cl tlx tly tlz]     s = |p|^2 = sh + sl, c = |t|^2 = ch + cl
// (every product of two tf32 numbers is exact in the fp32 accumulator; the split keeps 22 bits per operand).
// A block is 128 rows x 128 columns: tcgen05.mma kind::tf32 M=128 N=128 K=8, two K steps, issued twice -
// D1 = R C^T (TMEM lane = row) and D2 = C R^T (TMEM lane = column) - so that BOTH minima are per-thread
// reductions after a tcgen05.ld.32x32b: no shuffles, no shared-memory exchange.  Per pair the CUDA cores
// spend one TMEM read + one FMNMX per direction (measured 375 clk per block per SM for the epilogue,
// tools/tc_probe.cu) instead of 4 packed FMA-pipe ops + 2 FMNMX in the CUDA-core kernel.
//
// Error of the hot value against the reference's d: the omitted split terms are <= 2 * 2^-22 |a||t| per
// coordinate, s and c carry <= 3u each, the accumulator adds <= 16 roundings; measured over random tiles
// 2^-21.9 (|p|+|t|)^2 (tools/tc_probe.cu test 4); the filter assumes E = 2^-19 (|p|+|t|)^2, which by the
// argument in DESIGN.md section 4.1 gives err <= max(16 E rho^2, 4 E d), covered twice by the slack
// 1e-4 rho^2 + 2^-15 |x|.
//
// CTA = 10 warps over one tile of NB x 128 rows, swept twice (D1 blocks, then D2 blocks): warps 0-7 read the
// accumulators (lane quarter q = warp % 4, every second block) and keep the per-row candidate records in shared
// memory / emit the per-column records, warp 8 builds the C_j operands of the next column chunk, warp 9 (one
// elected lane) issues the MMAs.  TMEM holds a ring of four 128-column accumulators; accumulators and column
// buffers are handed over with mbarriers (tcgen05.commit on the MMA side).
#include "common.cuh"

namespace vpn {

constexpr int kTcBlk = 128;                    // rows per block = columns per chunk = MMA M = MMA N
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;
constexpr int kTcBlkBytes = kTcBlk * 64;       // 128 points x 16 tf32
constexpr float kTcBig = 1.0e30f;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_inf() { return __int_as_float(0x7f800000); }

// shared-memory matrix descriptor, K-major, no swizzle: 8-row groups SBO bytes apart, the two 16-byte K halves
// of one MMA LBO bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B tf32, both K-major, N=128, M=128
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcBlk >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a), "l"(b), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool tc_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ float tc_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }
__device__ __forceinline__ float tc_min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float tc_thr(float x, float rel, float abs_) { return __fadd_ru(__fmaf_ru(fabsf(x), rel, x), abs_); }

// one operand row (16 tf32) of point `idx` inside a 128-point tile: 8-row groups of 512 B, four 128-B core
// matrices (8 rows x 16 B) per group, one per 4 consecutive k
__device__ __forceinline__ void tc_store_operand(unsigned char* tile, int idx, float4 k0, float4 k1, float4 k2, float4 k3) {
  unsigned char* base = tile + (idx >> 3) * 512 + (idx & 7) * 16;
  *reinterpret_cast<float4*>(base) = k0;
  *reinterpret_cast<float4*>(base + 128) = k1;
  *reinterpret_cast<float4*>(base + 256) = k2;
  *reinterpret_cast<float4*>(base + 384) = k3;
}
// x, y, z centred.  Row operand R_i (is_row) or column operand C_j; returns |.|^2
__device__ __forceinline__ float tc_make_operand(unsigned char* tile, int idx, float x, float y, float z, bool is_row) {
  const float n2 = fmaf(z, z, fmaf(y, y, x * x));
  const float nh = tc_tf32(n2), nl = tc_tf32(n2 - nh);
  if (is_row) { x *= -2.f; y *= -2.f; z *= -2.f; }
  const float xh = tc_tf32(x), yh = tc_tf32(y), zh = tc_tf32(z);
  const float xl = tc_tf32(x - xh), yl = tc_tf32(y - yh), zl = tc_tf32(z - zh);
  if (is_row)
    tc_store_operand(tile, idx, make_float4(xh, xh, xl, yh), make_float4(yh, yl, zh, zh), make_float4(zl, nh, nl, 1.f),
                     make_float4(1.f, xl, yl, zl));
  else
    tc_store_operand(tile, idx, make_float4(xh, xl, xh, yh), make_float4(yl, yh, zh, zl), make_float4(zh, 1.f, 1.f, nh),
                     make_float4(nl, xl, yl, zl));
  return n2;
}

#ifdef VPN_TC_PROF
// cycle accounting of CTA (0,0,0): [0] mma wait colfull [1] mma wait empty [2] mma issue [3] epi(row) wait full
// [4] epi(row) ld [5] epi(row) min+fold [6] epi(col) wait full [7] epi(col) ld [8] epi(col) min+fold [9] setup [10] total
__device__ long long g_tc_prof[16];
#define TC_PROF_DECL long long pt_ = clock64(); const bool pon_ = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0
#define TC_PROF(slot) do { if (pon_) { long long n_ = clock64(); atomicAdd((unsigned long long*)&g_tc_prof[slot], (unsigned long long)(n_ - pt_)); pt_ = n_; } } while (0)
#else
#define TC_PROF_DECL
#define TC_PROF(slot)
#endif

// dynamic shared memory carve-up (NB = row blocks per tile)
struct TcSmem {
  unsigned char* rows; unsigned char* cols; float* rs_best; float* rs_thr; u64* rs_mask; float* colw;
  u64* bars; float* red; uint32_t* tmem_slot;
};
__host__ __device__ inline size_t tc_smem_bytes(int NB) {
  return (size_t)NB * kTcBlkBytes + 2 * kTcBlkBytes + (size_t)NB * kTcBlk * 16 + 2 * (size_t)NB * kTcBlk * 4 + 16 * 8 + 64 * 4 + 16;
}
__device__ __forceinline__ TcSmem tc_carve(unsigned char* p, int NB) {
  TcSmem s;
  s.rows = p; p += (size_t)NB * kTcBlkBytes;
  s.cols = p; p += 2 * kTcBlkBytes;
  s.rs_mask = reinterpret_cast<u64*>(p); p += (size_t)NB * kTcBlk * 8;
  s.rs_best = reinterpret_cast<float*>(p); p += (size_t)NB * kTcBlk * 4;
  s.rs_thr = reinterpret_cast<float*>(p); p += (size_t)NB * kTcBlk * 4;
  s.colw = reinterpret_cast<float*>(p); p += 2 * (size_t)NB * kTcBlk * 4;
  s.bars = reinterpret_cast<u64*>(p); p += 16 * 8;
  s.red = reinterpret_cast<float*>(p); p += 64 * 4;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

// grid: x = row tile (NB * 128 rows), y = column split, z = sample
__global__ void __launch_bounds__(kTcThreads, 1)
chamfer_tc_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                  float* __restrict__ rbest, u64* __restrict__ rmask,
                  float* __restrict__ cbest, unsigned* __restrict__ cmask,
                  float2* __restrict__ tslack, int* __restrict__ fallback,
                  int P, int M, int NB, int nchunks, int cps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const TcSmem sm = tc_carve(smem_raw, NB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_i = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int ntiles = gridDim.x, nsplit = gridDim.y;
  const int TM = NB * kTcBlk;
  TC_PROF_DECL;
#ifdef VPN_TC_PROF
  const long long pstart_ = pt_;
#endif
  const float* A = p1 + (size_t)b * P * 3;
  const float* T = p2 + (size_t)b * M * 3;
  const int c_first = split * cps;
  const int c_last = min(nchunks, c_first + cps);
  // barriers (8 bytes each): +0..+24 accumulator stage full, +32..+56 stage empty, +64/+72 column buffer full,
  // +80/+88 column buffer empty
  const uint32_t bar0 = tc_smem_u32(sm.bars);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) tc_mbar_init(bar0 + 8 * i, 1);
    for (int i = 4; i < 8; ++i) tc_mbar_init(bar0 + 8 * i, 4);          // the four warps (one per TMEM lane quarter) that read the stage
    for (int i = 8; i < 12; ++i) tc_mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(tc_smem_u32(sm.tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- tile centroid and radius (any centre is valid; the centroid keeps the radius, hence the slack, small)
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = tid; i < TM; i += kTcThreads) {
    const int row = min(tile_i * TM + i, P - 1);
    sx += A[3 * (size_t)row]; sy += A[3 * (size_t)row + 1]; sz += A[3 * (size_t)row + 2];
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { sm.red[warp * 4] = sx; sm.red[warp * 4 + 1] = sy; sm.red[warp * 4 + 2] = sz; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) s += sm.red[w * 4 + tid];
    sm.red[48 + tid] = s / (float)TM;
  }
  __syncthreads();
  const float cx = sm.red[48], cy = sm.red[49], cz = sm.red[50];
  float rho2 = 0.f;
  for (int i = tid; i < TM; i += kTcThreads) {
    const int row = min(tile_i * TM + i, P - 1);
    const float x = __fsub_rn(A[3 * (size_t)row], cx), y = __fsub_rn(A[3 * (size_t)row + 1], cy), z = __fsub_rn(A[3 * (size_t)row + 2], cz);
    const float n2 = tc_make_operand(sm.rows + (size_t)(i >> 7) * kTcBlkBytes, i & 127, x, y, z, true);
    rho2 = fmaxf(rho2, n2);
    if (!(n2 < kTcBig)) rho2 = tc_inf();
    sm.rs_best[i] = tc_inf(); sm.rs_thr[i] = tc_inf(); sm.rs_mask[i] = 0ull;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rho2 = fmaxf(rho2, __shfl_xor_sync(0xffffffffu, rho2, o));
  if (lane == 0) sm.red[warp * 4 + 3] = rho2;
  tc_fence_async_smem();                      // operand rows were written with generic stores; the MMA reads them through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  rho2 = sm.red[3];
  for (int w = 1; w < kTcThreads / 32; ++w) rho2 = fmaxf(rho2, sm.red[w * 4 + 3]);
  if (!(rho2 < kTcBig)) { if (tid == 0) atomicOr(&fallback[b], 1); rho2 = 0.f; }
  const float slack_rel = 3.0517578125e-05f;                      // 2^-15
  const float slack_abs = __fmaf_ru(1.0e-4f, rho2, 1e-36f);
  if (tid == 0 && split == 0) tslack[(size_t)b * ntiles + tile_i] = make_float2(slack_rel, slack_abs);
  const uint32_t tbase = *sm.tmem_slot;
  if (warp == 0) TC_PROF(9);

  // The tile is swept twice: phase 0 accumulates D1 = R C^T blocks (TMEM lane = row -> row minima), phase 1
  // D2 = C R^T blocks (TMEM lane = column -> column minima).  One direction at a time leaves all four 128-column
  // accumulators of TMEM to one ring, deep enough to cover the MMA round trip.
  const int nc = c_last - c_first;
  const int per_phase = nc * NB;
  if (warp == kTcEpiWarps) {
    // ===== column-operand builder (both phases) =====
    for (int cc = 0; cc < 2 * nc; ++cc) {
      const int c = c_first + (cc < nc ? cc : cc - nc);
      const int cb = cc & 1, use = cc >> 1;
      tc_mbar_wait(bar0 + 80 + 8 * cb, (use & 1) ^ 1);
      unsigned char* dst = sm.cols + cb * kTcBlkBytes;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = k * 32 + lane;
        const int col = min(c * kTcBlk + j, M - 1);
        const float x = __fsub_rn(T[3 * (size_t)col], cx), y = __fsub_rn(T[3 * (size_t)col + 1], cy), z = __fsub_rn(T[3 * (size_t)col + 2], cz);
        const float n2 = tc_make_operand(dst, j, x, y, z, false);
        if (!(n2 < kTcBig)) atomicOr(&fallback[b], 1);
      }
      tc_fence_async_smem();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(bar0 + 64 + 8 * cb);
    }
  } else if (warp == kTcEpiWarps + 1) {
    // ===== MMA issuer: the whole warp walks the loop (warp-uniform operands), one elected lane issues =====
    {
      const uint32_t rows_a = tc_smem_u32(sm.rows), cols_a = tc_smem_u32(sm.cols);
      const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
      const uint64_t drows = tc_desc(rows_a);
      uint32_t it = 0;
      for (int cc = 0; cc < 2 * nc; ++cc) {
        const int cb = cc & 1, cuse = cc >> 1;
        const bool phase1 = cc >= nc;
        tc_mbar_wait(bar0 + 64 + 8 * cb, cuse & 1);
        tc_fence_after();
        TC_PROF(0);
        const uint64_t dc0 = tc_desc(cols_a + cb * kTcBlkBytes), dc1 = dc0 + (256 >> 4);
        for (int r = 0; r < NB; ++r, ++it) {
          const uint32_t st = it & 3;
          const uint64_t dr0 = drows + (uint64_t)(r * (kTcBlkBytes >> 4)), dr1 = dr0 + (256 >> 4);
          tc_mbar_wait(bar0 + 32 + 8 * st, ((it >> 2) & 1) ^ 1);
          tc_fence_after();
          TC_PROF(1);
          const uint32_t d = tb + st * 128;
          if (tc_elect()) {
            if (!phase1) { tc_mma(d, dr0, dc0, 0); tc_mma(d, dr1, dc1, 1); }      // D1[row][col]
            else         { tc_mma(d, dc0, dr0, 0); tc_mma(d, dc1, dr1, 1); }      // D2[col][row]
            tc_commit(bar0 + 8 * st);
          }
          __syncwarp();
          TC_PROF(2);
        }
        if (tc_elect()) tc_commit(bar0 + 80 + 8 * cb);               // column buffer free once these MMAs retire
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp = 4 g + q reads TMEM lanes [32 q, 32 q + 32) of the blocks with it % 2 == g =====
    // Each thread owns one TMEM lane and reads its 128 accumulator columns as two halves; the loads of the next
    // half (or of this warp's next block) are in flight while the current half is reduced.
    const int q = warp & 3, g = warp >> 2;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int li = q * 32 + lane;                                    // row in block (phase 0) / column in chunk (phase 1)
    const int total = 2 * per_phase;
    float v0[32], v1[32];
    if (g < total) {
      tc_mbar_wait(bar0 + 8 * g, 0);
      tc_fence_after();
      tc_ld32(tlane + g * 128, v0); tc_ld32(tlane + g * 128 + 32, v1);
    }
    for (int it = g; it < total; it += 2) {
      const uint32_t st = it & 3;
      float v2[32], v3[32];
      tc_wait_ld();
      if (q == 0 && g == 0) TC_PROF(3);
      tc_ld32(tlane + st * 128 + 64, v2); tc_ld32(tlane + st * 128 + 96, v3);
      float m0 = tc_inf(), m1 = tc_inf();
#pragma unroll
      for (int k = 0; k < 32; k += 2) { m0 = tc_min3(m0, v0[k], v0[k + 1]); m1 = tc_min3(m1, v1[k], v1[k + 1]); }
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(bar0 + 32 + 8 * st);             // all 128 columns are in registers: the accumulator can be overwritten
      if (q == 0 && g == 0) TC_PROF(4);
      if (it + 2 < total) {
        const uint32_t nst = (it + 2) & 3;
        tc_mbar_wait(bar0 + 8 * nst, ((it + 2) >> 2) & 1);
        tc_fence_after();
        tc_ld32(tlane + nst * 128, v0); tc_ld32(tlane + nst * 128 + 32, v1);
      }
#pragma unroll
      for (int k = 0; k < 32; k += 2) { m0 = tc_min3(m0, v2[k], v2[k + 1]); m1 = tc_min3(m1, v3[k], v3[k + 1]); }
      const float m = fminf(m0, m1);
      const bool phase1 = it >= per_phase;
      const int itp = phase1 ? it - per_phase : it;
      const int cl = itp / NB, r = itp - cl * NB;                    // chunk (relative to c_first), row block
      if (!phase1) {
        const int ri = r * kTcBlk + li;
        if (m <= sm.rs_thr[ri]) {
          const u64 bit = 1ull << cl;
          const float best = sm.rs_best[ri];
          const float tm = tc_thr(m, slack_rel, slack_abs);
          const u64 mask = (tm < best) ? 0ull : sm.rs_mask[ri];
          sm.rs_mask[ri] = mask | bit;
          if (m < best) { sm.rs_best[ri] = m; sm.rs_thr[ri] = tm; }
        }
      } else {
        float* cw = sm.colw + (size_t)(cl & 1) * NB * kTcBlk;
        cw[r * kTcBlk + li] = m;
        if (r >= NB - 2) {
          // the two warps of this lane quarter meet once per chunk; warp g = 1 (it holds r = NB - 1) merges
          asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");
          if (r == NB - 1) {
            const int col = (c_first + cl) * kTcBlk + li;
            if (col < M) {
              float best = tc_inf();
              for (int i = 0; i < NB; ++i) best = fminf(best, cw[i * kTcBlk + li]);
              const float t = tc_thr(best, slack_rel, slack_abs);
              unsigned mask = 0;
              for (int i = 0; i < NB; ++i) mask |= (cw[i * kTcBlk + li] <= t) ? (1u << i) : 0u;
              const size_t o = ((size_t)b * ntiles + tile_i) * M + col;
              cbest[o] = best; cmask[o] = mask;
            }
          }
        }
      }
      if (q == 0 && g == 0) TC_PROF(5);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef VPN_TC_PROF
  if (pon_ && warp == 0) atomicAdd((unsigned long long*)&g_tc_prof[10], (unsigned long long)(clock64() - pstart_));
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
  for (int i = tid; i < TM; i += kTcThreads) {
    const int row = tile_i * TM + i;
    if (row < P) {
      const size_t o = ((size_t)b * nsplit + split) * P + row;
      rbest[o] = sm.rs_best[i]; rmask[o] = sm.rs_mask[i];
    }
  }
}

// Samples whose coordinates are not finite or too large for the centred expansion (fallback[b] != 0) are
// redone here with the reference's arithmetic, thread per point, no candidate filter: correct, not fast.
// grid: x = slice, y = sample
__global__ void __launch_bounds__(256)
chamfer_flagged_kernel(const float* __restrict__ p1, const float* __restrict__ p2, float* __restrict__ min1,
                       int* __restrict__ idx1, float* __restrict__ min2, int* __restrict__ idx2,
                       const int* __restrict__ fallback, int P, int M) {
  const int b = blockIdx.y;
  if (fallback[b] == 0) return;
  const float* A = p1 + (size_t)b * P * 3;
  const float* T = p2 + (size_t)b * M * 3;
  for (int dir = 0; dir < 2; ++dir) {
    const float* X = dir ? T : A; const float* Y = dir ? A : T;
    const int nx = dir ? M : P, ny = dir ? P : M;
    float* mn = (dir ? min2 + (size_t)b * M : min1 + (size_t)b * P);
    int* ix = (dir ? idx2 + (size_t)b * M : idx1 + (size_t)b * P);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
      const float x = X[3 * (size_t)i], y = X[3 * (size_t)i + 1], z = X[3 * (size_t)i + 2];
      float bv = tc_inf(); int bj = 0;
      // torch.min over sqrt(d): first index of the smallest value; NaN propagates as in torch (first NaN wins)
      bool nan_seen = false;
      for (int j = 0; j < ny; ++j) {
        // (x - y)^2 == (y - x)^2 bit for bit, so the direction of the difference does not matter
        const float dx = __fsub_rn(x, Y[3 * (size_t)j]), dy = __fsub_rn(y, Y[3 * (size_t)j + 1]), dz = __fsub_rn(z, Y[3 * (size_t)j + 2]);
        const float v = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        if (!nan_seen) {
          if (v != v) { bv = v; bj = j; nan_seen = true; }
          else if (v < bv || j == 0) { bv = v; bj = j; }
        }
      }
      mn[i] = bv; ix[i] = bj;
    }
  }
}

int chamfer_tc_max_blocks() { return 16; }
size_t chamfer_tc_smem_bytes(int NB) { return tc_smem_bytes(NB); }

int chamfer_tc_launch(const float* p1, const float* p2, float* rbest, u64* rmask, float* cbest, unsigned* cmask,
                      float2* tslack, int* fallback, int B, int P, int M, int NB, int ntiles, int nsplit,
                      int nchunks, int cps, cudaStream_t s) {
  static int attr_for = 0;
  const size_t smem = tc_smem_bytes(NB);
  if (attr_for < (int)smem) {
    cudaError_t e = cudaFuncSetAttribute(chamfer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(16));
    if (e != cudaSuccess) { vpn_set_error("chamfer tc: smem attribute: %s", cudaGetErrorString(e)); return VPN_ERR_CUDA; }
    attr_for = (int)tc_smem_bytes(16);
  }
  chamfer_tc_kernel<<<dim3(ntiles, nsplit, B), kTcThreads, smem, s>>>(p1, p2, rbest, rmask, cbest, cmask, tslack, fallback,
                                                                      P, M, NB, nchunks, cps);
  return vpn_check_launch("chamfer_tc_kernel");
}

int chamfer_flagged_launch(const float* p1, const float* p2, float* min1, int* idx1, float* min2, int* idx2,
                           const int* fallback, int B, int P, int M, cudaStream_t s) {
  chamfer_flagged_kernel<<<dim3(32, B), 256, 0, s>>>(p1, p2, min1, idx1, min2, idx2, fallback, P, M);
  return vpn_check_launch("chamfer_flagged_kernel");
}

}  // namespace vpn
