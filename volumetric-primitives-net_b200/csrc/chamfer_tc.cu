// Chamfer nearest-neighbour search: tensor-core candidate filter for sm_100a (tcgen05 + TMEM).
//
// Same contract as chamfer_tiled_kernel (chamfer_tiled.cu): ONE pass over the (P x M) squared-distance matrix
// of modules/loss/chamfer_distance.py:14-23 leaves, for every row (predicted point) and every column (target),
// the smallest hot value and a bit mask of the places that can still hold the arg-min; the exact recovery
// kernels then redo only those places with the reference's own arithmetic, so min and arg-min stay bit-exact.
//
// What moves to the tensor core: the distance itself.  With coordinates centred on the row tile's centroid and
// scaled by a power of two S (so that every magnitude fits fp16: S * max(|p|, |t|) < 128),
//     S^2 |p - t|^2 = |P|^2 + |T|^2 - 2 P.T = R_i . C_j,     a 16-term dot product of fp16 numbers
//     R_i = [ahx ahx alx alx  ahy ahy aly aly  ahz ahz alz alz  sh sl 1  1 ]      a = -2P = ah + al (fp16 split),
//     C_j = [thx tlx thx tlx  thy tly thy tly  thz tlz thz tlz  1  1  ch cl]      s = |P|^2 = sh + sl, c = |T|^2 = ch + cl
// (every product of two fp16 numbers is exact in the fp32 accumulator; the split keeps 22 bits per operand).
// One tcgen05.mma kind::f16 M=128 N=256 K=16 fills one 128 x 256 accumulator ("stage"); it is issued in both
// orientations - D1 = R C^T (TMEM lane = row) and D2 = C R^T (TMEM lane = column) - so that BOTH minima are
// per-thread reductions after a tcgen05.ld.32x32b: no shuffles, no shared-memory exchange.  Per pair the CUDA
// cores spend one TMEM read + half an FMNMX3 per direction instead of 4 packed FMA-pipe ops + 2 FMNMX in the
// CUDA-core kernel.  (A tf32 version with K = 8 x 2 was measured first: each K step re-reads and re-writes the
// whole accumulator, 373 clk per stage on the tensor side alone; kind::f16 needs one K step.)
//
// Error of the hot value against the reference's d (scaled units): the split keeps a and t to 2^-22 relative,
// s and c carry <= 3u each, the accumulator adds <= 16 roundings: <= 2^-19.4 (|P|+|T|)^2; fp16 subnormal low
// parts add <= 2^-21 (|P|+|T|), which is below 2^-19 (|P|+|T|)^2 once |P|+|T| >= 1/4 and below 2^-23 otherwise.
// With E = 2^-19 (|P|+|T|)^2 the argument in DESIGN.md section 4.1 gives err <= max(16 E rho^2, 4 E d) for each of
// the two values compared, i.e. a threshold of 32 E rho^2 + 8 E |x|; the kernel uses four times that,
// 2.5e-4 rho^2 + 2^-20 + 2^-14 |x| (scaled units).  Measured error over random tiles (tools/tc_err.cu): 2^-19.85
// relative, 2^-23.7 absolute for |P|+|T| < 1/4.
//
// Pruning (chamfer_prep.cu): the targets arrive Morton-sorted, so a 128-column chunk is a compact patch, and 128
// consecutive rows are a patch of one primitive.  A stage whose two patches are further apart (box gap) than a distance
// every row of the block (column of the chunk) is known to achieve elsewhere is not computed at all: the skip bits are
// derived once per CTA in the prologue from the per-block boxes and bounds, and the MMA and epilogue warps walk the same
// bit masks.  The row and the column direction prune independently (phase 0 / phase 1).
//
// CTA = 20 warps over one tile of NB x 128 rows, swept twice (D1 blocks, then D2 blocks): warps 0-15 read the
// accumulators and keep the per-row candidate records in shared memory / emit the per-column records, warps 16-17
// build the C_j operands of the next 256 columns (one half each), warps 18-19 (one elected lane each) issue the MMAs of
// one column half each.  TMEM holds four 128 x 128 accumulators; accumulators and column buffers are handed over with
// mbarriers (tcgen05.commit on the MMA side).
// The accumulator is FP16 (round 2): the tensor core accumulates in f32 as before and rounds the result once; the
// epilogue then reads two values per register and reduces them with packed 16-bit integer mins, 2.6 x fewer ALU cycles
// per block than f32 + FMNMX3; the filter's slack grows by the rounding (2^-10 relative), the recovery stays exact.
#include <cuda_fp16.h>
#include "common.cuh"

namespace vpn {

constexpr int kTcBlk = 128;                    // rows per block = columns per chunk = MMA M = MMA N
constexpr int kTcEpiWarps = 16;
constexpr int kTcThreads = (kTcEpiWarps + 4) * 32;       // 16 epilogue warps, one operand builder and one MMA issuer per half
constexpr int kTcBlkBytes = kTcBlk * 32;       // 128 points x 16 fp16
constexpr float kTcBig = 1.0e30f;
constexpr int kTcDefaultHunits = 0;            // 32-column units per block reduced on the FP16 pipe (the rest: ALU pipe)

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_inf() { return __int_as_float(0x7f800000); }

// shared-memory matrix descriptor, K-major, no swizzle: 8-row groups SBO = 256 bytes apart, the two 16-byte K
// halves (8 fp16 each) of a row LBO = 128 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B f16, both K-major, N=256, M=128
constexpr uint32_t kTcIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(2 * kTcBlk >> 3) << 17) | ((uint32_t)(kTcBlk >> 4) << 24);

constexpr uint32_t kTcIdescHalf = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcBlk >> 4) << 24);   // N = 128
// the same with an FP16 accumulator (c_format = 0): the tensor core accumulates as before and rounds the result once, to
// nearest (tools/tc_f16acc.cu: f16 result == RN16(f32 result) on every probed tile); TMEM keeps one f16 per 32-bit column
constexpr uint32_t kTcIdescHalfF16 = (0u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTcBlk >> 3) << 17) | ((uint32_t)(kTcBlk >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate, uint32_t idesc = kTcIdesc) {
#if defined(VPN_TC_VARIANT) && VPN_TC_VARIANT == 3       // probe: epilogue alone
  return;
#endif
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
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
// 64 accumulator columns as 32 registers of two f16 each (column 2k in the low half of register k)
__device__ __forceinline__ void tc_ld32_pack(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
// min of two pairs of f16.  The two spellings compute the same for the numbers met here (no NaN: the sums are finite) and
// alternate along the reduction tree so that ptxas keeps 2-input HMNMX2 instead of fusing pairs into the 3-input VHMNMX,
// which runs at a fraction of the rate (tools/tc_f16acc.cu, clk per 128 x 128 block per SM with 16 epilogue warps: f32 +
// FMNMX3 385, VHMNMX 519, 3-input packed integer min VIMNMX3.S16x2 254, HMNMX2 133; TMEM loads alone 65).
__device__ __forceinline__ uint32_t tc_hmin2(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t tc_hmin2n(uint32_t a, uint32_t b) { uint32_t r; asm("min.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// min over 16 registers of two f16, per half
__device__ __forceinline__ uint32_t tc_hmin16(const uint32_t* v) {
  const uint32_t t0 = tc_hmin2(v[0], v[1]), t1 = tc_hmin2(v[2], v[3]), t2 = tc_hmin2(v[4], v[5]), t3 = tc_hmin2(v[6], v[7]);
  const uint32_t t4 = tc_hmin2(v[8], v[9]), t5 = tc_hmin2(v[10], v[11]), t6 = tc_hmin2(v[12], v[13]), t7 = tc_hmin2(v[14], v[15]);
  const uint32_t s0 = tc_hmin2n(t0, t1), s1 = tc_hmin2n(t2, t3), s2 = tc_hmin2n(t4, t5), s3 = tc_hmin2n(t6, t7);
  return tc_hmin2n(tc_hmin2(s0, s1), tc_hmin2(s2, s3));
}
// min of three pairs of 16-bit integers (ptxas fuses the two into one VIMNMX3.S16x2, ALU pipe)
__device__ __forceinline__ uint32_t tc_min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r, t;
  asm("min.s16x2 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(t), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t tc_min_s16x2(uint32_t a, uint32_t b) { uint32_t r; asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t tc_max_s16x2(uint32_t a, uint32_t b) { uint32_t r; asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// min over 16 registers of packed 16-bit values, per half
__device__ __forceinline__ uint32_t tc_imin16(const uint32_t* s) {
  uint32_t x = s[0];
#pragma unroll
  for (int k = 1; k < 15; k += 2) x = tc_min3_s16x2(x, s[k], s[k + 1]);
  return tc_min_s16x2(x, s[15]);
}
__device__ __forceinline__ float tc_min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float tc_thr(float x, float rel, float abs_) { return __fadd_ru(__fmaf_ru(fabsf(x), rel, x), abs_); }

// (a, b) -> hi = (f16(a), f16(b)) and lo = (f16(a - hi.a), f16(b - hi.b)), first element in the low half: x = hi + lo with
// hi, lo fp16 (lo may be subnormal).  One packed conversion per pair (F2FP) instead of two scalar F2F, which run on the
// slow conversion unit: the operand builders are on the critical path of a pruned sweep.
__device__ __forceinline__ uint32_t tc_cvt2(float first, float second) {
  uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first)); return r;
}
__device__ __forceinline__ void tc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = tc_cvt2(a, b);
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = tc_cvt2(a - back.x, b - back.y);
}
// X, Y, Z centred and scaled.  Writes the row operand R_i (is_row) or column operand C_j of point `idx` of a
// 128-point tile: 8-row groups of 256 B, two 128-B core matrices (8 rows x 16 B) per group, k 0-7 and k 8-15.
// Returns |.|^2 (scaled).
__device__ __forceinline__ float tc_make_operand(unsigned char* tile, int idx, float x, float y, float z, bool is_row) {
  const float n2 = fmaf(z, z, fmaf(y, y, x * x));
  if (is_row) { x *= -2.f; y *= -2.f; z *= -2.f; }
  uint32_t hxy, lxy, hzn, lzn;
  tc_split2(x, y, hxy, lxy); tc_split2(z, n2, hzn, lzn);
  const uint32_t ones = 0x3c003c00u;                                 // (1, 1)
  uint4 k0, k1;
  if (is_row) {                                                      // [xh xh xl xl  yh yh yl yl] [zh zh zl zl  nh nl 1 1]
    k0 = make_uint4(__byte_perm(hxy, hxy, 0x1010), __byte_perm(lxy, lxy, 0x1010), __byte_perm(hxy, hxy, 0x3232), __byte_perm(lxy, lxy, 0x3232));
    k1 = make_uint4(__byte_perm(hzn, hzn, 0x1010), __byte_perm(lzn, lzn, 0x1010), __byte_perm(hzn, lzn, 0x7632), ones);
  } else {                                                           // [xh xl xh xl  yh yl yh yl] [zh zl zh zl  1 1 nh nl]
    const uint32_t wx = __byte_perm(hxy, lxy, 0x5410), wy = __byte_perm(hxy, lxy, 0x7632), wz = __byte_perm(hzn, lzn, 0x5410);
    k0 = make_uint4(wx, wx, wy, wy);
    k1 = make_uint4(wz, wz, ones, __byte_perm(hzn, lzn, 0x7632));
  }
  unsigned char* base = tile + (idx >> 3) * 256 + (idx & 7) * 16;
  *reinterpret_cast<uint4*>(base) = k0;
  *reinterpret_cast<uint4*>(base + 128) = k1;
  return n2;
}

// dynamic shared memory carve-up (NB = row blocks per tile)
struct TcSmem {
  unsigned char* rows; unsigned char* cols; float* rs_best; uint4* rs_mask; float* colw;
  u64* bars; float* red; uint32_t* tmem_slot; uint32_t* skip;
};
// plan words of a tile (global): [0,64) per chunk c the row blocks r (bit r) whose 128 x 128 block is pruned for the row
// direction, [64,128) the same for the column direction.
constexpr int kTcPlanWords = 128;
constexpr int kTcSkipWords = 128;
__host__ __device__ inline size_t tc_smem_bytes(int NB) {
  return (size_t)NB * kTcBlkBytes + 4 * kTcBlkBytes + 2 * (size_t)NB * kTcBlk * 20 + 2 * 2 * 2 * kTcBlk * 16 + 16 * 8 + 96 * 4 + 16 +
         kTcSkipWords * 4;
}
__device__ __forceinline__ TcSmem tc_carve(unsigned char* p, int NB) {
  TcSmem s;
  s.rows = p; p += (size_t)NB * kTcBlkBytes;
  s.cols = p; p += 4 * kTcBlkBytes;                                   // two buffers of 256 columns
  s.rs_best = reinterpret_cast<float*>(p); p += 2 * (size_t)NB * kTcBlk * 4;      // [column half][row]
  s.rs_mask = reinterpret_cast<uint4*>(p); p += 2 * (size_t)NB * kTcBlk * 16;     // [column half][row]: 4 bits per chunk pair (32-column units)
  s.colw = reinterpret_cast<float*>(p); p += 2 * 2 * 2 * kTcBlk * 16;             // column record exchange: [p][turn][h][column] float4
  s.bars = reinterpret_cast<u64*>(p); p += 16 * 8;
  s.red = reinterpret_cast<float*>(p); p += 96 * 4;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p); p += 16;
  s.skip = reinterpret_cast<uint32_t*>(p);
  return s;
}

// Minima of this thread's TMEM lane over the 128 accumulator columns starting at taddr, one per 32-column UNIT (u[0..3]):
// the records name candidate units, not whole blocks, so that the exact recovery redoes a quarter of the pairs.
// 64 values are live at a time (the kernel runs 19 warps: 104 registers per thread); each unit is reduced as two
// interleaved chains.  tcgen05.wait::ld has no groups - it waits for every outstanding load of the thread - so a warp
// cannot overlap its own loads with its own reduction; the overlap comes from the other three epilogue warps of the SMSP.
__device__ __forceinline__ float tc_lane_min4x32(uint32_t taddr, uint32_t empty_bar, int lane, float (&u)[4]) {
#if defined(VPN_TC_VARIANT) && VPN_TC_VARIANT == 2       // probe: tensor side alone (no TMEM reads, no reduction)
  __syncwarp();
  if (lane == 0) tc_mbar_arrive(empty_bar);
  u[0] = u[1] = u[2] = u[3] = __int_as_float(0x7fc00000);   // NaN: no record is ever updated
  return u[0];
#else
  float va[32], vb[32];
  tc_ld32(taddr, va); tc_ld32(taddr + 32, vb);
  tc_wait_ld();
  float a0 = tc_inf(), a1 = tc_inf(), b0 = tc_inf(), b1 = tc_inf();
#if defined(VPN_TC_VARIANT) && VPN_TC_VARIANT == 1       // probe: TMEM reads, no reduction
  a0 = va[1] + vb[2];
#else
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    a0 = tc_min3(a0, va[k], va[k + 1]); a1 = tc_min3(a1, va[16 + k], va[17 + k]);
    b0 = tc_min3(b0, vb[k], vb[k + 1]); b1 = tc_min3(b1, vb[16 + k], vb[17 + k]);
  }
#endif
  const float m0 = fminf(a0, a1), m1 = fminf(b0, b1);
  tc_ld32(taddr + 64, va); tc_ld32(taddr + 96, vb);
  tc_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) tc_mbar_arrive(empty_bar);    // the block's accumulator is free again while the last minima are taken
  a0 = tc_inf(); a1 = tc_inf(); b0 = tc_inf(); b1 = tc_inf();
#if defined(VPN_TC_VARIANT) && VPN_TC_VARIANT == 1
  a0 = va[3] + vb[7];
#else
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    a0 = tc_min3(a0, va[k], va[k + 1]); a1 = tc_min3(a1, va[16 + k], va[17 + k]);
    b0 = tc_min3(b0, vb[k], vb[k + 1]); b1 = tc_min3(b1, vb[16 + k], vb[17 + k]);
  }
#endif
  const float m2 = fminf(a0, a1), m3 = fminf(b0, b1);
  u[0] = m0; u[1] = m1; u[2] = m2; u[3] = m3;
  return fminf(tc_min3(m0, m1, m2), m3);
#endif
}

// The same for an FP16 accumulator: 128 columns arrive as 64 registers of two f16, all loaded before the single wait, so
// the accumulator is released before any reduction.  The reduction runs on packed pairs, per 32-column unit either
//   * as 16-bit INTEGERS on the ALU pipe (tc_imin16, 8 VIMNMX3.S16x2): non-negative f16 numbers order like their bit
//     patterns; a negative hot value (a squared distance of ~0 whose rounding errors won) is a negative integer and wins
//     the min, in any order among negatives - it is then clamped to +0, which is at least as close to the true value (>= 0),
//     so the error bound holds for the clamped number; or
//   * as f16 on the FP16 pipe (tc_hmin16, 15 HMNMX2), negative values ordered as numbers (and clamped all the same).
// HUNITS of the four units take the second route.  With nothing else running the FP16 pipe is the faster one
// (tools/tc_f16acc.cu, clk per 128 x 128 block per SM with 16 epilogue warps: all integer 254, all f16 133, two and two
// 110), but inside the filter, next to the MMAs, every unit moved to it makes the kernel slower (C2 filter stage 0.595 ms
// with 0 units, 0.610 / 0.625 / 0.645 / 0.674 with 1 / 2 / 3 / 4): the default is all integer.  +inf (overflow: |P - T|^2 >=
// 65520 in scaled units) is the largest pattern either way.
template <int HUNITS>
__device__ __forceinline__ float tc_lane_min4x32_h(uint32_t taddr, uint32_t empty_bar, int lane, float (&u)[4]) {
  uint32_t a[32], b[32];
  tc_ld32_pack(taddr, a); tc_ld32_pack(taddr + 64, b);
  tc_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) tc_mbar_arrive(empty_bar);
  const uint32_t m0 = (HUNITS >= 4) ? tc_hmin16(a) : tc_imin16(a), m1 = (HUNITS >= 3) ? tc_hmin16(a + 16) : tc_imin16(a + 16);
  const uint32_t m2 = (HUNITS >= 2) ? tc_hmin16(b) : tc_imin16(b), m3 = (HUNITS >= 1) ? tc_hmin16(b + 16) : tc_imin16(b + 16);
  // (even, odd) column minima of two units -> (even0, even1) and (odd0, odd1): the min of the two is (u0, u1).  The integer
  // min is right for a mixed pair too: both routes leave ordinary f16 bit patterns.
  uint32_t u01 = tc_min_s16x2(__byte_perm(m0, m1, 0x5410), __byte_perm(m0, m1, 0x7632));
  uint32_t u23 = tc_min_s16x2(__byte_perm(m2, m3, 0x5410), __byte_perm(m2, m3, 0x7632));
  u01 = tc_max_s16x2(u01, 0u); u23 = tc_max_s16x2(u23, 0u);
  const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&u01)), f23 = __half22float2(*reinterpret_cast<const __half2*>(&u23));
  u[0] = f01.x; u[1] = f01.y; u[2] = f23.x; u[3] = f23.y;
  return fminf(tc_min3(u[0], u[1], u[2]), u[3]);
}

// nibble k of x -> low nibble of byte k
__device__ __forceinline__ u64 tc_spread_nibbles(uint32_t v) {
  u64 x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  return x;
}
// squared gap between two axis-aligned boxes (8 floats: lo xyz, hi xyz); an empty box gives +inf
__device__ __forceinline__ float tc_box_gap2(const float* __restrict__ a, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float g = fmaxf(0.f, fmaxf(a[k] - b[3 + k], b[k] - a[3 + k]));
    s = fmaf(g, g, s);
  }
  return s;
}

// Does pass (phase, chunk pair j) contain a block that is not pruned?  (The operand builder and the MMA issuers must
// agree on which column buffers exist.)  raw: the plan words of the phase (bit r of word c: block (r, c) pruned).
__device__ __forceinline__ bool tc_pass_needed(int j, int nc, int NB, const uint32_t* raw) {
  const uint32_t live = ~raw[2 * j] | ((2 * j + 1 < nc) ? ~raw[2 * j + 1] : 0u);
  return (live & ((1u << NB) - 1u)) != 0u;
}

// ---- plan: which stages of which tile are pruned, and the order the tiles are run in ---------------------------------
// One CTA per tile (tile_i, split, b).  Stage (row block r, chunk c) is prunable for the row direction when
// gap(box_r, box_c)^2 > T_r and for the column direction when gap^2 > U_c (chamfer_prep.cu): every pair of the stage is
// then at least sqrt(gap2) apart while T / U are distances the block's rows / the chunk's columns certainly achieve
// elsewhere (exact arithmetic).  1e-5 relative covers the roundings on both sides (~1e-6); a NaN gap or an infinite
// bound compares false: not skipped.  Row blocks past the end of the cloud are always prunable.
// Output per tile: kTcPlanWords mask words + the number of live 128 x 128 blocks (the tile's work).
constexpr int kPlanThreads = 128;
static_assert(kPlanThreads == kTcPlanWords, "one plan word per thread");
__global__ void __launch_bounds__(kPlanThreads)
chamfer_tc_plan_kernel(const float* __restrict__ cbox, const float* __restrict__ rbox, const float* __restrict__ rthr,
                       const float* __restrict__ cub, uint32_t* __restrict__ plan_masks, int* __restrict__ plan_work,
                       int ncta, int ntiles, int nsplit, int nrb_total, int NB, int nchunks, int cps) {
  __shared__ uint32_t skipR[64], skipC[64];
  __shared__ int s_live[kPlanThreads / 32];
  __shared__ float4 s_rbox[16][2];                                    // the tile's row-block boxes (8 floats each) and bounds
  __shared__ float s_rthr[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;                                         // one CTA per tile
  const int tile_i = cta % ntiles, split = (cta / ntiles) % nsplit, b = cta / (ntiles * nsplit);
  const int c_first = split * cps, nc = min(nchunks, c_first + cps) - c_first;
  if (tid < 64) { skipR[tid] = 0u; skipC[tid] = 0u; }
  if (tid < 2 * NB) {
    const int r = tid >> 1, rb = tile_i * NB + r;
    if (rb < nrb_total) s_rbox[r][tid & 1] = reinterpret_cast<const float4*>(rbox + ((size_t)b * nrb_total + rb) * 8)[tid & 1];
  } else if (tid >= 32 && tid < 32 + NB) {
    const int r = tid - 32, rb = tile_i * NB + r;
    s_rthr[r] = (rb < nrb_total) ? __fmul_ru(rthr[(size_t)b * nrb_total + rb], 1.00001f) : 0.f;
  }
  __syncthreads();
  // thread = (chunk c, half of the row blocks): the chunk's box stays in registers (two 16-byte loads), the row boxes are
  // broadcast reads of shared memory.  (One (r, c) pair per thread and turn with twelve scalar global loads per pair: 11 us.)
  {
    const int c = tid & 63, r0 = (tid >> 6) * (NB / 2);
    if (c < nc) {
      const float4 c0 = reinterpret_cast<const float4*>(cbox + ((size_t)b * nchunks + c_first + c) * 8)[0];
      const float4 c1 = reinterpret_cast<const float4*>(cbox + ((size_t)b * nchunks + c_first + c) * 8)[1];
      const float cb8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      const float ub = __fmul_ru(cub[(size_t)b * nchunks + c_first + c], 1.00001f);
      uint32_t mr = 0u, mc = 0u;
      for (int r = r0; r < r0 + NB / 2; ++r) {
        bool pr = true, pc = true;
        if (tile_i * NB + r < nrb_total) {
          const float4 a0 = s_rbox[r][0], a1 = s_rbox[r][1];
          const float rb8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float gap2 = tc_box_gap2(rb8, cb8);
          pr = gap2 > s_rthr[r];
          pc = gap2 > ub;
        }
        mr |= pr ? (1u << r) : 0u;
        mc |= pc ? (1u << r) : 0u;
      }
      atomicOr(&skipR[c], mr);
      atomicOr(&skipC[c], mc);
    }
  }
  __syncthreads();
  uint32_t* out = plan_masks + (size_t)cta * kTcPlanWords;
  int live = 0;                                                       // work in 128 x 128 blocks (half stages)
  {
    const int c = tid & 63;
    const uint32_t m = (c < nc) ? (tid < 64 ? skipR[c] : skipC[c]) : 0xffffffffu;
    if (c < nc) live = NB - __popc(m & ((1u << NB) - 1u));
    out[tid] = m;                                                     // kPlanThreads == kTcPlanWords
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
  if (lane == 0) s_live[warp] = live;
  __syncthreads();
  if (tid == 0) plan_work[cta] = s_live[0] + s_live[1] + s_live[2] + s_live[3];
}

// Counting sort of the tiles by work, heaviest first (work <= 2048 blocks).  One CTA.  The order among tiles of equal
// work depends on atomics; it only affects which SM runs which tile, never a result.
constexpr int kOrderThreads = 1024, kOrderBins = 2112;
__global__ void __launch_bounds__(kOrderThreads)
chamfer_tc_order_kernel(const int* __restrict__ plan_work, int* __restrict__ plan_order, int ncta) {
  __shared__ int hist[kOrderBins];
  const int tid = threadIdx.x;
  for (int i = tid; i < kOrderBins; i += kOrderThreads) hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < ncta; i += kOrderThreads) atomicAdd(&hist[min(max(plan_work[i], 0), kOrderBins - 1)], 1);
  __syncthreads();
  // exclusive prefix over the bins, heaviest bin first: thread t owns the kPer consecutive bins below kOrderBins - t kPer
  {
    constexpr int kPer = (kOrderBins + kOrderThreads - 1) / kOrderThreads;
    __shared__ int wsum[32];
    const int lane = tid & 31, warp = tid >> 5;
    int cnt[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) { const int w = kOrderBins - 1 - (tid * kPer + k); cnt[k] = (w >= 0) ? hist[w] : 0; sum += cnt[k]; }
    int inc = sum;
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
    int run = wsum[warp] + inc - sum;
#pragma unroll
    for (int k = 0; k < kPer; ++k) { const int w = kOrderBins - 1 - (tid * kPer + k); if (w >= 0) hist[w] = run; run += cnt[k]; }
  }
  __syncthreads();
  for (int i = tid; i < ncta; i += kOrderThreads) plan_order[atomicAdd(&hist[min(max(plan_work[i], 0), kOrderBins - 1)], 1)] = i;
}

// grid: x = tile (row tile NB * 128 rows, column split, sample) in the order of the plan
//
// Work unit ("stage") = one 128-lane x 256-column accumulator: tcgen05.mma kind::f16 M=128 N=256 K=16, one commit.
// TMEM (512 columns) holds two stages.  The column chunks of the split are taken in ADJACENT pairs (2j, 2j + 1) - after
// the Morton sort two neighbouring patches of the target shape: the 256-column operand buffer holds chunk 2j in its
// first half and chunk 2j + 1 in its second.
//   phase 0 (row minima)   : stage (j, r) = R_r (128 rows) x [C_2j C_2j+1]^T; TMEM lane = row, the thread of column
//                            half h reduces the 128 values of chunk 2j + h;
//   phase 1 (column minima): stage (chunk, rp) = C_chunk (128 columns) x [R_2rp R_2rp+1]^T; TMEM lane = column, the
//                            thread of half h reduces over the 128 rows of block 2 rp + h.
// Roles: warps 0-15 epilogue (warp = 8 p + 4 h + q: accumulator buffer p, column half h, TMEM lane quarter q), warps 16-17
// build the column operands of half 0 / 1, warps 18-19 issue the MMAs of half 0 / 1 (one elected lane).
// Measured (tools/tc_var.cu, clk per stage per SM): tensor side alone 324, epilogue alone 558 (TMEM reads 136 and
// 128 FMNMX3 = 256 do not overlap: they share the register-file write port), together 580.
template <int HUNITS>
__global__ void __launch_bounds__(kTcThreads, 1)
chamfer_tc_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                  float* __restrict__ rbest, u64* __restrict__ rmask,
                  float* __restrict__ cbest, u64* __restrict__ cmask,
                  float2* __restrict__ tslack, int* __restrict__ fallback, const float* __restrict__ tmax,
                  const uint32_t* __restrict__ plan_masks, const int* __restrict__ plan_order, u64* __restrict__ stats,
                  int ntiles, int nsplit, int P, int M, int NB, int nchunks, int cps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
#ifdef VPN_TC_COUNTERS
  const long long t_start = clock64();
#endif
  const TcSmem sm = tc_carve(smem_raw, NB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1-D grid.  With a plan (chamfer_tc_plan_kernel + chamfer_tc_order_kernel) block x runs the x-th heaviest tile: CTAs
  // are dispatched in block order, so the heavy ones start first and the light ones fill the tail (without it the SMs
  // idled 13 % of the kernel waiting for late heavy tiles).
  const int cta = plan_order ? plan_order[blockIdx.x] : (int)blockIdx.x;
  const int tile_i = cta % ntiles, split = (cta / ntiles) % nsplit, b = cta / (ntiles * nsplit);
  const int TM = NB * kTcBlk;
  const float* A = p1 + (size_t)b * P * 3;
  const float* T = p2 + (size_t)b * M * 3;
  const int c_first = split * cps;
  const int c_last = min(nchunks, c_first + cps);
  // barriers (8 bytes each): +0..+24 accumulator full [half g][buffer p], +32..+56 accumulator empty [g][p], +64/+72
  // column operand buffer full, +80/+88 column operand buffer empty
  const uint32_t bar0 = tc_smem_u32(sm.bars);
  const uint32_t bar_full = bar0, bar_empty = bar0 + 32, bar_cfull = bar0 + 64, bar_cempty = bar0 + 80;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      tc_mbar_init(bar_full + 8 * i, 1);
      tc_mbar_init(bar_empty + 8 * i, 4);                              // the four epilogue warps (lane quarters) that read the buffer
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(bar_cfull + 8 * i, 2);                              // both builders fill a column buffer
      tc_mbar_init(bar_cempty + 8 * i, 2);                             // both issuers release a column buffer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(tc_smem_u32(sm.tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- tile centroid, radius and scale (any centre is valid; the centroid keeps the radius, hence the slack, small).
  // Each thread keeps its rows (<= 7 of the 2048) in registers across the three steps.
  constexpr int kRowsPerThread = (16 * kTcBlk + kTcThreads - 1) / kTcThreads;      // 4
  float rx[kRowsPerThread], ry[kRowsPerThread], rz[kRowsPerThread];
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int i = tid + k * kTcThreads;
    rx[k] = ry[k] = rz[k] = 0.f;
    if (i < TM) {
      const int row = min(tile_i * TM + i, P - 1);
      rx[k] = A[3 * (size_t)row]; ry[k] = A[3 * (size_t)row + 1]; rz[k] = A[3 * (size_t)row + 2];
      sx += rx[k]; sy += ry[k]; sz += rz[k];
    }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { sm.red[warp * 4] = sx; sm.red[warp * 4 + 1] = sy; sm.red[warp * 4 + 2] = sz; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) s += sm.red[w * 4 + tid];
    sm.red[80 + tid] = s / (float)TM;
  }
  __syncthreads();
  const float cx = sm.red[80], cy = sm.red[81], cz = sm.red[82];
  float rho2 = 0.f;
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    if (tid + k * kTcThreads < TM) {
      rx[k] = __fsub_rn(rx[k], cx); ry[k] = __fsub_rn(ry[k], cy); rz[k] = __fsub_rn(rz[k], cz);
      const float n2 = fmaf(rz[k], rz[k], fmaf(ry[k], ry[k], rx[k] * rx[k]));
      rho2 = (n2 < kTcBig) ? fmaxf(rho2, n2) : tc_inf();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rho2 = fmaxf(rho2, __shfl_xor_sync(0xffffffffu, rho2, o));
  if (lane == 0) sm.red[warp * 4 + 3] = rho2;
  __syncthreads();
  rho2 = sm.red[3];
  for (int w = 1; w < kTcThreads / 32; ++w) rho2 = fmaxf(rho2, sm.red[w * 4 + 3]);
  // scale: |P|, |T| < 128 in scaled units.  |t - c| <= sqrt(3) (max|t_k| + max|c_k|) for every target of the sample.
  float reach = fmaxf(sqrtf(rho2), 1.7320509f * (tmax[b] + fmaxf(fabsf(cx), fmaxf(fabsf(cy), fabsf(cz))))) * 1.001f;
  if (!(reach < 1.0e15f) || !(tmax[b] < 1.0e15f)) { if (tid == 0) atomicOr(&fallback[b], 1); reach = 1.f; rho2 = 0.f; }      // tmax: NaN if any target coordinate is
  int ex = 0;
  frexpf(reach, &ex);                                               // reach < 2^ex
  ex = max(-40, min(50, ex));
  const float S = scalbnf(1.f, 7 - ex), invS2 = scalbnf(1.f, 2 * ex - 14);
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int i = tid + k * kTcThreads;
    if (i < TM) {
      tc_make_operand(sm.rows + (size_t)(i >> 7) * kTcBlkBytes, i & 127, rx[k] * S, ry[k] * S, rz[k] * S, true);
      sm.rs_best[i] = tc_inf(); sm.rs_best[TM + i] = tc_inf();
      sm.rs_mask[i] = make_uint4(0u, 0u, 0u, 0u); sm.rs_mask[TM + i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc_fence_async_smem();                      // operand rows were written with generic stores; the MMA reads them through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all hot values and thresholds below are in scaled units (x S^2); records are stored unscaled
  // the true arg-min and the best hot value each err by <= max(16 E rho^2, 4 E d): the threshold needs 32 E rho^2 + 8 E |x|
  // = 6.1e-5 rho^2 + 2^-16 |x| at E = 2^-19; both are taken 4 x larger (measured E over random tiles: 2^-19.85)
  // The accumulator is FP16: the hot value h is RN16 of the f32 hot value f, |h - f| <= u |f| + s with
  // u = 2^-11, s = 2^-25 (subnormal range).  If the f32 values satisfy x_f <= b_f (1 + r) + a (the two lines above), then
  // x_h <= x_f (1 + u) + s and b_f <= (b_h + s) / (1 - u) give x_h <= b_h (1 + r + 2 u + ...) + a (1 + u) + 3 s: the
  // relative slack grows by 2 u (1 + u + r) < 1.125 x 2^-10, the absolute one by the factor 1.001 and 2^-22.
  const float slack_rel = 6.103515625e-05f + 1.0986328125e-03f;   // 2^-14 + 1.125 x 2^-10
  const float slack_abs = __fmaf_ru(2.5025e-4f, rho2 * (S * S), 1.1930e-06f);      // (2.5e-4 rho^2 + 2^-20: fp16 subnormal low parts) x 1.001 + 2^-22
  if (tid == 0 && split == 0) tslack[(size_t)b * ntiles + tile_i] = make_float2(slack_rel, slack_abs * invS2);
  const uint32_t tbase = *sm.tmem_slot;

  const int nc = c_last - c_first;             // chunks of this split (<= 64)
  const int hc = (nc + 1) >> 1;                // chunk pairs (<= 32)
  // ---- stage skip masks of this tile (computed by chamfer_tc_plan_kernel).  Without a plan every stage is computed.
  const uint32_t* rawR = sm.skip; const uint32_t* rawC = sm.skip + 64;
  for (int i = tid; i < kTcPlanWords; i += kTcThreads) sm.skip[i] = plan_masks ? plan_masks[(size_t)cta * kTcPlanWords + i] : 0u;
  __syncthreads();
  // statistics: stats[0] 128 x 128 blocks, stats[1] blocks skipped - counted by an MMA warp's lanes in parallel, two atomics per CTA
  // (thread 0 doing it alone, plus cycle counters, delayed epilogue warp 0 and with it every stage: +3 %).  A probe build
  // (-DVPN_TC_COUNTERS) adds the cycle counters of epilogue warp 0: [2] prologue, [3] row phase, [4] column phase, [5]
  // tail, and [6] / [7] live stages per phase, [8] live chunks, [9] operand passes built.
  if (warp == kTcEpiWarps + 2 && stats != nullptr) {
    unsigned s0 = 0, s1 = 0;
    for (int c = lane; c < nc; c += 32) { s0 += __popc(rawR[c] & ((1u << NB) - 1u)); s1 += __popc(rawC[c] & ((1u << NB) - 1u)); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (lane == 0) {
      atomicAdd(&stats[0], (u64)(2 * nc * NB));                         // 128 x 128 blocks, both directions
      atomicAdd(&stats[1], (u64)(s0 + s1));
#ifdef VPN_TC_COUNTERS
      atomicAdd(&stats[6], (u64)(nc * NB - s0));
      atomicAdd(&stats[7], (u64)(nc * NB - s1));
#endif
    }
  }
#ifdef VPN_TC_COUNTERS
  long long t_mark = 0;
  if (tid == 0 && stats != nullptr) {
    unsigned livechunks = 0, passes = 0;
    for (int c = 0; c < nc; ++c) livechunks += ((~rawC[c]) & ((1u << NB) - 1u)) ? 1u : 0u;
    for (int cc = 0; cc < 2 * hc; ++cc) passes += tc_pass_needed(cc < hc ? cc : cc - hc, nc, NB, cc < hc ? rawR : rawC) ? 1u : 0u;
    atomicAdd(&stats[8], (u64)livechunks);
    atomicAdd(&stats[9], (u64)passes);
    t_mark = clock64();
    atomicAdd(&stats[2], (u64)(t_mark - t_start));
  }
#define TC_MARK(slot) if (tid == 0 && stats != nullptr) { const long long t_ = clock64(); atomicAdd(&stats[slot], (u64)(t_ - t_mark)); t_mark = t_; }
#else
#define TC_MARK(slot)
#endif
  if (warp >= kTcEpiWarps && warp < kTcEpiWarps + 2) {
    // ===== column-operand builder of half g (one pass per phase): the 128 columns of chunk 2 j + g, 4 per lane =====
    // Only passes that contain a live stage are built (buffer = built-pass count & 1), and of such a pass only the halves
    // that have one (the other builder just takes part in the hand-over).  The 12 coordinate loads of the NEXT pass are
    // issued before this warp waits for that pass's buffer, so their latency hides behind the MMAs that still read it:
    // with most stages pruned a pass is short, and an L2 round trip per pass would set the pace.  (One warp building both
    // halves, with scalar f32 -> f16 conversions, took ~1000 clk per pass - as long as the pass's MMAs and epilogues.)
    const int g = warp - kTcEpiWarps;
    uint32_t seq = 0;
    float tx[4], ty[4], tz[4];
    const uint32_t nbmask = (1u << NB) - 1u;
    auto next_needed = [&](int cc) { while (cc < 2 * hc && !tc_pass_needed(cc < hc ? cc : cc - hc, nc, NB, cc < hc ? rawR : rawC)) ++cc; return cc; };
    auto half_live = [&](int cc) {
      const int c = 2 * (cc < hc ? cc : cc - hc) + g;
      return c < nc && ((~(cc < hc ? rawR : rawC)[c]) & nbmask) != 0u;
    };
    auto load_pass = [&](int cc) {
      if (!half_live(cc)) return;
      const int chunk = c_first + 2 * (cc < hc ? cc : cc - hc) + g;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = min(chunk * kTcBlk + k * 32 + lane, M - 1);
        tx[k] = T[3 * (size_t)col]; ty[k] = T[3 * (size_t)col + 1]; tz[k] = T[3 * (size_t)col + 2];
      }
    };
    int cc = next_needed(0);
    if (cc < 2 * hc) load_pass(cc);
    while (cc < 2 * hc) {
      const int cb = seq & 1, use = seq >> 1;
      ++seq;
      tc_mbar_wait(bar_cempty + 8 * cb, (use & 1) ^ 1);
      if (half_live(cc)) {
        unsigned char* dst = sm.cols + cb * 2 * kTcBlkBytes + g * kTcBlkBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x = __fsub_rn(tx[k], cx), y = __fsub_rn(ty[k], cy), z = __fsub_rn(tz[k], cz);
          const float n2 = tc_make_operand(dst, k * 32 + lane, x * S, y * S, z * S, false);
          if (!(n2 < 20000.f)) atomicOr(&fallback[b], 1);           // cannot happen for finite targets (|T| < 128)
        }
        tc_fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(bar_cfull + 8 * cb);
      cc = next_needed(cc + 1);
      if (cc < 2 * hc) load_pass(cc);
    }
  } else if (warp >= kTcEpiWarps + 2) {
    // ===== MMA issuer of half g: the whole warp walks the loop (warp-uniform operands), one elected lane issues =====
    // Work unit = one 128 x 128 block, tcgen05.mma kind::f16 M = 128, N = 128, K = 16 into the half's accumulator buffer p
    // (TMEM columns [256 p + 128 g, + 128)), which belongs to the four epilogue warps (g, ., p).
    //   phase 0: blocks (row block r, chunk 2 j + g), buffer p = r & 1;
    //   phase 1: blocks (chunk c, row block r) with r & 1 == g, buffer p = c & 1 - the two chunks of a pass are interleaved.
    const int g = warp - (kTcEpiWarps + 2);
    const uint32_t rows_a = tc_smem_u32(sm.rows), cols_a = tc_smem_u32(sm.cols);
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0) + g * 128;
    const uint64_t drows = tc_desc(rows_a);
    const uint32_t my_full = bar_full + 16 * g, my_empty = bar_empty + 16 * g;
    const uint32_t nbmask = (1u << NB) - 1u;
    uint32_t uses[2] = {0u, 0u};                                      // blocks issued into buffer p so far
    uint32_t seq = 0;
    // one block: wait until the buffer's previous block has been read, issue, signal `full` when the MMA retires
    auto issue = [&](int p, uint64_t da, uint64_t db) {
      const uint32_t n = p ? uses[1] : uses[0];
      tc_mbar_wait(my_empty + 8 * p, (n & 1) ^ 1);
      tc_fence_after();
      if (tc_elect()) {
        tc_mma(tb + p * 256, da, db, 0, kTcIdescHalfF16);
        tc_commit(my_full + 8 * p);
      }
      __syncwarp();
      if (p) ++uses[1]; else ++uses[0];
    };
    for (int cc = 0; cc < 2 * hc; ++cc) {
      const int j = cc < hc ? cc : cc - hc;
      if (!tc_pass_needed(j, nc, NB, cc < hc ? rawR : rawC)) continue;
      const int cb = seq & 1, cuse = seq >> 1;
      ++seq;
      tc_mbar_wait(bar_cfull + 8 * cb, cuse & 1);
      tc_fence_after();
      const uint64_t dcols = tc_desc(cols_a + cb * 2 * kTcBlkBytes);
      uint32_t issued = 0;
      if (cc < hc) {
        const uint64_t dc = dcols + (uint64_t)(g * (kTcBlkBytes >> 4));
        for (uint32_t live = (2 * j + g < nc) ? (~rawR[2 * j + g] & nbmask) : 0u; live; live &= live - 1) {
          const int r = __ffs((int)live) - 1;
          issue(r & 1, drows + (uint64_t)(r * (kTcBlkBytes >> 4)), dc);                   // D[row][column of the chunk]
          ++issued;
        }
      } else {
        uint32_t live0 = ~rawC[2 * j] & nbmask & (0x55555555u << g);
        uint32_t live1 = (2 * j + 1 < nc) ? (~rawC[2 * j + 1] & nbmask & (0x55555555u << g)) : 0u;
        while (live0 | live1) {
          if (live0) {
            const int r = __ffs((int)live0) - 1; live0 &= live0 - 1;
            issue(0, dcols, drows + (uint64_t)(r * (kTcBlkBytes >> 4)));                  // D[column][row of the block]
            ++issued;
          }
          if (live1) {
            const int r = __ffs((int)live1) - 1; live1 &= live1 - 1;
            issue(1, dcols + (uint64_t)(kTcBlkBytes >> 4), drows + (uint64_t)(r * (kTcBlkBytes >> 4)));
            ++issued;
          }
        }
      }
      // column buffer free once this half's MMAs on it retire (a half without a block in the pass just arrives)
      if (tc_elect()) { if (issued) tc_commit(bar_cempty + 8 * cb); else tc_mbar_arrive(bar_cempty + 8 * cb); }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp = 8 p + 4 h + q reads accumulator buffer p of half h: TMEM lanes [32 q, 32 q + 32), columns
    // [256 p + 128 h, + 128).  The four warps of an SMSP (q) work on four different blocks at any time: while one waits
    // for its TMEM loads the others reduce.  (With one 256-column stage shared by two warps per SMSP they moved in lock
    // step - both loading, then both reducing - and the load latency was exposed: 580 clk per stage against 256 of FMNMX3
    // issue; two independent warps per SMSP: 353 clk per 128 x 128 block, the serial load -> reduce chain of a warp.)
    const int q = warp & 3, h = (warp >> 2) & 1, p = warp >> 3;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + p * 256 + h * 128;
    const int li = q * 32 + lane;                                    // row in block (phase 0) / column in chunk (phase 1)
    const uint32_t my_full = bar_full + 16 * h + 8 * p, my_empty = bar_empty + 16 * h + 8 * p;
    uint32_t it = 0;                                                 // blocks read from my buffer so far
    // ---- phase 0: row minima; record (h, row): best value and, per chunk pair j, four bits naming the 32-column units of
    // chunk 2 j + h that may hold the arg-min (word j >> 3, nibble j & 7).  Row block r belongs to the warps with p = r & 1:
    // a record has one writer.
    float* my_best = sm.rs_best + (size_t)h * TM;
    uint4* my_mask = sm.rs_mask + (size_t)h * TM;
    const uint32_t nbmask = (1u << NB) - 1u;
    for (int j = 0; j < hc; ++j) {
      if (2 * j + h >= nc) break;
      // only the live blocks are visited (a per-block `if skipped continue` cost as many instructions as the blocks left)
      for (uint32_t live = ~rawR[2 * j + h] & nbmask & (0x55555555u << p); live; live &= live - 1) {
        const int r = __ffs((int)live) - 1;
        tc_mbar_wait(my_full, it & 1);
        tc_fence_after();
        float mu[4];
        const float m = tc_lane_min4x32_h<HUNITS>(tlane, my_empty, lane, mu);
        ++it;
        const int ri = r * kTcBlk + li;
        const float best = my_best[ri];
        if (m <= tc_thr(best, slack_rel, slack_abs)) {
          // units within the slack of the best value known now (the final filter uses the final, smaller, best)
          const float t = tc_thr(fminf(best, m), slack_rel, slack_abs);
          const uint32_t nib = (mu[0] <= t ? 1u : 0u) | (mu[1] <= t ? 2u : 0u) | (mu[2] <= t ? 4u : 0u) | (mu[3] <= t ? 8u : 0u);
          if (tc_thr(m, slack_rel, slack_abs) < best) my_mask[ri] = make_uint4(0u, 0u, 0u, 0u);   // everything recorded so far is out
          reinterpret_cast<uint32_t*>(my_mask + ri)[j >> 3] |= nib << ((j & 7) * 4);
          if (m < best) my_best[ri] = m;
        }
      }
    }
    TC_MARK(3)
    // ---- phase 1: column minima of the chunks c with c & 1 == p.  Thread = one column of the chunk; it keeps a running
    // record (best, 64-bit mask of the tile's 32-row UNITS: bit 4 r + u) of the row blocks r (r & 1 == h) it sees, in
    // REGISTERS, with the same update rule as the row records; at the end of a chunk the two warps (h = 0 / 1) of this lane
    // quarter and parity exchange their records through shared memory (16 bytes per column) and the warp whose turn it
    // is merges the two and writes the (tile, column) record.  (A per-block value array in shared memory merged in two
    // passes per chunk cost ~400 clk per chunk on the critical path.)
    int cseq = 0;
    float4* xch = reinterpret_cast<float4*>(sm.colw) + (size_t)p * 4 * kTcBlk;       // [turn][h][128] exchange slots of parity p
    for (int c = p; c < nc; c += 2) {
      const uint32_t livec = ~rawC[c] & nbmask;                      // row blocks of this chunk that are computed
      if (livec == 0u) {
        // chunk pruned for the whole tile: record (+inf, no candidate) - it is never a candidate in the recovery.  No
        // exchange, no barrier, and it does not take part in the two warps' alternation (cseq counts live chunks).
        const int col = (c_first + c) * kTcBlk + li;
        if (h == 0 && col < M) { const size_t o = ((size_t)b * ntiles + tile_i) * M + col; cbest[o] = tc_inf(); cmask[o] = 0ull; }
        continue;
      }
      float best = tc_inf(); u64 mask = 0ull;
      for (uint32_t live = livec & (0x55555555u << h); live; live &= live - 1) {
        const int r = __ffs((int)live) - 1;
        tc_mbar_wait(my_full, it & 1);
        tc_fence_after();
        float mu[4];
        const float m = tc_lane_min4x32_h<HUNITS>(tlane, my_empty, lane, mu);
        ++it;
        if (m <= tc_thr(best, slack_rel, slack_abs)) {
          const float t = tc_thr(fminf(best, m), slack_rel, slack_abs);
          const uint32_t nib = (mu[0] <= t ? 1u : 0u) | (mu[1] <= t ? 2u : 0u) | (mu[2] <= t ? 4u : 0u) | (mu[3] <= t ? 8u : 0u);
          const float tm = tc_thr(m, slack_rel, slack_abs);
          mask = ((tm < best) ? 0ull : mask) | ((u64)nib << (4 * r));
          best = fminf(best, m);
        }
      }
      float4* slot = xch + (size_t)(cseq & 1) * 2 * kTcBlk;
      slot[h * kTcBlk + li] = make_float4(best, __uint_as_float((uint32_t)mask), __uint_as_float((uint32_t)(mask >> 32)), 0.f);
      // the two warps (h = 0 / 1) of this lane quarter and parity meet once per live chunk; they take turns merging
      asm volatile("bar.sync %0, 64;" :: "r"(1 + q + 4 * p) : "memory");
      if ((cseq & 1) == h) {
        const int col = (c_first + c) * kTcBlk + li;
        if (col < M) {
          const float4 other = slot[(h ^ 1) * kTcBlk + li];
          const float b2 = fminf(best, other.x);
          const float t = tc_thr(b2, slack_rel, slack_abs);
          const u64 omask = (u64)__float_as_uint(other.y) | ((u64)__float_as_uint(other.z) << 32);
          const u64 m2 = ((best <= t) ? mask : 0ull) | ((other.x <= t) ? omask : 0ull);
          const size_t o = ((size_t)b * ntiles + tile_i) * M + col;
          cbest[o] = b2 * invS2; cmask[o] = m2;
        }
      }
      ++cseq;
    }
  }
  TC_MARK(4)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
  const size_t mask_plane = (size_t)gridDim.x / ntiles * P;           // B * nsplit * P records per plane
  for (int i = tid; i < TM; i += kTcThreads) {
    const int row = tile_i * TM + i;
    if (row < P) {
      const float b0 = sm.rs_best[i], b1 = sm.rs_best[TM + i];
      const float best = fminf(b0, b1);
      const float t = tc_thr(best, slack_rel, slack_abs);
      // record (h, row) nibble j names the units of chunk 2 j + h of the split: unit 4 (2 j + h) + u = bit 8 j + 4 h + u of the
      // 256-bit row mask, stored as four planes of u64 (plane w = units [64 w, 64 w + 64))
      const uint4 k0 = sm.rs_mask[i], k1 = sm.rs_mask[TM + i];
      const bool in0 = b0 <= t, in1 = b1 <= t;
      const size_t o = ((size_t)b * nsplit + split) * P + row;
      rbest[o] = best * invS2;
      rmask[o] = (in0 ? tc_spread_nibbles(k0.x) : 0ull) | (in1 ? (tc_spread_nibbles(k1.x) << 4) : 0ull);
      rmask[mask_plane + o] = (in0 ? tc_spread_nibbles(k0.y) : 0ull) | (in1 ? (tc_spread_nibbles(k1.y) << 4) : 0ull);
      rmask[2 * mask_plane + o] = (in0 ? tc_spread_nibbles(k0.z) : 0ull) | (in1 ? (tc_spread_nibbles(k1.z) << 4) : 0ull);
      rmask[3 * mask_plane + o] = (in0 ? tc_spread_nibbles(k0.w) : 0ull) | (in1 ? (tc_spread_nibbles(k1.w) << 4) : 0ull);
    }
  }
  TC_MARK(5)
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

// largest |coordinate| of every sample's targets (the tensor-core kernel scales by it).  grid: x = sample
__global__ void __launch_bounds__(256)
chamfer_tc_bounds_kernel(const float* __restrict__ p2, float* __restrict__ tmax, int M) {
  __shared__ float red[8];
  const float* T = p2 + (size_t)blockIdx.x * M * 3;
  float m = 0.f;
  for (int i = threadIdx.x; i < 3 * M; i += 256) {
    const float v = fabsf(T[i]);
    m = (v <= m) ? m : v;                                           // NaN sticks
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const float x = __shfl_xor_sync(0xffffffffu, m, o); m = (x <= m) ? m : x; }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = (red[w] <= m) ? m : red[w];
    tmax[blockIdx.x] = m;
  }
}

size_t chamfer_tc_smem_bytes(int NB) { return tc_smem_bytes(NB); }
int chamfer_tc_plan_words() { return kTcPlanWords; }


// p2: the targets the filter sweeps (Morton-sorted copy when cbox != NULL); tmax filled by the caller (chamfer_prep_launch)
// when with_bounds == 0, by chamfer_tc_bounds_kernel here otherwise.
// cbox != NULL: the pruned sweep; plan_masks (ncta x kTcPlanWords words), plan_work and plan_order (ncta ints) are scratch.
int chamfer_tc_launch(const float* p1, const float* p2, float* rbest, u64* rmask, float* cbest, u64* cmask,
                      float2* tslack, int* fallback, float* tmax, const float* cbox, const float* rbox, const float* rthr,
                      const float* cub, u64* stats, uint32_t* plan_masks, int* plan_work, int* plan_order, int with_bounds,
                      int B, int P, int M, int NB, int ntiles, int nsplit, int nchunks, int cps, cudaStream_t s) {
  static DeviceOnce once;
  const size_t smem = tc_smem_bytes(NB);
  {
    cudaError_t e = set_dyn_smem(chamfer_tc_kernel<0>, (int)tc_smem_bytes(16), once);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
      void (*kern)(const float*, const float*, float*, u64*, float*, u64*, float2*, int*, const float*, const uint32_t*, const int*, u64*,
                   int, int, int, int, int, int, int) = k == 0 ? chamfer_tc_kernel<1> : (k == 1 ? chamfer_tc_kernel<2> : (k == 2 ? chamfer_tc_kernel<3> : chamfer_tc_kernel<4>));
      static DeviceOnce more[4];
      e = set_dyn_smem(kern, (int)tc_smem_bytes(16), more[k]);
    }
    if (e != cudaSuccess) { vpn_set_error("chamfer tc: smem attribute: %s", cudaGetErrorString(e)); return VPN_ERR_CUDA; }
  }
  int rc;
  if (with_bounds) {
    chamfer_tc_bounds_kernel<<<B, 256, 0, s>>>(p2, tmax, M);
    if ((rc = vpn_check_launch("chamfer_tc_bounds_kernel"))) return rc;
  }
  const long long ncta_ll = (long long)ntiles * nsplit * B;
  if (ncta_ll > 0x7fffffffLL) { vpn_set_error("chamfer tc: too many tiles"); return VPN_ERR_SHAPE; }
  const int ncta = (int)ncta_ll;
  if (cbox != nullptr) {
    chamfer_tc_plan_kernel<<<ncta, kPlanThreads, 0, s>>>(
        cbox, rbox, rthr, cub, plan_masks, plan_work, ncta, ntiles, nsplit, (P + kTcBlk - 1) / kTcBlk, NB, nchunks, cps);
    if ((rc = vpn_check_launch("chamfer_tc_plan_kernel"))) return rc;
    chamfer_tc_order_kernel<<<1, kOrderThreads, 0, s>>>(plan_work, plan_order, ncta);
    if ((rc = vpn_check_launch("chamfer_tc_order_kernel"))) return rc;
  }
  // units (of the four of a 128-column block) reduced on the FP16 pipe, the others on the ALU pipe; vpn_set_tuning("tc_hunits",
  // 1 + units) overrides the default for probes
  const int tune_h = tuning_value(kTuneTcHunits);
  const int hunits = (tune_h >= 1 && tune_h <= 5) ? tune_h - 1 : kTcDefaultHunits;
#define VPN_TC_LAUNCH(H) chamfer_tc_kernel<H><<<ncta, kTcThreads, smem, s>>>(p1, p2, rbest, rmask, cbest, cmask, tslack, fallback, tmax, \
                                                   cbox ? plan_masks : nullptr, cbox ? plan_order : nullptr, stats, \
                                                   ntiles, nsplit, P, M, NB, nchunks, cps)
  switch (hunits) {
    case 0: VPN_TC_LAUNCH(0); break;
    case 1: VPN_TC_LAUNCH(1); break;
    case 2: VPN_TC_LAUNCH(2); break;
    case 3: VPN_TC_LAUNCH(3); break;
    default: VPN_TC_LAUNCH(4); break;
  }
#undef VPN_TC_LAUNCH
  return vpn_check_launch("chamfer_tc_kernel");
}

int chamfer_flagged_launch(const float* p1, const float* p2, float* min1, int* idx1, float* min2, int* idx2,
                           const int* fallback, int B, int P, int M, cudaStream_t s) {
  chamfer_flagged_kernel<<<dim3(32, B), 256, 0, s>>>(p1, p2, min1, idx1, min2, idx2, fallback, P, M);
  return vpn_check_launch("chamfer_flagged_kernel");
}

}  // namespace vpn
