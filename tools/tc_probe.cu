// tcgen05 probe for the Chamfer tensor-core filter (B200, sm_100a):
//   test 1  correctness of D = A * B^T (kind::tf32, M=128, N=128, K=16 as two K=8 steps) with both operands in
//           shared memory in the canonical K-major no-swizzle layout, read back with tcgen05.ld.32x32b;
//   test 2  tensor time of one "block" of the filter (two 128x128x16 products) issued back to back;
//   test 3  epilogue rate: tcgen05.ld of 128 columns per thread + 128 min ops, 8 warps per CTA;
//   test 4  accumulation error of the 3xTF32 split distance against the exact fp32 value.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;                       // descriptor version (Blackwell)
  return d;                              // layout type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);       // both K-major
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(COLS) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

// operand tile: 128 rows x 16 tf32, canonical K-major no-swizzle: 8-row groups of 512 B, each made of four
// 128-B core matrices (8 rows x 16 B), one per group of 4 consecutive k.
__device__ __forceinline__ void store_row(float* tile, int row, const float (&x)[16]) {
  unsigned char* base = reinterpret_cast<unsigned char*>(tile) + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4)
    *reinterpret_cast<float4*>(base + k4 * 128) = make_float4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]);
}
constexpr uint32_t kLBO = 128, kSBO = 512, kKStepBytes = 256;

__device__ __forceinline__ float tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// ---- test 1 / 4: one CTA, D (128x128) = A (128x16) * B (128x16)^T --------------------------------------
__global__ void __launch_bounds__(128) k_correct(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + 8192);
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    float x[16];
    for (int k = 0; k < 16; ++k) x[k] = A[tid * 16 + k];
    store_row(sA, tid, x);
    for (int k = 0; k < 16; ++k) x[k] = B[tid * 16 + k];
    store_row(sB, tid, x);
  }
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 128);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    mma_tf32(tbase, make_desc(a0, kLBO, kSBO), make_desc(b0, kLBO, kSBO), idesc, 0);
    mma_tf32(tbase, make_desc(a0 + kKStepBytes, kLBO, kSBO), make_desc(b0 + kKStepBytes, kLBO, kSBO), idesc, 1);
    mma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after();
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_wait_ld();
    for (int k = 0; k < 32; ++k) D[(warp * 32 + lane) * 128 + c + k] = v[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<128>(tbase);
}

// ---- test 2: tensor time per filter block (two 128x128x16 products) --------------------------------------
__global__ void __launch_bounds__(128) k_mma_rate(float* out, int iters, int per_commit) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + 8192);
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  { float x[16]; for (int k = 0; k < 16; ++k) x[k] = 1.0f / (1 + tid + k); store_row(sA, tid, x); store_row(sB, tid, x); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_slot;
  long long t0 = clock64();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 128);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint64_t da0 = make_desc(a0, kLBO, kSBO), da1 = make_desc(a0 + kKStepBytes, kLBO, kSBO);
    const uint64_t db0 = make_desc(b0, kLBO, kSBO), db1 = make_desc(b0 + kKStepBytes, kLBO, kSBO);
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < per_commit; ++j) {
        const uint32_t d = tbase + ((it * per_commit + j) & 1) * 256;
        mma_tf32(d, da0, db0, idesc, 0); mma_tf32(d, da1, db1, idesc, 1);                    // D1 = rows x cols
        mma_tf32(d + 128, db0, da0, idesc, 0); mma_tf32(d + 128, db1, da1, idesc, 1);        // D2 = cols x rows
      }
      mma_commit(smem_u32(&mbar));
      mbar_wait(smem_u32(&mbar), phase); phase ^= 1;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / ((float)iters * per_commit);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free<512>(tbase);
}

// ---- test 3: epilogue rate: every thread reads 128 TMEM columns of its lane and takes the minimum ----------
template <int MINOP>   // 0: ld only, 1: ld + FMNMX3, 2: ld + FMNMX
__global__ void __launch_bounds__(256) k_epi_rate(float* out, int iters) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float a[32], b[32], c[32], d[32];
    const uint32_t t = taddr + (it & 1) * 256;
    tmem_ld32(t, a); tmem_ld32(t + 32, b); tmem_ld32(t + 64, c); tmem_ld32(t + 96, d);
    tmem_wait_ld();
    if (MINOP == 1) {
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(a[k]), "f"(a[k + 1]));
        asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(b[k]), "f"(b[k + 1]));
        asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m2) : "f"(c[k]), "f"(c[k + 1]));
        asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m3) : "f"(d[k]), "f"(d[k + 1]));
      }
    } else if (MINOP == 2) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m0) : "f"(a[k]));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m1) : "f"(b[k]));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m2) : "f"(c[k]));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m3) : "f"(d[k]));
      }
    } else {
      m0 = fminf(m0, a[0]); m1 = fminf(m1, b[1]); m2 = fminf(m2, c[2]); m3 = fminf(m3, d[3]);
    }
  }
  long long t1 = clock64();
  float m = fminf(fminf(m0, m1), fminf(m2, m3));
  if (m == 123.456f) out[1] = m;
  if (tid == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / (float)iters;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free<512>(tbase);
}

// test 6: stream of MMAs issued the way the product kernel does (whole warp walks the loop, one elected lane issues):
// `per` MMAs of shape 128 x N x 8 then one commit, `steps` times; never waits except at the end.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
template <int N>
__global__ void __launch_bounds__(320) k_mma_stream(float* out, int steps, int per, int readers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar[5];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int t = 0; t < 18; ++t) if (tid < 128) { float x[16]; for (int k = 0; k < 16; ++k) x[k] = 1.0f / (1 + tid + k + t); store_row(reinterpret_cast<float*>(smem + t * 8192), tid, x); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) { for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&mbar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 9) {
    constexpr uint32_t idesc = make_idesc(128, N);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t da = make_desc(s0, kLBO, kSBO), db = make_desc(s0 + 16 * 8192, kLBO, kSBO);
    const uint32_t m0 = smem_u32(&mbar[0]);
    for (int it = 0; it < steps; ++it) {
      if (elect_one()) {
        for (int j = 0; j < per; ++j) {
          const uint64_t a = da + (uint64_t)(((it * per + j) & 7) * 512);
          const uint32_t d = tbase + (N == 256 ? ((it * per + j) & 1) * 256 : ((it * per + j) & 3) * 128);
          mma_tf32(d, a, db, idesc, 0);
          mma_tf32(d, a + 16, db + 16, idesc, 1);
        }
        mma_commit(m0 + 8 * (it & 3));
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(m0 + 32);
    __syncwarp();
    mbar_wait(m0 + 32, 0);
  } else if (readers && warp < 8) {
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    float mm = 1e30f;
    for (int it = 0; it < steps * per / 2; ++it) {
      float a[32], b[32];
      tmem_ld32(taddr + (it & 3) * 128, a); tmem_ld32(taddr + (it & 3) * 128 + 32, b);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 32; k += 2) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(mm) : "f"(a[k]), "f"(a[k + 1])); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(mm) : "f"(b[k]), "f"(b[k + 1])); }
    }
    if (mm == 123.456f) out[1] = mm;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free<512>(tbase);
}

template <int N>
static void run_stream(float* dout, int sms, int per, int readers) {
  const int steps = 8000 / per;
  CK(cudaFuncSetAttribute(k_mma_stream<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 8192));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_mma_stream<N><<<sms, 320, 18 * 8192>>>(dout, 16, per, readers); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_mma_stream<N><<<sms, 320, 18 * 8192>>>(dout, steps, per, readers); CK(cudaGetLastError());
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double clk = ms * 1e-3 * 1.965e9 / (steps * per);
  printf("test6 MMA stream N=%d, %d accumulator(s) per commit, TMEM readers %d: %.1f clk per accumulator (2 x 128x%dx8) = %.1f clk per 128x128x16\n",
         N, per, readers, clk, N, clk * 128 / N);
}

// 16 warps: each thread reads 128 columns as two halves of 64, the second half's loads in flight while the first is reduced
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_epi_rate2(float* out, int iters) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  float m0 = 1e30f, m1 = 1e30f;
  long long t0 = clock64();
  float a[32], b[32];
  tmem_ld32(taddr, a); tmem_ld32(taddr + 32, b);
  for (int it = 0; it < iters; ++it) {
    float c[32], d[32];
    tmem_wait_ld();
    tmem_ld32(taddr + 64, c); tmem_ld32(taddr + 96, d);
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(a[k]), "f"(a[k + 1]));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(b[k]), "f"(b[k + 1]));
    }
    tmem_wait_ld();
    tmem_ld32(taddr, a); tmem_ld32(taddr + 32, b);
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(c[k]), "f"(c[k + 1]));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(d[k]), "f"(d[k + 1]));
    }
  }
  tmem_wait_ld();
  long long t1 = clock64();
  float m = fminf(m0, m1) + a[0] + b[0];
  if (m == 123.456f) out[1] = m;
  if (tid == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / (float)iters;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_free<512>(tbase);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  CK(cudaFuncSetAttribute(k_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  CK(cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  float *dA, *dB, *dD, *dout;
  CK(cudaMalloc(&dA, 128 * 16 * 4)); CK(cudaMalloc(&dB, 128 * 16 * 4)); CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&dout, 64));
  std::vector<float> A(128 * 16), B(128 * 16), D(128 * 128);

  // test 1: small integers (exact in tf32, exact sums): any layout / descriptor mistake shows as a mismatch
  for (int i = 0; i < 128; ++i) for (int k = 0; k < 16; ++k) { A[i * 16 + k] = (float)((i * 7 + k * 3) % 11 - 5); B[i * 16 + k] = (float)((i * 5 + k * 2) % 13 - 6); }
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  k_correct<<<1, 128, 16384>>>(dA, dB, dD); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
    float ref = 0.f; for (int k = 0; k < 16; ++k) ref += A[i * 16 + k] * B[j * 16 + k];
    if (ref != D[i * 128 + j]) { if (bad < 5) printf("  mismatch D[%d][%d] = %g, want %g\n", i, j, D[i * 128 + j], ref); ++bad; }
  }
  printf("test1 integer product: %d mismatches of 16384\n", bad);

  // test 4: 3xTF32 split distance vs exact fp32 (centred data, |p| <= rho, |t| <= 1)
  {
    srand(12345);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };  // rna to 10 bits
    std::vector<float> P(128 * 3), T(128 * 3);
    const float rho = 0.12f;
    for (int i = 0; i < 128; ++i) for (int c = 0; c < 3; ++c) { P[i * 3 + c] = rnd() * rho; T[i * 3 + c] = rnd() * (i < 64 ? 0.15f : 1.0f); }
    for (int i = 0; i < 128; ++i) {
      float a[3], s = 0.f, c2 = 0.f;
      for (int c = 0; c < 3; ++c) { a[c] = -2.f * P[i * 3 + c]; }
      s = fmaf(P[i * 3 + 2], P[i * 3 + 2], fmaf(P[i * 3 + 1], P[i * 3 + 1], P[i * 3] * P[i * 3]));
      c2 = fmaf(T[i * 3 + 2], T[i * 3 + 2], fmaf(T[i * 3 + 1], T[i * 3 + 1], T[i * 3] * T[i * 3]));
      float* r = &A[i * 16]; float* q = &B[i * 16];
      for (int c = 0; c < 3; ++c) {
        float ah = tf(a[c]), al = tf(a[c] - ah), th = tf(T[i * 3 + c]), tl = tf(T[i * 3 + c] - th);
        r[3 * c] = ah; r[3 * c + 1] = ah; r[3 * c + 2] = al; q[3 * c] = th; q[3 * c + 1] = tl; q[3 * c + 2] = th;
        r[13 + c] = al; q[13 + c] = tl;
      }
      float sh = tf(s), sl = tf(s - sh), ch = tf(c2), cl = tf(c2 - ch);
      r[9] = sh; r[10] = sl; r[11] = 1.f; r[12] = 1.f; q[9] = 1.f; q[10] = 1.f; q[11] = ch; q[12] = cl;
    }
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    k_correct<<<1, 128, 16384>>>(dA, dB, dD); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst_rel = 0;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
      double dx = (double)P[i * 3] - T[j * 3], dy = (double)P[i * 3 + 1] - T[j * 3 + 1], dz = (double)P[i * 3 + 2] - T[j * 3 + 2];
      double ex = dx * dx + dy * dy + dz * dz;
      double np = sqrt((double)P[i * 3] * P[i * 3] + (double)P[i * 3 + 1] * P[i * 3 + 1] + (double)P[i * 3 + 2] * P[i * 3 + 2]);
      double nt = sqrt((double)T[j * 3] * T[j * 3] + (double)T[j * 3 + 1] * T[j * 3 + 1] + (double)T[j * 3 + 2] * T[j * 3 + 2]);
      double err = fabs((double)D[i * 128 + j] - ex), scale = (np + nt) * (np + nt);
      if (err / scale > worst) worst = err / scale;
      if (ex > 0 && err / ex > worst_rel) worst_rel = err / ex;
    }
    printf("test4 3xTF32 distance: max |e - d| / (|p|+|t|)^2 = %.3e (= 2^%.2f), max relative to d = %.3e\n", worst, log2(worst), worst_rel);
  }

  // test 2
  for (int pc : {1, 2, 4}) {
    k_mma_rate<<<sms, 128, 16384>>>(dout, 2000, pc); CK(cudaDeviceSynchronize());
    float clk; CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test2 tensor clk per block (2 x 128x128x16 tf32), %d block(s) per commit: %.1f clk  -> %.1f pairs/clk/SM\n", pc, clk, 16384.0 / clk);
  }
  // test 3
  {
    float clk;
    k_epi_rate<0><<<sms, 256>>>(dout, 4000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test3 epilogue, tcgen05.ld only        : %.1f clk per 128 columns per thread (8 warps)\n", clk);
    k_epi_rate<1><<<sms, 256>>>(dout, 4000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test3 epilogue, ld + 64 FMNMX3         : %.1f clk\n", clk);
    k_epi_rate<2><<<sms, 256>>>(dout, 4000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test3 epilogue, ld + 128 FMNMX         : %.1f clk\n", clk);
  }
  run_stream<128>(dout, sms, 1, 0); run_stream<128>(dout, sms, 4, 0); run_stream<128>(dout, sms, 1, 1);
  run_stream<256>(dout, sms, 1, 0); run_stream<256>(dout, sms, 2, 0); run_stream<64>(dout, sms, 1, 0); run_stream<64>(dout, sms, 4, 0);
  {
    float clk;
    k_epi_rate2<8><<<sms, 256>>>(dout, 4000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test5 pipelined epilogue, 8 warps x 128 columns : %.1f clk per iteration (= per block)\n", clk);
    k_epi_rate2<16><<<sms, 512>>>(dout, 4000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&clk, dout, 4, cudaMemcpyDeviceToHost));
    printf("test5 pipelined epilogue, 16 warps x 128 columns: %.1f clk per iteration (= per 2 blocks)\n", clk);
  }
  return 0;
}
