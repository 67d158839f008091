// Probe of the operand layouts of the tensor-core phase B of scan_moments (fp32c), B200 / sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe_t umma_probe_t.cu && ./umma_probe_t
//
// The contraction of phase B is  M[moment][bin] += sum_i Phi[moment][i] * E[i][bin]  over the points i of a tile:
//   A = Phi^T stack, 128 rows (19 hi + 19 lo parts of the weighted row, the rest zero) x K points, K-major  (written by the scale step)
//   B = E tile, K points x N = 256 bins, K-major (32 points contiguous per bin row: a phase-A thread owns 4 consecutive points and
//       writes one 16-byte chunk per bin).  MN-major B (bins contiguous per point) was probed first: it works for bf16 with
//       SWIZZLE_128B, but for 32-bit operands only the SWIZZLE_128B_BASE32B layout is accepted (CUTLASS sm100_common.inl:92), whose
//       4-row atoms put a quarter-warp's stores two to a bank group.
//   D = 128 x 256 fp32 in tensor memory
// Two passes: E_hi as tf32 (kind::tf32, K = 8), E_lo = e - E_hi as bf16 (kind::f16, K = 16) against a bf16 copy of Phi_hi.
// Checked here: K-major SWIZZLE_128B (tf32) and SWIZZLE_64B (bf16, 32 points = 64-byte rows) descriptors for A and B at N = 256,
// both kinds onto one accumulator, the 16x256b fragment of tcgen05.ld, and the cycles per 32-point tile.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128(uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); }
__device__ __forceinline__ uint32_t sw64(uint32_t lin) { return lin ^ (((lin >> 7) & 3u) << 4); }
// layout: 0 none, 2 = 128B, 4 = 64B swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// c = f32, a/b format (1 bf16, 2 tf32), A K-major, B MN-major (bit 16)
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N, int b_mn) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
               :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 lanes x 256 bits: thread t holds (lane t/4, columns 2(t%4), +1) in r0, r1 and (lane t/4 + 8, same columns) in r2, r3
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int M = 128, TP = 32, NB = 256;   // rows of A, points per tile, bins
constexpr int TMEM_COLS = 512;
// shared-memory plan of one tile
constexpr int O_A32 = 0;                 // 128 rows x 128 B, K-major SW128                       16 KB
constexpr int O_E32 = 16384;             // 256 rows (bins) x 128 B (32 points), K-major SW128     32 KB
constexpr int O_A16 = 49152;             // 128 rows x 64 B, K-major SW64 (16 groups of 512 B)     8 KB
constexpr int O_E16 = 57344;             // 256 rows (bins) x 64 B (32 points), K-major SW64       16 KB
constexpr int SMEM = 73728;

// mode bit 0: tf32 pass, bit 1: bf16 pass
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A32, const float* __restrict__ E32, const __nv_bfloat16* __restrict__ A16,
                                                    const __nv_bfloat16* __restrict__ E16, float* __restrict__ D, float* __restrict__ D2, int mode, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * TP; i += 128) {                       // A32[row][point]
    const int r = i / TP, p = i % TP;
    *reinterpret_cast<float*>(smem + O_A32 + sw128((r >> 3) * 1024 + (r & 7) * 128 + p * 4)) = A32[i];
  }
  for (int i = tid; i < TP * NB; i += 128) {                      // E32[point][bin]
    const int p = i / NB, b = i % NB;
    *reinterpret_cast<float*>(smem + O_E32 + sw128((b >> 3) * 1024 + (b & 7) * 128 + p * 4)) = E32[i];
  }
  for (int i = tid; i < M * TP; i += 128) {                       // A16[row][point], 64-byte rows
    const int r = i / TP, p = i % TP;
    *reinterpret_cast<__nv_bfloat16*>(smem + O_A16 + sw64((r >> 3) * 512 + (r & 7) * 64 + p * 2)) = A16[i];
  }
  for (int i = tid; i < TP * NB; i += 128) {                      // E16[point][bin]
    const int p = i / NB, b = i % NB;
    *reinterpret_cast<__nv_bfloat16*>(smem + O_E16 + sw64((b >> 3) * 512 + (b & 7) * 64 + p * 2)) = E16[i];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem;
  if (tid == 0) {
    uint32_t acc = 0;
    const uint32_t sb = smem_u32(smem);
    if (mode & 1) {
      const uint32_t id = make_idesc(2, M, NB, 0);
      for (int kb = 0; kb < TP / 8; ++kb) {
        const uint64_t da = make_desc(sb + O_A32 + kb * 32, 16, 1024, 2);
        const uint64_t db = make_desc(sb + O_E32 + kb * 32, 16, 1024, 2);
        mma_tf32(tbase, da, db, id, acc); acc = 1;
      }
    }
    if (mode & 2) {
      const uint32_t id = make_idesc(1, M, NB, 0);
      for (int j = 0; j < TP / 16; ++j) {
        const uint64_t da = make_desc(sb + O_A16 + j * 32, 16, 512, 4);
        const uint64_t db = make_desc(sb + O_E16 + j * 32, 16, 512, 4);
        mma_f16(tbase, da, db, id, acc); acc = 1;
      }
    }
    umma_commit(smem_u32(&s_bar));
  }
  mbar_wait(smem_u32(&s_bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < NB; c0 += 16) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) D[(size_t)tid * NB + c0 + j] = v[j];
  }
  // the 16x256b shape on rows 0..15 of warp 0's quarter, columns 0..7: D2[t][0..3]
  if (warp == 0) {
    float v[4];
    tmem_ld_16x256b(tbase, v);
    for (int j = 0; j < 4; ++j) D2[tid * 4 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(TMEM_COLS) : "memory");
}

// the instruction stream of the real kernel: per 32-point tile 4 x tf32 (K = 8) + 2 x bf16 (K = 16)  [mix = 0]  or 8 x tf32 [mix = 1]
// or 4 x tf32 only [mix = 2]; two shared-memory stages and, every 16 tiles, the other half of tensor memory.  Warps 1..7 store to
// shared memory at `sts_per_tile` 16-byte stores per thread per tile meanwhile (phase A writing the next tile).
__global__ void __launch_bounds__(256) tile_rate_kernel(long long* out, int tiles, int mix, int sts_per_tile, int nb) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ volatile int s_tile;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * SMEM / 4; i += 256) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) s_tile = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem;
  if (tid == 0) {
    const uint32_t sb0 = smem_u32(smem);
    const uint32_t id32 = make_idesc(2, M, nb, 0), id16 = make_idesc(1, M, nb, 0);
    uint32_t parity[2] = {0, 0};
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const int s = t & 1;
      if (t >= 2) { mbar_wait(smem_u32(&s_bar[s]), parity[s]); parity[s] ^= 1; }     // stage s is free again
      s_tile = t;
      const uint32_t sb = sb0 + s * SMEM;
      const uint32_t d = tbase + (((t >> 4) & 1) ? 256u : 0u);
      for (int kb = 0; kb < 4; ++kb)
        mma_tf32(d, make_desc(sb + O_A32 + kb * 32, 16, 1024, 2), make_desc(sb + O_E32 + kb * 32, 16, 1024, 2), id32, 1);
      if (mix == 0)
        for (int j = 0; j < 2; ++j)
          mma_f16(d, make_desc(sb + O_A16 + j * 32, 16, 512, 4), make_desc(sb + O_E16 + j * 32, 16, 512, 4), id16, 1);
      if (mix == 1)
        for (int kb = 0; kb < 4; ++kb)
          mma_tf32(d, make_desc(sb + O_A32 + kb * 32, 16, 1024, 2), make_desc(sb + O_E32 + kb * 32, 16, 1024, 2), id32, 1);
      umma_commit(smem_u32(&s_bar[s]));
    }
    for (int s = 0; s < 2; ++s) mbar_wait(smem_u32(&s_bar[(tiles + s) & 1]), parity[(tiles + s) & 1]);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    s_tile = -1;
  } else if (warp > 0 && sts_per_tile > 0) {
    // a store stream paced by the tile counter (never touches the operand tiles: its own 32 KB region at the end)
    float4* dst = reinterpret_cast<float4*>(smem + 2 * SMEM) + (tid - 32);
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    int seen = -2;
    for (;;) {
      const int t = s_tile;
      if (t == -1) break;
      if (t == seen) continue;
      seen = t;
      for (int j = 0; j < sts_per_tile; ++j) dst[(j & 7) * 224] = v;
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "n"(TMEM_COLS) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float bf(__nv_bfloat16 h) { uint16_t b; memcpy(&b, &h, 2); uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  CK(cudaSetDevice(0));
  std::vector<float> A(M * TP), E(TP * NB);
  std::vector<__nv_bfloat16> A16(M * TP), E16(TP * NB);
  srand(2);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  for (auto& v : A) v = rnd() - 0.3f;
  for (auto& v : E) v = rnd();
  for (auto& v : A16) v = __float2bfloat16(rnd() - 0.5f);
  for (auto& v : E16) v = __float2bfloat16(rnd() - 0.5f);
  float *dA, *dE, *dD, *dD2; __nv_bfloat16 *dA16, *dE16;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dE, E.size() * 4)); CK(cudaMalloc(&dD, M * NB * 4)); CK(cudaMalloc(&dD2, 128 * 4));
  CK(cudaMalloc(&dA16, A16.size() * 2)); CK(cudaMalloc(&dE16, E16.size() * 2));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dE16, E16.data(), E16.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM + 1024));
  CK(cudaFuncSetAttribute(tile_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SMEM + 32768 + 1024));
  std::vector<float> D(M * NB), D2(128);
  for (int variant = 0; variant < 1; ++variant)
    for (int mode = 1; mode <= 3; ++mode) {
      CK(cudaMemset(dD, 0, M * NB * 4));
      probe_kernel<<<1, 128, SMEM + 1024>>>(dA, dE, dA16, dE16, dD, dD2, mode, variant);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, M * NB * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D2.data(), dD2, 128 * 4, cudaMemcpyDeviceToHost));
      double err = 0, mx = 0;
      for (int r = 0; r < M; ++r)
        for (int b = 0; b < NB; ++b) {
          double s = 0;
          if (mode & 1) for (int p = 0; p < TP; ++p) s += (double)tf32_trunc(A[r * TP + p]) * (double)tf32_trunc(E[p * NB + b]);
          if (mode & 2) for (int p = 0; p < TP; ++p) s += (double)bf(A16[r * TP + p]) * (double)bf(E16[p * NB + b]);
          err = fmax(err, fabs(D[r * NB + b] - s)); mx = fmax(mx, fabs(s));
        }
      // 16x256b shape check: thread t = (lane t/4, cols 2(t%4), +1 | lane t/4+8, same cols)
      double e2 = 0;
      for (int t = 0; t < 32; ++t) {
        const int l = t / 4, c = 2 * (t % 4);
        e2 = fmax(e2, fabs(D2[t * 4 + 0] - D[l * NB + c])); e2 = fmax(e2, fabs(D2[t * 4 + 1] - D[l * NB + c + 1]));
        e2 = fmax(e2, fabs(D2[t * 4 + 2] - D[(l + 8) * NB + c])); e2 = fmax(e2, fabs(D2[t * 4 + 3] - D[(l + 8) * NB + c + 1]));
      }
      printf("[transposed, %s%s, K-major A and B, N = 256] max|D - ref| / max|ref| = %.3e   (16x256b fragment mismatch %.1e)\n", (mode & 1) ? "tf32 SW128 " : "",
             (mode & 2) ? "bf16 SW64" : "", err / mx, e2);
    }
  long long* dT; CK(cudaMalloc(&dT, 8 * 148));
  const char* mixn[] = {"4 tf32 + 2 bf16", "8 tf32", "4 tf32"};
  for (int nb : {256, 128, 48})
    for (int mix = 0; mix < 3; ++mix)
      for (int sts : {0, 12, 48}) {
        const int tiles = 512;
        tile_rate_kernel<<<148, 256, 2 * SMEM + 32768 + 1024>>>(dT, tiles, mix, sts, nb);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        long long t[148]; CK(cudaMemcpy(t, dT, 8 * 148, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)t[i];
        avg /= 148.0 * tiles;
        printf("tile rate N=%3d bins, %-16s per 32 points, %3d STS.128/thread/tile from 7 warps: %7.1f cycles per tile = %5.1f per point\n", nb, mixn[mix], sts, avg, avg / 32.0);
      }
  printf("umma_probe_t done\n");
  return 0;
}
