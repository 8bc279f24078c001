// lmhead_tc.cuh — LM head on the 5th-generation tensor cores (tcgen05 + TMEM), OPT-IN
// (LSK_LMHEAD_TC=1).  Written after the round-1 GPU budget was spent: compiled for sm_100a, SASS
// inspected (UTCHMMA / LDTM / UBLKCP), NEVER EXECUTED — every wait is bounded and traps instead of
// hanging.  The default LM head is gemm_skinny_kernel<NT, PRO_RMS, EPI_LMHEAD> (mma.sync).
//
//   logits[v, m] = sum_k Wlm[v, k] * rmsnorm(x)[m, k]        v: local vocab rows, m <= 16 tokens
//
// This is the one GEMM of the decode path whose N side is large enough (vocab >= 32000: 250+
// tiles of 128 rows) for a 128-row UMMA tile to keep every SM busy without split-K, so it is where
// tcgen05 fits (DESIGN.md §3.1 explains why the layer GEMMs stay on 16-row mma.sync fragments).
// Swap-AB: the WEIGHTS are the A operand (M = 128 vocab rows, K-major, streamed by TMA bulk copies
// into a shared-memory ring), the normalised activations are the B operand (N = 16 token columns,
// K-major, resident in shared memory), the accumulator D[128 x 16] fp32 lives in TMEM (16 of 512
// columns, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
//
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one
// lane issues tcgen05.mma / tcgen05.commit), warps 2..5 = epilogue (tcgen05.ld 32x32b: warp w
// owns TMEM lanes 32 (w % 4) .., i.e. 32 vocab rows of the tile; optional logits store + running
// arg-max with the engine's "lowest index wins" rule).  Warps 1..5 share the RMSNorm prologue.
//
// Shared-memory operand layouts (SWIZZLE_NONE canonical K-major: 8 rows x 16 B core matrices):
//   A stage  = 128 rows x 64 k = 16 KiB, packed by pack_canonical_kernel exactly as it lies in
//              HBM: core(row group i = 0..15, k chunk j = 0..7) at (i * 8 + j) * 128 B
//              -> descriptor LBO = 128 B (next k chunk), SBO = 1024 B (next 8 rows)
//   B (whole K) = core(k chunk j, token group i = 0..1) at (j * 2 + i) * 128 B
//              -> descriptor LBO = 256 B, SBO = 128 B
//   one tcgen05.mma consumes K = 16 (two k chunks): per stage 4 MMAs, descriptors advance by
//   256 B (A) and 512 B (B).
#pragma once
#include "gemm_skinny.cuh"

namespace lsk {

constexpr int kTcThreads = 192;
constexpr int kTcTileRows = 128;
constexpr int kTcStageK = 64;
constexpr int kTcStageBytes = kTcTileRows * kTcStageK * 2;   // 16 KiB
constexpr int kTcTokens = 16;                                // UMMA N
constexpr int kTcMaxStages = 6;
constexpr int BAR_TC_EPI = 9;                                // named barrier of the 4 epilogue warps
constexpr int BAR_TC_PRO = 10;                               // named barrier of warps 1..5 (prologue)
constexpr int kTcProThreads = kTcThreads - 32;               // everyone but the producer warp
constexpr int kTcHeaderBytes = 2048;                         // mbarriers, TMEM slot, small scratch
constexpr long long kTcTimeoutCycles = 2000000000LL;         // ~1 s: trap instead of hanging

struct LmHeadTcArgs {
  const unsigned char* W;     // canonical-packed weights [n_tiles][K / 64][16 KiB]
  int n_tiles;                // ceil(local vocab / 128)
  int K;                      // hidden (multiple of 64)
  int M;                      // valid token rows (<= 16)
  int n_stages;               // ring depth
  const float* x_f32;         // residual rows [M][x_ld]
  int x_ld;
  const __nv_bfloat16* norm_w;
  float eps;
  float* logits;              // optional [M][logits_ld]
  int logits_ld;
  int n_valid_rows;           // local vocab rows
  int vocab_off;              // global id of local row 0
  float* part_val;            // [grid][16]
  int* part_idx;
};

__host__ __device__ inline size_t lmhead_tc_smem_bytes(int K, int n_stages) {
  return (size_t)kTcHeaderBytes + (size_t)n_stages * kTcStageBytes + (size_t)kTcTokens * K * 2;
}

// ---- tcgen05 / TMEM primitives (PTX strings as in CUTLASS cute/arch/{mma_sm100_umma,copy_sm100,
//      tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc32(uint32_t* smem_dst) {      // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(32u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc32(uint32_t taddr) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(32u) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {            // arrives when prior MMAs finish
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// SWIZZLE_NONE K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);     // version 1, layout type 0
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// bounded mbarrier wait: a protocol bug must surface as a launch failure, not as a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && clock64() - t0 > kTcTimeoutCycles) __trap();
  }
}

// natural [rows, K] bf16 -> canonical tiles (see header); rows >= n_rows are zero
__global__ void pack_canonical_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld, int64_t row0,
                                      int64_t n_rows, int64_t K, uint4* __restrict__ dst, int64_t n_tiles) {
  const int64_t kst = K / kTcStageK;
  const int64_t total = n_tiles * kst * (kTcStageBytes / 16);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t blk = c / (kTcStageBytes / 16);          // (tile, k stage)
    const int in = (int)(c % (kTcStageBytes / 16));        // 16-byte chunk inside the stage
    const int64_t tile = blk / kst, s = blk % kst;
    const int core = in >> 3, r = in & 7;                  // core = i * 8 + j
    const int i = core >> 3, j = core & 7;
    const int64_t row = tile * kTcTileRows + i * 8 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < n_rows) v = *reinterpret_cast<const uint4*>(src + (row0 + row) * src_ld + s * kTcStageK + j * 8);
    dst[c] = v;
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
lmhead_tc_kernel(const LmHeadTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kTcMaxStages;
  uint64_t* tfull_bar = empty_bar + kTcMaxStages;           // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                     // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* stat = reinterpret_cast<float*>(smem + 256);       // [6 warps][16] + rstd[16]   (448 B)
  float* xval = reinterpret_cast<float*>(smem + 768);       // [4 warps][16] cross-warp arg-max
  int* xi = reinterpret_cast<int*>(smem + 1024);            // [4 warps][16]
  unsigned char* ring = smem + kTcHeaderBytes;
  unsigned char* xb = ring + (size_t)a.n_stages * kTcStageBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.n_stages;
  const int n_kst = a.K / kTcStageK;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc32(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
    // ============================================================ TMA PRODUCER (weights are static:
    // it runs ahead of the PDL dependency, like gemm_producer)
    uint32_t q = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
      for (int s = 0; s < n_kst; ++s, ++q) {
        const int st = q % NS;
        mbar_wait_bounded(&empty_bar[st], ((q / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[st], kTcStageBytes);
        tma_bulk_g2s(ring + (size_t)st * kTcStageBytes,
                     a.W + ((size_t)tile * n_kst + s) * kTcStageBytes, kTcStageBytes, &full_bar[st]);
      }
    }
    pdl_wait();                                      // completion stays transitive along the PDL chain
  } else {
    pdl_wait();
    // -------------------------------------------------------------- prologue (warps 1..5): RMSNorm
    // of the token rows -> bf16 B operand in the canonical K-major layout.  The producer warp must
    // NOT take part: its progress depends on the MMAs, which depend on this prologue.
    const int ptid = tid - 32, pwarp = warp - 1;
    const int nvec = a.K >> 2;
    for (int m = 0; m < a.M; ++m) {
      const float4* xr = reinterpret_cast<const float4*>(a.x_f32 + (size_t)m * a.x_ld);
      float ss = 0.f;
      for (int idx = ptid; idx < nvec; idx += kTcProThreads) {
        const float4 v = xr[idx];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      ss = warp_sum(ss);
      if (lane == 0) stat[pwarp * 16 + m] = ss;
    }
    bar_sync(BAR_TC_PRO, kTcProThreads);
    if (ptid < a.M) {
      float tot = 0.f;
      for (int w = 0; w < kTcProThreads / 32; ++w) tot += stat[w * 16 + ptid];
      stat[6 * 16 + ptid] = rsqrtf(tot / (float)a.K + a.eps);
    }
    bar_sync(BAR_TC_PRO, kTcProThreads);
    for (int m = 0; m < kTcTokens; ++m) {
      const float rstd = m < a.M ? stat[6 * 16 + m] : 0.f;
      const float4* xr = reinterpret_cast<const float4*>(a.x_f32 + (size_t)(m < a.M ? m : 0) * a.x_ld);
      for (int idx = ptid; idx < nvec; idx += kTcProThreads) {   // 4 consecutive k: half a 16-B chunk
        uint2 o = make_uint2(0u, 0u);
        if (m < a.M) {
          const float4 v = xr[idx];
          const uint2 wv = *reinterpret_cast<const uint2*>(a.norm_w + idx * 4);
          o.x = pack_bf16x2(bf16_lo(wv.x) * (v.x * rstd), bf16_hi(wv.x) * (v.y * rstd));
          o.y = pack_bf16x2(bf16_lo(wv.y) * (v.z * rstd), bf16_hi(wv.y) * (v.w * rstd));
        }
        const int k = idx * 4, j = k >> 3;
        *reinterpret_cast<uint2*>(xb + ((size_t)j * 2 + (m >> 3)) * 128 + (m & 7) * 16 + (k & 7) * 2) = o;
      }
    }
    // generic-proxy writes -> visible to the tensor core's async-proxy reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bar_sync(BAR_TC_PRO, kTcProThreads);
  }

  if (warp == 1) {
    // ============================================================ MMA ISSUER (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcTileRows, kTcTokens);
      const uint32_t xb_addr = smem_u32(xb);
      uint32_t q = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_bounded(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);      // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)buf * kTcTokens;
        for (int s = 0; s < n_kst; ++s, ++q) {
          const int st = q % NS;
          mbar_wait_bounded(&full_bar[st], (q / NS) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)st * kTcStageBytes);
#pragma unroll
          for (int k = 0; k < kTcStageK / 16; ++k) {
            const uint64_t da = umma_desc(a_addr + k * 256, 128, 1024);
            const uint64_t db = umma_desc(xb_addr + (uint32_t)(s * 8 + k * 2) * 256, 256, 128);
            umma_bf16_ss(d_addr, da, db, idesc, (s > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[st]);               // frees the ring slot once these MMAs have read it
        }
        umma_commit(&tfull_bar[buf]);                // accumulator of this tile complete
      }
    }
  } else if (warp >= 2) {
    // ============================================================ EPILOGUE (4 warps, 128 vocab rows)
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    float best_v[kTcTokens];
    int best_i[kTcTokens];
#pragma unroll
    for (int m = 0; m < kTcTokens; ++m) { best_v[m] = -INFINITY; best_i[m] = 0x7fffffff; }
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait_bounded(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kTcTokens, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);  // the MMA warp may overwrite this buffer
      const int orow = tile * kTcTileRows + quarter * 32 + lane;
      const bool valid = orow < a.n_valid_rows;
#pragma unroll
      for (int m = 0; m < kTcTokens; ++m) {
        if (m < a.M) {
          const float val = __uint_as_float(v[m]);
          if (valid && a.logits != nullptr) a.logits[(size_t)m * a.logits_ld + orow] = val;
          float bv = valid ? val : -INFINITY;
          int bi = valid ? orow : 0x7fffffff;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
          }
          if (better(bv, bi, best_v[m], best_i[m])) { best_v[m] = bv; best_i[m] = bi; }
        }
      }
    }
    // cross-warp: 4 candidates per token -> one (value, global index) per CTA
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < kTcTokens; ++m) { xval[(warp - 2) * kTcTokens + m] = best_v[m]; xi[(warp - 2) * kTcTokens + m] = best_i[m]; }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(BAR_TC_EPI), "r"(128) : "memory");
    if (warp == 2 && lane < a.M) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int w = 0; w < 4; ++w)
        if (better(xval[w * kTcTokens + lane], xi[w * kTcTokens + lane], bv, bi)) {
          bv = xval[w * kTcTokens + lane];
          bi = xi[w * kTcTokens + lane];
        }
      a.part_val[blockIdx.x * kMaxRows + lane] = bv;
      a.part_idx[blockIdx.x * kMaxRows + lane] = (bi == 0x7fffffff) ? 0x7fffffff : bi + a.vocab_off;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc32(tmem_base);
}

}  // namespace lsk
