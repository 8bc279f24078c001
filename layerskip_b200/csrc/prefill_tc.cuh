// prefill_tc.cuh — the prompt pass on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The reference's prefill is `forward_early` / `forward_remainder` on s = T_p rows
// (self_speculation/llama_model_utils.py:213-276, 363-383): a real contraction, unlike the decode
// steps.  The decode kernel (gemm_skinny.cuh) carries at most 16 token rows, i.e. one pass over
// the weights per 16 prompt tokens; here a pass carries 128 tokens:
//
//     out[tok, f] = sum_k act[tok, k] * W[f, k]          tok < 128 per launch, f = output feature
//
// Swap-AB UMMA: the WEIGHTS are the A operand (M = 128 output features per tile), the ACTIVATIONS
// the B operand (N = 128 tokens), both K-major in the canonical SWIZZLE_NONE core-matrix layout and
// both streamed by TMA bulk copies (16 KiB per operand per 64-wide k stage) through one shared-
// memory ring; the accumulator D[128 features x 128 tokens] fp32 lives in TMEM (128 of 512 columns,
// double-buffered: the epilogue of tile i overlaps the MMAs of tile i+1).  Weights come from HBM
// once per launch; the activation block (<= 128 x K bf16) is re-read per feature tile from L2.
// Roofline: 7B layer = 405 MB of weights / 6.5 TB/s = 62 us vs 52 GFLOP / 1.6 PFLOP/s = 33 us:
// HBM-bound at 128 tokens, i.e. the prompt costs ONE weight pass per 128 tokens instead of eight.
//
// Roles (192 threads, as lmhead_tc.cuh): warp 0 = TMA producer (one lane), warp 1 = TMEM
// allocator + single-lane tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld 32x32b: warp w
// owns TMEM lanes 32 (w % 4) .. + 31 = 32 output features, all 128 token columns).
//
// Operand layouts (common.cuh: canon_offset): K-major SWIZZLE_128B — stage (128 rows x 64 k) =
//   16 KiB, row r at r*128 B, 16-byte chunk c of a row stored at c ^ (r & 7).  Descriptor: layout
//   type 2 (SWIZZLE_128B), SBO = 1024 B (next 8-row atom), LBO unused (1); one tcgen05.mma eats
//   K = 16 = 32 bytes of a row: 4 MMAs per stage, start address advancing by 32 B.
//   Weights: [tile][k stage][16 KiB], packed once with the SAME row permutations as the decode
//   layout (rotary pairs / gate-up pairs sit 8 rows apart).  Activations: [k stage][16 KiB] with
//   rows = tokens, written in this layout by the producing kernel.
//
// Grid: one CTA per work item wave; a work item = (feature tile, k split).  Row-parallel GEMMs with
// few feature tiles (O / down projections: hidden / 128 = 32 tiles at 7B) split K so that ~all SMs
// stream weights; their fp32 partial tiles go to per-split buffers that the NEXT kernel
// (rms_canon_kernel: residual add + RMSNorm) sums in fixed order — deterministic, no atomics.
#pragma once
#include "lmhead_tc.cuh"
#include "misc_kernels.cuh"

namespace lsk {

constexpr int kPfTokens = 128;                 // UMMA N: token rows per prefill pass
constexpr int kPfStageBytes = 2 * kTcStageBytes;   // A stage + B stage
constexpr int kPfMaxStages = 6;
constexpr int kPfTmemCols = 256;               // 2 accumulators x 128 token columns

enum { PF_EPI_QKV = 0, PF_EPI_STORE = 2, PF_EPI_SILU = 3 };

struct PrefillGemmArgs {
  const unsigned char* W;      // canonical weights [n_tiles][n_kst][16 KiB]
  const unsigned char* X;      // canonical activations [n_kst][16 KiB] (rows = tokens)
  int n_tiles;                 // ceil(n_rows / 128)
  int n_rows;                  // valid output features
  int n_kst;                   // K / 64 (K padded to 64)
  int M;                       // valid token rows (<= 128)
  int n_stages;
  int k_splits;                // work items per tile (STORE only; 1 otherwise)
  // STORE: out_f32[k split][tok][f]  (split stride = kPfTokens * out_ld floats)
  float* out_f32;
  int out_ld;
  // SILU: act canonical [inter_pad / 64][16 KiB]; packed rows (16-row groups: 8 gate, 8 up)
  unsigned char* act_canon;
  // QKV: RoPE, q -> q_out natural [tok][q_ld], k/v -> paged pool
  __nv_bfloat16* q_out;
  int q_ld;
  __nv_bfloat16* kpool;
  __nv_bfloat16* vpool;
  const int* page_table;
  int pos0;                    // position of token row 0 (prefill: base length is 0)
  const float2* rope;
  int head_dim;
  int q_rows, kv_rows, n_kv_heads;
};

__host__ __device__ inline size_t prefill_tc_smem_bytes(int n_stages) {
  return (size_t)kTcHeaderBytes + (size_t)n_stages * kPfStageBytes;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start address,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B, version 1, layout type 2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__device__ __forceinline__ void tmem_alloc_cols(uint32_t* smem_dst, uint32_t cols) {      // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cols(uint32_t taddr, uint32_t cols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// natural [rows, K] bf16 -> canonical tiles with a row permutation (`mode`: misc_kernels.cuh map_row)
// and an input-column slice (tensor-parallel row-parallel GEMMs); dst rows beyond n_rows and k
// beyond K are zero.  Packed row pr of dst = source row map^-1: we iterate SOURCE rows and scatter.
__global__ void pack_canonical_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld,
                                           int64_t src_row0, int64_t src_col0, int64_t n_rows, int64_t K,
                                           unsigned char* __restrict__ dst, int64_t dst_row0, int mode,
                                           int64_t n_kst, int hd) {
  const int64_t chunks = K >> 3;                               // 16-byte chunks per source row
  const int64_t total = n_rows * chunks;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = c / chunks, kc = c % chunks;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (src_row0 + r) * src_ld + src_col0 + kc * 8);
    const int64_t pr = dst_row0 + map_row(mode, r, hd);
    const int64_t tile = pr >> 7;
    *reinterpret_cast<uint4*>(dst + (size_t)tile * n_kst * kTcStageBytes + canon_offset((int)(pr & 127), (int)(kc * 8))) = v;
  }
}

// x[tok] += sum of the `n_part` partial rows (fixed order; the split-K / tensor-parallel partials of
// the previous row-parallel GEMM), written back, then RMSNorm -> bf16 activations in the operand
// layout (the rounding point of a bf16 HF model, modeling_llama.py:52-70).  One CTA per token row.
__global__ void __launch_bounds__(256)
rms_canon_kernel(float* __restrict__ x, int x_ld, const float* __restrict__ part, int n_part, size_t part_stride,
                 const __nv_bfloat16* __restrict__ norm_w, float eps, int K, unsigned char* __restrict__ dst) {
  __shared__ float s_part[8];
  pdl_launch_dependents();
  pdl_wait();
  const int tok = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)tok * x_ld);
  const int nvec = K >> 2;
  float ss = 0.f;
  for (int i = tid; i < nvec; i += 256) {
    float4 v = xr[i];
    for (int p = 0; p < n_part; ++p) {
      const float4 d = *reinterpret_cast<const float4*>(part + (size_t)p * part_stride + (size_t)tok * x_ld + i * 4);
      v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
    }
    if (n_part > 0) xr[i] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if (lane == 0) s_part[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += s_part[w];
  const float rstd = rsqrtf(tot / (float)K + eps);
  if (dst == nullptr) return;                       // residual update only
  for (int i = tid; i < nvec; i += 256) {
    const float4 v = xr[i];                         // own writes: same thread, same address
    const uint2 wv = *reinterpret_cast<const uint2*>(norm_w + i * 4);
    uint2 o;
    o.x = pack_bf16x2(bf16_lo(wv.x) * (v.x * rstd), bf16_hi(wv.x) * (v.y * rstd));
    o.y = pack_bf16x2(bf16_lo(wv.y) * (v.z * rstd), bf16_hi(wv.y) * (v.w * rstd));
    *reinterpret_cast<uint2*>(dst + canon_offset(tok, i * 4)) = o;
  }
}

// tensor parallel: out[tok] = sum of the local split-K partials (fixed order), all-reduced afterwards
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int n_part, size_t part_stride, int ld, int K,
                       float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int tok = blockIdx.x;
  for (int i = threadIdx.x; i < (K >> 2); i += 256) {
    float4 v = *reinterpret_cast<const float4*>(part + (size_t)tok * ld + i * 4);
    for (int p = 1; p < n_part; ++p) {
      const float4 d = *reinterpret_cast<const float4*>(part + (size_t)p * part_stride + (size_t)tok * ld + i * 4);
      v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
    }
    *reinterpret_cast<float4*>(out + (size_t)tok * ld + i * 4) = v;
  }
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
prefill_gemm_tc_kernel(const PrefillGemmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kPfMaxStages;
  uint64_t* tfull_bar = empty_bar + kPfMaxStages;            // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  unsigned char* ring = smem + kTcHeaderBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.n_stages;
  const int KS = (EPI == PF_EPI_STORE && a.k_splits > 1) ? a.k_splits : 1;
  const int n_items = a.n_tiles * KS;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_cols(tmem_slot, kPfTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  pdl_wait();          // the activations (B operand) are the previous kernel's output

  if (warp == 0) {
    if (lane == 0) {
      // ============================================================ TMA PRODUCER
      uint32_t q = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / KS, ks = item - tile * KS;
        const int s_lo = (int)((long long)a.n_kst * ks / KS), s_hi = (int)((long long)a.n_kst * (ks + 1) / KS);
        for (int s = s_lo; s < s_hi; ++s, ++q) {
          const int st = q % NS;
          mbar_wait_bounded(&empty_bar[st], ((q / NS) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[st], kPfStageBytes);
          unsigned char* dst = ring + (size_t)st * kPfStageBytes;
          tma_bulk_g2s(dst, a.W + ((size_t)tile * a.n_kst + s) * kTcStageBytes, kTcStageBytes, &full_bar[st]);
          tma_bulk_g2s(dst + kTcStageBytes, a.X + (size_t)s * kTcStageBytes, kTcStageBytes, &full_bar[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA ISSUER (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcTileRows, kPfTokens);
      uint32_t q = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int tile = item / KS, ks = item - tile * KS;
        const int s_lo = (int)((long long)a.n_kst * ks / KS), s_hi = (int)((long long)a.n_kst * (ks + 1) / KS);
        (void)tile;
        const int buf = it & 1;
        mbar_wait_bounded(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);      // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)buf * kPfTokens;
        for (int s = s_lo; s < s_hi; ++s, ++q) {
          const int st = q % NS;
          mbar_wait_bounded(&full_bar[st], (q / NS) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)st * kPfStageBytes);
          const uint32_t b_addr = a_addr + kTcStageBytes;
#pragma unroll
          for (int k = 0; k < kTcStageK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            umma_bf16_ss(d_addr, da, db, idesc, (s > s_lo || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[st]);               // frees the ring slot once these MMAs have read it
        }
        umma_commit(&tfull_bar[buf]);                // accumulator of this tile complete
      }
    }
  } else {
    // ================================================================ EPILOGUE (4 warps x 32 features)
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / KS, ks = item - tile * KS;
      const int buf = it & 1;
      mbar_wait_bounded(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const int prow = tile * kTcTileRows + quarter * 32 + lane;       // packed output row (feature)
      const bool valid = prow < a.n_rows;
      const int r16 = prow & 15;
#pragma unroll 1
      for (int c0 = 0; c0 < kPfTokens; c0 += 32) {
        if (c0 >= a.M) break;                        // warp-uniform: no token rows beyond M
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kPfTokens + c0), v);
        if (EPI == PF_EPI_STORE) {
          float* outp = a.out_f32 + (size_t)ks * kPfTokens * a.out_ld + prow;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int tok = c0 + j;
            if (tok < a.M && valid) outp[(size_t)tok * a.out_ld] = __uint_as_float(v[j]);
          }
        } else if (EPI == PF_EPI_SILU) {
          // 16-row groups: rows 0..7 gate, rows 8..15 up of the same 8 features (MAP_GATE / MAP_UP)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float mine = __uint_as_float(v[j]);
            const float other = __shfl_xor_sync(0xffffffffu, mine, 8);
            const int tok = c0 + j;
            if (r16 < 8 && tok < a.M && valid) {
              const float sg = mine / (1.f + __expf(-mine));
              const int kidx = (prow >> 4) * 8 + r16;
              *reinterpret_cast<__nv_bfloat16*>(a.act_canon + canon_offset(tok, kidx)) = __float2bfloat16_rn(sg * other);
            }
          }
        } else {  // PF_EPI_QKV
          // everything that depends only on the output row is computed once per thread; the RoPE
          // factors of the 32 tokens are fetched up front (independent loads, one round trip)
          const int HD = a.head_dim, half = HD >> 1;
          const bool is_q = prow < a.q_rows, is_k = !is_q && prow < a.q_rows + a.kv_rows;
          const int rel = is_q ? prow : (is_k ? prow - a.q_rows : prow - a.q_rows - a.kv_rows);
          const int head = rel / HD, inhead = rel - head * HD;
          const bool upper = r16 >= 8;                              // upper row of a rotary pair
          const int d = (inhead >> 4) * 8 + (r16 & 7);              // pair index (q / k rows)
          const int dd = (is_q || is_k) ? (upper ? d + half : d) : inhead;
          const float2* __restrict__ rope = a.rope;
          float2 cs[32];
          if (is_q || is_k) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int tok = min(c0 + j, a.M - 1);
              cs[j] = __ldg(rope + (size_t)(a.pos0 + tok) * half + d);
            }
          }
          const int page_lo = a.page_table[(a.pos0 + c0) >> 6];
          const int page_hi = a.page_table[(a.pos0 + min(c0 + 31, a.M - 1)) >> 6];
          __nv_bfloat16* __restrict__ qo = a.q_out + head * HD + dd;
          __nv_bfloat16* __restrict__ pool = is_k ? a.kpool : a.vpool;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float mine = __uint_as_float(v[j]);
            const float other = __shfl_xor_sync(0xffffffffu, mine, 8);
            const int tok = c0 + j;
            if (tok >= a.M || !valid) continue;
            const int pos = a.pos0 + tok;
            // lower row of the pair holds x[d] (lo), upper row x[d + half] (hi):
            //   out_lo = lo cos - hi sin,  out_hi = hi cos + lo sin   (rotate_half, modeling_llama.py:138-168)
            float outv = mine;
            if (is_q || is_k) outv = upper ? mine * cs[j].x + other * cs[j].y : mine * cs[j].x - other * cs[j].y;
            const __nv_bfloat16 ob = __float2bfloat16_rn(outv);
            if (is_q) {
              qo[(size_t)tok * a.q_ld] = ob;
            } else {
              const int page = ((pos >> 6) == ((a.pos0 + c0) >> 6)) ? page_lo : page_hi;
              pool[kv_elem_offset(HD, page, a.n_kv_heads, head, pos & 63, dd)] = ob;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);  // the MMA warp may overwrite this buffer
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_cols(tmem_base, kPfTmemCols);
}

}  // namespace lsk
