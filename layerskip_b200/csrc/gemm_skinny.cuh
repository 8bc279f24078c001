// gemm_skinny.cuh — weight-streaming "skinny" GEMM for decode shapes:
//     y[m, n] = sum_k x[m, k] * W[n, k]        m <= 16 tokens,  W bf16 [N, K] (HF layout)
// HBM-bound by construction (arithmetic intensity <= 16 flop/B): the kernel is a stream of
// fully coalesced 128-bit loads of PRE-SHUFFLED weights straight into mma.sync A-fragments,
// with the (tiny) activation block resident in shared memory as the B operand.
//
// Packed weight layout (built once by pack_rows_kernel, see engine.cu):
//   tile   = 16 consecutive output rows,  super-block (sb) = 32 consecutive k
//   for each (tile, sb): 2 MMAs x 32 lanes x 16 B   = 1 KiB, contiguous, tile-major
//   lane (g = lane/4, t = lane%4), MMA j holds  a0..a3 =
//        W[g   ][32sb + 8t + 4j + {0,1}],  W[g+8][.. same ..],
//        W[g   ][32sb + 8t + 4j + {2,3}],  W[g+8][.. same ..]
//   i.e. the physical k order inside a super-block is permuted so that a lane's B operand for
//   both MMAs is ONE 128-bit shared-memory load of x[token g][32sb + 8t .. 8t+7].
//
// Numerics are batch-invariant: an output element is accumulated in the same order whatever
// the number of token rows, so a row computed alone (draft / AR, m = 1) is bit-identical to
// the same row computed inside a verify block (m = D+1).
#pragma once
#include "common.cuh"

namespace lsk {

enum { PRO_RMS = 0, PRO_BF16 = 1 };
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_STORE = 2, EPI_SILU = 3, EPI_LMHEAD = 4 };

constexpr int kGemmWarps = 16;
constexpr int kGemmThreads = kGemmWarps * 32;
constexpr int kPrefetch = 8;  // super-blocks in flight per warp (16 x 16 B per lane)

struct GemmArgs {
  const uint4* W;       // packed weights
  int n_tiles;          // N / 16
  int nsb;              // K / 32
  int K;
  int ks_log2;          // log2 of the K-split across the warps of a CTA
  int M;                // valid token rows
  // ---- prologue
  const float* x_f32;   // PRO_RMS: residual-stream rows [M][x_ld] fp32
  int x_ld;
  const __nv_bfloat16* norm_w;
  float eps;
  const __nv_bfloat16* x_bf16;  // PRO_BF16: activations [M][xb_ld]
  int xb_ld;
  // ---- epilogue: RESID (+=) / STORE (=)
  float* out_f32;
  int out_ld;
  // ---- SILU: act[m][8*tile + r] = silu(gate) * up
  __nv_bfloat16* act;
  int act_ld;
  // ---- QKV: RoPE, q -> q_out, k/v -> paged cache
  __nv_bfloat16* q_out;
  int q_ld;
  __nv_bfloat16* kpool;  // layer base
  __nv_bfloat16* vpool;
  const int* page_table;
  const int* base_len;
  int pos_off;
  const float2* rope;    // [max_pos][64] (cos, sin)
  int q_rows;            // local q rows (heads * 128)
  int kv_rows;           // local kv rows
  int n_kv_heads;        // local
  // ---- LMHEAD
  float* logits;         // optional [M][logits_ld]
  int logits_ld;
  int n_valid_rows;      // local vocab rows (<= n_tiles*16)
  int vocab_off;         // global id of local row 0
  float* part_val;       // [grid][16]
  int* part_idx;
};

__host__ __device__ inline int gemm_x_stride_bytes(int K) {
  return ((2 * K + 127) / 128) * 128 + 64;   // == 64 (mod 128): conflict-free LDS.128
}

template <int NT>
__host__ __device__ inline size_t gemm_smem_bytes(int K, int epi, int ks_log2) {
  size_t s = (size_t)NT * 8 * gemm_x_stride_bytes(K);     // activations
  s += (size_t)kGemmWarps * NT * 128 * 4;                 // cross-warp reduction
  s += (size_t)(kGemmWarps + 1) * NT * 8 * 4;             // rms partials + rstd
  if (epi == EPI_LMHEAD) s += (size_t)NT * 8 * ((kGemmWarps >> ks_log2) * 16) * 4;
  return s;
}

template <int NT, int PRO, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_skinny_kernel(const GemmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int XS = gemm_x_stride_bytes(a.K);
  unsigned char* xs = smem;
  float* red = reinterpret_cast<float*>(smem + (size_t)NT * 8 * XS);
  float* stat = red + kGemmWarps * NT * 128;
  float* lg = stat + (kGemmWarps + 1) * NT * 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KS = 1 << a.ks_log2;
  const int TPC = kGemmWarps >> a.ks_log2;
  const int tile_local = warp >> a.ks_log2;
  const int ks = warp & (KS - 1);
  const int sb_per = a.nsb >> a.ks_log2;
  const int sb0 = ks * sb_per;
  const int n_groups = (a.n_tiles + TPC - 1) / TPC;

  uint4 abuf[kPrefetch][2];
  auto prefetch_head = [&](int grp) {
    const int tile = grp * TPC + tile_local;
    if (grp < n_groups && tile < a.n_tiles) {
      const uint4* wp = a.W + ((size_t)tile * a.nsb + sb0) * 64 + lane;
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i)
        if (i < sb_per) {
          abuf[i][0] = ldg_stream(wp + i * 64);
          abuf[i][1] = ldg_stream(wp + i * 64 + 32);
        }
    }
  };

  // Weights never depend on the previous kernel: get the stream going before the dependency.
  prefetch_head(blockIdx.x);
  pdl_launch_dependents();
  pdl_wait();

  // ------------------------------------------------------------------ prologue: x -> smem
  if (PRO == PRO_RMS) {
    // Two passes over the (L1/L2-resident) residual rows: sum of squares, then normalise ->
    // bf16 (rounding point of a bf16 HF model: modeling_llama.py:52-70).  Low register use so
    // the weight prefetch issued above stays in flight.
    const int nvec = a.K >> 2;  // float4 per row
    float ss[NT * 8];
#pragma unroll
    for (int m = 0; m < NT * 8; ++m) {
      ss[m] = 0.f;
      if (m < a.M) {
        const float4* xr = reinterpret_cast<const float4*>(a.x_f32 + (size_t)m * a.x_ld);
        for (int idx = tid; idx < nvec; idx += kGemmThreads) {
          const float4 v = xr[idx];
          ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < NT * 8; ++m) {
      const float tot = warp_sum(ss[m]);
      if (lane == 0) stat[warp * (NT * 8) + m] = tot;
    }
    __syncthreads();
    if (tid < NT * 8) {
      float tot = 0.f;
      for (int w = 0; w < kGemmWarps; ++w) tot += stat[w * (NT * 8) + tid];
      stat[kGemmWarps * NT * 8 + tid] = rsqrtf(tot / (float)a.K + a.eps);
    }
    __syncthreads();
    for (int m = 0; m < NT * 8; ++m) {
      if (m < a.M) {
        const float rstd = stat[kGemmWarps * NT * 8 + m];
        const float4* xr = reinterpret_cast<const float4*>(a.x_f32 + (size_t)m * a.x_ld);
        for (int idx = tid; idx < nvec; idx += kGemmThreads) {
          const float4 v = xr[idx];
          const uint2 wv = *reinterpret_cast<const uint2*>(a.norm_w + idx * 4);
          uint2 o;
          o.x = pack_bf16x2(bf16_lo(wv.x) * (v.x * rstd), bf16_hi(wv.x) * (v.y * rstd));
          o.y = pack_bf16x2(bf16_lo(wv.y) * (v.z * rstd), bf16_hi(wv.y) * (v.w * rstd));
          *reinterpret_cast<uint2*>(xs + (size_t)m * XS + idx * 8) = o;
        }
      } else {
        for (int idx = tid; idx < (a.K >> 3); idx += kGemmThreads)
          *reinterpret_cast<uint4*>(xs + (size_t)m * XS + idx * 16) = make_uint4(0, 0, 0, 0);
      }
    }
    __syncthreads();
  } else {
    const int nvec = a.K >> 3;  // uint4 (8 bf16) per row
    for (int m = 0; m < NT * 8; ++m) {
      for (int idx = tid; idx < nvec; idx += kGemmThreads) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (m < a.M) v = *reinterpret_cast<const uint4*>(a.x_bf16 + (size_t)m * a.xb_ld + idx * 8);
        *reinterpret_cast<uint4*>(xs + (size_t)m * XS + idx * 16) = v;
      }
    }
    __syncthreads();
  }

  const unsigned char* xlane = xs + (size_t)g * XS + t * 16;

  float best_v = -INFINITY;  // LMHEAD running arg-max (warp `m` owns token m)
  int best_i = 0x7fffffff;

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int tile = grp * TPC + tile_local;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

    if (tile < a.n_tiles) {
      const uint4* wp = a.W + ((size_t)tile * a.nsb + sb0) * 64 + lane;
      for (int sb = 0; sb < sb_per; sb += kPrefetch) {
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) {
          if (sb + i < sb_per) {
            const uint4 a0 = abuf[i][0], a1 = abuf[i][1];
            if (sb + i + kPrefetch < sb_per) {
              abuf[i][0] = ldg_stream(wp + (size_t)(sb + i + kPrefetch) * 64);
              abuf[i][1] = ldg_stream(wp + (size_t)(sb + i + kPrefetch) * 64 + 32);
            }
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const uint4 b = *reinterpret_cast<const uint4*>(
                  xlane + (size_t)n * 8 * XS + (size_t)(sb0 + sb + i) * 64);
              mma_bf16_16816(acc[n], a0.x, a0.y, a0.z, a0.w, b.x, b.y);
              mma_bf16_16816(acc[n], a1.x, a1.y, a1.z, a1.w, b.z, b.w);
            }
          }
        }
      }
    }
    // keep the HBM stream busy across the reduction / epilogue of this group
    prefetch_head(grp + gridDim.x);

#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float* r = red + ((warp * NT + n) * 16) * 8;
      *reinterpret_cast<float2*>(r + g * 8 + 2 * t) = make_float2(acc[n][0], acc[n][1]);
      *reinterpret_cast<float2*>(r + (g + 8) * 8 + 2 * t) = make_float2(acc[n][2], acc[n][3]);
    }
    __syncthreads();

    auto ksum = [&](int tl, int n, int row, int tok) {
      float s = 0.f;
      for (int k = 0; k < KS; ++k) s += red[((((tl << a.ks_log2) + k) * NT + n) * 16 + row) * 8 + tok];
      return s;
    };

    if (EPI == EPI_QKV || EPI == EPI_SILU) {
      const int items = TPC * NT * 64;
      for (int it = tid; it < items; it += kGemmThreads) {
        const int tok = it & 7, r = (it >> 3) & 7, n = (it >> 6) % NT, tl = (it >> 6) / NT;
        const int m = n * 8 + tok;
        const int tl_tile = grp * TPC + tl;
        if (m >= a.M || tl_tile >= a.n_tiles) continue;
        const float lo = ksum(tl, n, r, tok), hi = ksum(tl, n, r + 8, tok);
        if (EPI == EPI_SILU) {
          const float s = lo / (1.f + __expf(-lo));
          a.act[(size_t)m * a.act_ld + tl_tile * 8 + r] = __float2bfloat16_rn(s * hi);
        } else {
          const int pr = tl_tile * 16;                  // first packed row of the tile
          const int pos = *a.base_len + a.pos_off + m;
          if (pr < a.q_rows + a.kv_rows) {              // q or k: rotary pair (d, d + 64)
            const bool is_q = pr < a.q_rows;
            const int rel = is_q ? pr : pr - a.q_rows;
            const int head = rel >> 7, tt = (rel & 127) >> 4;
            const int d = tt * 8 + r;
            const float2 cs = a.rope[(size_t)pos * 64 + d];
            const float o_lo = lo * cs.x - hi * cs.y;
            const float o_hi = hi * cs.x + lo * cs.y;
            if (is_q) {
              __nv_bfloat16* q = a.q_out + (size_t)m * a.q_ld + head * 128;
              q[d] = __float2bfloat16_rn(o_lo);
              q[d + 64] = __float2bfloat16_rn(o_hi);
            } else {
              const int page = a.page_table[pos >> 6];
              __nv_bfloat16* kd = a.kpool +
                  ((size_t)(page * a.n_kv_heads + head) * kPageTokens + (pos & 63)) * 128;
              kd[d] = __float2bfloat16_rn(o_lo);
              kd[d + 64] = __float2bfloat16_rn(o_hi);
            }
          } else {                                       // v: natural order, no rotation
            const int rel = pr - a.q_rows - a.kv_rows;
            const int head = rel >> 7, d0 = rel & 127;
            const int page = a.page_table[pos >> 6];
            __nv_bfloat16* vd = a.vpool +
                ((size_t)(page * a.n_kv_heads + head) * kPageTokens + (pos & 63)) * 128;
            vd[d0 + r] = __float2bfloat16_rn(lo);
            vd[d0 + r + 8] = __float2bfloat16_rn(hi);
          }
        }
      }
    } else {
      const int items = TPC * NT * 128;
      for (int it = tid; it < items; it += kGemmThreads) {
        const int row = it & 15, tok = (it >> 4) & 7, n = (it >> 7) % NT, tl = (it >> 7) / NT;
        const int m = n * 8 + tok;
        const int tl_tile = grp * TPC + tl;
        if (tl_tile >= a.n_tiles) continue;
        const int orow = tl_tile * 16 + row;
        const float v = (m < a.M) ? ksum(tl, n, row, tok) : 0.f;
        if (EPI == EPI_RESID) {
          if (m < a.M) a.out_f32[(size_t)m * a.out_ld + orow] += v;
        } else if (EPI == EPI_STORE) {
          if (m < a.M) a.out_f32[(size_t)m * a.out_ld + orow] = v;
        } else {  // LMHEAD
          if (m < a.M && a.logits != nullptr && orow < a.n_valid_rows)
            a.logits[(size_t)m * a.logits_ld + orow] = v;
          lg[m * (TPC * 16) + tl * 16 + row] = v;
        }
      }
      if (EPI == EPI_LMHEAD) {
        __syncthreads();
        if (warp < NT * 8 && warp < a.M) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int i = lane; i < TPC * 16; i += 32) {
            const int orow = grp * TPC * 16 + i;
            if (orow < a.n_valid_rows) {
              const float v = lg[warp * (TPC * 16) + i];
              if (better(v, orow, bv, bi)) { bv = v; bi = orow; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
          }
          if (better(bv, bi, best_v, best_i)) { best_v = bv; best_i = bi; }
        }
      }
    }
    __syncthreads();
  }

  if (EPI == EPI_LMHEAD) {
    if (warp < NT * 8 && warp < a.M && lane == 0) {
      a.part_val[blockIdx.x * kMaxRows + warp] = best_v;
      a.part_idx[blockIdx.x * kMaxRows + warp] =
          (best_i == 0x7fffffff) ? 0x7fffffff : best_i + a.vocab_off;
    }
  }
}

}  // namespace lsk
