// attention.cuh — paged split-KV attention for decode / verify blocks (<= 16 query tokens).
//
// One CTA = (kv head, split).  The query rows of a CTA are all tokens x all q-heads that share
// the kv head (GQA), processed 16 rows at a time as the M side of mma.sync m16n8k16.  A split
// owns the 64-key groups  s, s + n_splits, ...  (by ABSOLUTE key index, so the partition seen
// by a query at position p does not depend on how many rows are in flight — this keeps the
// result batch-invariant).  Rounding points mirror a bf16 HF model: q/k/v bf16, scores and
// softmax fp32, probabilities rounded to bf16 for P.V, fp32 accumulate
// (transformers modeling_llama.py:187-221).
//
// Partials (m, l, O) are merged in fixed order: 4 warps inside the CTA, then the splits by the
// last CTA to finish for that kv head (atomic ticket).  Deterministic.
#pragma once
#include "common.cuh"

namespace lsk {

constexpr int kAttnThreads = 128;
constexpr int kKeyGroup = 64;                 // keys per CTA iteration (one KV page)
constexpr int kKvRowBytes = kHeadDim * 2 + 16;  // padded smem row: conflict-free LDS / ldmatrix

struct AttnArgs {
  const __nv_bfloat16* q;      // [M][q_ld] post-RoPE
  int q_ld;
  __nv_bfloat16* out;          // [M][out_ld]
  int out_ld;
  const __nv_bfloat16* kpool;  // layer base
  const __nv_bfloat16* vpool;
  const int* page_table;
  const int* base_len;
  int pos_off;
  int M;
  int group;                   // q heads per kv head
  int n_kv_heads;              // local
  int n_splits;
  float scale;                 // head_dim^-0.5
  float* part_o;               // [kv][split][rows_pad][128]
  float* part_ml;              // [kv][split][rows_pad][2]
  int rows_pad;                // group * 16 rounded up to 16
  int* tickets;                // [kv]
};

__global__ void __launch_bounds__(kAttnThreads)
attn_splitkv_kernel(const AttnArgs a) {
  __shared__ __align__(128) unsigned char sm_raw[2 * kKeyGroup * kKvRowBytes];
  __shared__ int s_last;
  unsigned char* ks = sm_raw;
  unsigned char* vs = sm_raw + kKeyGroup * kKvRowBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kvh = blockIdx.x, split = blockIdx.y;

  pdl_launch_dependents();
  pdl_wait();

  const int base = *a.base_len + a.pos_off;          // position of token row 0
  const int n_keys = base + a.M;                      // keys visible to the last row
  const int n_kgroups = (n_keys + kKeyGroup - 1) / kKeyGroup;
  const int R = a.group * a.M;                        // real query rows (token-major)
  const int n_rb = (R + 15) / 16;

  for (int rb = 0; rb < n_rb; ++rb) {
    // ---- Q fragments for rows rb*16 + {g, g+8}
    uint32_t qf[8][4];
    {
      const int r0 = rb * 16 + g, r1 = r0 + 8;
      const __nv_bfloat16* q0 = nullptr;
      const __nv_bfloat16* q1 = nullptr;
      if (r0 < R) q0 = a.q + (size_t)(r0 / a.group) * a.q_ld + (kvh * a.group + r0 % a.group) * kHeadDim;
      if (r1 < R) q1 = a.q + (size_t)(r1 / a.group) * a.q_ld + (kvh * a.group + r1 % a.group) * kHeadDim;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        qf[k][0] = q0 ? *reinterpret_cast<const uint32_t*>(q0 + k * 16 + 2 * t) : 0u;
        qf[k][1] = q1 ? *reinterpret_cast<const uint32_t*>(q1 + k * 16 + 2 * t) : 0u;
        qf[k][2] = q0 ? *reinterpret_cast<const uint32_t*>(q0 + k * 16 + 8 + 2 * t) : 0u;
        qf[k][3] = q1 ? *reinterpret_cast<const uint32_t*>(q1 + k * 16 + 8 + 2 * t) : 0u;
      }
    }
    const int row0 = rb * 16 + g, row1 = row0 + 8;
    const int lim0 = (row0 < R) ? base + row0 / a.group : -1;   // last visible key index
    const int lim1 = (row1 < R) ? base + row1 / a.group : -1;

    float o[16][4];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kg = split; kg < n_kgroups; kg += a.n_splits) {
      __syncthreads();   // previous iteration's readers are done with ks / vs
      {
        const int page = a.page_table[kg];   // kKeyGroup == kPageTokens
        const __nv_bfloat16* kp = a.kpool + (size_t)(page * a.n_kv_heads + kvh) * kPageTokens * kHeadDim;
        const __nv_bfloat16* vp = a.vpool + (size_t)(page * a.n_kv_heads + kvh) * kPageTokens * kHeadDim;
        for (int c = tid; c < kKeyGroup * 16; c += kAttnThreads) {
          const int key = c >> 4, ch = c & 15;
          uint4 kv4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
          if (kg * kKeyGroup + key < n_keys) {
            kv4 = *reinterpret_cast<const uint4*>(kp + key * kHeadDim + ch * 8);
            vv4 = *reinterpret_cast<const uint4*>(vp + key * kHeadDim + ch * 8);
          }
          *reinterpret_cast<uint4*>(ks + key * kKvRowBytes + ch * 16) = kv4;
          *reinterpret_cast<uint4*>(vs + key * kKvRowBytes + ch * 16) = vv4;
        }
      }
      __syncthreads();

      // ---- S = Q K^T for this warp's 16 keys (two n8 tiles)
      float s[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const unsigned char* kr = ks + (warp * 16 + n * 8 + g) * kKvRowBytes + (k * 16 + 2 * t) * 2;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + 16);
          mma_bf16_16816(s[n], qf[k][0], qf[k][1], qf[k][2], qf[k][3], b0, b1);
        }
      }
      // ---- scale + causal mask + online softmax (rows g and g+8)
      const int key0 = kg * kKeyGroup + warp * 16 + 2 * t;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int kidx = key0 + n * 8 + c;
          s[n][c] = (kidx <= lim0) ? s[n][c] * a.scale : -INFINITY;
          s[n][2 + c] = (kidx <= lim1) ? s[n][2 + c] * a.scale : -INFINITY;
          mx0 = fmaxf(mx0, s[n][c]);
          mx1 = fmaxf(mx1, s[n][2 + c]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float nm0 = fmaxf(m0, mx0), nm1 = fmaxf(m1, mx1);
      const float sc0 = (nm0 == -INFINITY) ? 1.f : __expf(m0 - nm0);
      const float sc1 = (nm1 == -INFINITY) ? 1.f : __expf(m1 - nm1);
      m0 = nm0;
      m1 = nm1;
      float p[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          p[n][c] = (nm0 == -INFINITY) ? 0.f : __expf(s[n][c] - nm0);
          p[n][2 + c] = (nm1 == -INFINITY) ? 0.f : __expf(s[n][2 + c] - nm1);
        }
      }
      l0 = l0 * sc0 + p[0][0] + p[0][1] + p[1][0] + p[1][1];
      l1 = l1 * sc1 + p[0][2] + p[0][3] + p[1][2] + p[1][3];
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        o[d][0] *= sc0; o[d][1] *= sc0; o[d][2] *= sc1; o[d][3] *= sc1;
      }
      // ---- O += P V  (P as bf16 A fragments, V through ldmatrix.trans)
      const uint32_t pa0 = pack_bf16x2(p[0][0], p[0][1]);
      const uint32_t pa1 = pack_bf16x2(p[0][2], p[0][3]);
      const uint32_t pa2 = pack_bf16x2(p[1][0], p[1][1]);
      const uint32_t pa3 = pack_bf16x2(p[1][2], p[1][3]);
      {
        const int mat = lane >> 3;
        const unsigned char* vrow = vs + (warp * 16 + (mat & 1) * 8 + (lane & 7)) * kKvRowBytes +
                                    (mat >> 1) * 16;
#pragma unroll
        for (int d = 0; d < 16; d += 2) {
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vrow + d * 16);
          mma_bf16_16816(o[d], pa0, pa1, pa2, pa3, vb[0], vb[1]);
          mma_bf16_16816(o[d + 1], pa0, pa1, pa2, pa3, vb[2], vb[3]);
        }
      }
    }

    // ---- merge the 4 warps (fixed order) through shared memory
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    __syncthreads();   // everyone is done with ks / vs: reuse as merge buffer
    float* mo = reinterpret_cast<float*>(sm_raw);            // [4 warps][16 rows][128]
    float* mml = mo + 4 * 16 * kHeadDim;                      // [4][16][2]
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      *reinterpret_cast<float2*>(mo + ((warp * 16 + g) * kHeadDim) + d * 8 + 2 * t) =
          make_float2(o[d][0], o[d][1]);
      *reinterpret_cast<float2*>(mo + ((warp * 16 + g + 8) * kHeadDim) + d * 8 + 2 * t) =
          make_float2(o[d][2], o[d][3]);
    }
    if (t == 0) {
      mml[(warp * 16 + g) * 2] = m0;
      mml[(warp * 16 + g) * 2 + 1] = l0;
      mml[(warp * 16 + g + 8) * 2] = m1;
      mml[(warp * 16 + g + 8) * 2 + 1] = l1;
    }
    __syncthreads();
    {
      // thread -> (row = tid / 8, 16 dims = (tid % 8) * 16 ..)
      const int row = tid >> 3, dseg = (tid & 7) * 16;
      float mm = -INFINITY;
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, mml[(w * 16 + row) * 2]);
      float ll = 0.f;
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      for (int w = 0; w < 4; ++w) {
        const float mw = mml[(w * 16 + row) * 2];
        const float f = (mw == -INFINITY) ? 0.f : __expf(mw - mm);
        ll += mml[(w * 16 + row) * 2 + 1] * f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += mo[(w * 16 + row) * kHeadDim + dseg + i] * f;
      }
      const size_t prow = ((size_t)(kvh * a.n_splits + split) * a.rows_pad + rb * 16 + row);
      float* po = a.part_o + prow * kHeadDim + dseg;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(po + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      if ((tid & 7) == 0) {
        a.part_ml[prow * 2] = mm;
        a.part_ml[prow * 2 + 1] = ll;
      }
    }
    __syncthreads();   // merge buffer free before the next row block refills ks / vs
  }

  // ---- cross-split merge by the last CTA of this kv head
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int tk = atomicAdd(&a.tickets[kvh], 1);
    s_last = (tk == a.n_splits - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int item = tid; item < R * 8; item += kAttnThreads) {
    const int row = item >> 3, dseg = (item & 7) * 16;
    float mm = -INFINITY;
    for (int s = 0; s < a.n_splits; ++s)
      mm = fmaxf(mm, __ldcg(a.part_ml + ((size_t)(kvh * a.n_splits + s) * a.rows_pad + row) * 2));
    float ll = 0.f;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int s = 0; s < a.n_splits; ++s) {
      const size_t prow = (size_t)(kvh * a.n_splits + s) * a.rows_pad + row;
      const float ms = __ldcg(a.part_ml + prow * 2);
      const float f = (ms == -INFINITY) ? 0.f : __expf(ms - mm);
      ll += __ldcg(a.part_ml + prow * 2 + 1) * f;
      const float4* po = reinterpret_cast<const float4*>(a.part_o + prow * kHeadDim + dseg);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(po + i);
        acc[4 * i] += v.x * f; acc[4 * i + 1] += v.y * f;
        acc[4 * i + 2] += v.z * f; acc[4 * i + 3] += v.w * f;
      }
    }
    const float inv = 1.f / ll;
    const int tok = row / a.group, hq = kvh * a.group + row % a.group;
    __nv_bfloat16* op = a.out + (size_t)tok * a.out_ld + hq * kHeadDim + dseg;
#pragma unroll
    for (int i = 0; i < 16; i += 2)
      *reinterpret_cast<uint32_t*>(op + i) = pack_bf16x2(acc[i] * inv, acc[i + 1] * inv);
  }
  if (tid == 0) a.tickets[kvh] = 0;
}

}  // namespace lsk
