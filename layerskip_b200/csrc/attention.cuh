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
#include <cooperative_groups.h>

#include "common.cuh"

namespace lsk {
namespace cg = cooperative_groups;

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
  int n_pages;                 // entries in page_table
};

constexpr int kAttnTeamSmem = 2 * kKeyGroup * kKvRowBytes + 128;   // K/V staging + flag
constexpr int BAR_TEAM0 = 8;                                       // named barriers 8..10

__device__ __forceinline__ void team_sync(int bar_id) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(kAttnThreads) : "memory");
}

// One (kv head, split) work item, executed by a TEAM of 4 warps (`tid` in 0..127) that owns
// `sm_raw` (kAttnTeamSmem bytes) and named barrier `bar_id`.  Used by the stand-alone kernel
// (one team per CTA) and by the step megakernel (three teams per CTA).
// Phase 1 of a (kv head, split) work item: online-softmax partial (m, l, O) of this split's key
// groups for every query row, 4 warps merged in fixed order, written to po[row][128] /
// pml[row][2] (global scratch OR shared memory — the cluster kernel keeps it on chip).
__device__ __forceinline__ void attn_partial(const AttnArgs& a, int kvh, int split,
                                             unsigned char* sm_raw, int tid, int bar_id,
                                             float* __restrict__ po_base, float* __restrict__ pml_base) {
  unsigned char* ks = sm_raw;
  unsigned char* vs = sm_raw + kKeyGroup * kKvRowBytes;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // the first page-table entry does not depend on the sequence length: fetch it alongside it
  const int page_first = a.page_table[split < a.n_pages ? split : 0];
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
      team_sync(bar_id);   // previous iteration's readers are done with ks / vs
      {
        const int page = (kg == split) ? page_first : a.page_table[kg];   // kKeyGroup == kPageTokens
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
      team_sync(bar_id);

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
    team_sync(bar_id);   // everyone is done with ks / vs: reuse as merge buffer
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
    team_sync(bar_id);
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
      const size_t prow = (size_t)(rb * 16 + row);
      float* po = po_base + prow * kHeadDim + dseg;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(po + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      if ((tid & 7) == 0) {
        pml_base[prow * 2] = mm;
        pml_base[prow * 2 + 1] = ll;
      }
    }
    team_sync(bar_id);   // merge buffer free before the next row block refills ks / vs
  }

}

// One (kv head, split) work item with the GLOBAL-scratch merge: partial -> global, then the last
// team to finish for this kv head (atomic ticket) merges the splits in fixed order.  Used by the
// step megakernel (three teams per CTA) and as the non-cluster fallback kernel.
__device__ __forceinline__ void attn_team(const AttnArgs& a, int kvh, int split,
                                          unsigned char* sm_raw, int tid, int bar_id) {
  volatile int* s_last = reinterpret_cast<volatile int*>(sm_raw + 2 * kKeyGroup * kKvRowBytes);
  const size_t blk = (size_t)(kvh * a.n_splits + split) * a.rows_pad;
  attn_partial(a, kvh, split, sm_raw, tid, bar_id, a.part_o + blk * kHeadDim, a.part_ml + blk * 2);
  const int R = a.group * a.M;
  // ---- cross-split merge by the last CTA of this kv head
  __threadfence();
  team_sync(bar_id);
  if (tid == 0) {
    const int tk = atomicAdd(&a.tickets[kvh], 1);
    *s_last = (tk == a.n_splits - 1);
  }
  team_sync(bar_id);
  if (!*s_last) return;
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

// Fixed-order merge of the splits' partials for one (row, 16-dim segment); `ml(s)` / `o(s)` return
// split s's (m, l) pair and O segment — from global scratch or from a peer CTA's shared memory.
template <typename FML, typename FO>
__device__ __forceinline__ void merge_splits_write(const AttnArgs& a, int kvh, int row, int dseg,
                                                   FML ml, FO o) {
  // all (m, l) pairs first, then the O segments in batches of independent loads: 1 + 4 remote
  // latencies instead of a dependent chain of 2 x n_splits (same arithmetic, same order)
  constexpr int kMaxSplits = 8;
  float ms[kMaxSplits], ls[kMaxSplits];
#pragma unroll
  for (int s = 0; s < kMaxSplits; ++s) {
    ms[s] = -INFINITY; ls[s] = 0.f;
    if (s < a.n_splits) { const float* p = ml(s); ms[s] = p[0]; ls[s] = p[1]; }
  }
  float mm = -INFINITY;
#pragma unroll
  for (int s = 0; s < kMaxSplits; ++s) mm = fmaxf(mm, ms[s]);
  float f[kMaxSplits];
  float ll = 0.f;
#pragma unroll
  for (int s = 0; s < kMaxSplits; ++s) {
    f[s] = (ms[s] == -INFINITY) ? 0.f : __expf(ms[s] - mm);
    if (s < a.n_splits) ll += ls[s] * f[s];
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 v[kMaxSplits];
#pragma unroll
    for (int s = 0; s < kMaxSplits; ++s)
      v[s] = (s < a.n_splits) ? reinterpret_cast<const float4*>(o(s))[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < kMaxSplits; ++s) {
      if (s < a.n_splits) {
        acc[4 * i] += v[s].x * f[s]; acc[4 * i + 1] += v[s].y * f[s];
        acc[4 * i + 2] += v[s].z * f[s]; acc[4 * i + 3] += v[s].w * f[s];
      }
    }
  }
  const float inv = 1.f / ll;
  const int tok = row / a.group, hq = kvh * a.group + row % a.group;
  __nv_bfloat16* op = a.out + (size_t)tok * a.out_ld + hq * kHeadDim + dseg;
#pragma unroll
  for (int i = 0; i < 16; i += 2)
    *reinterpret_cast<uint32_t*>(op + i) = pack_bf16x2(acc[i] * inv, acc[i + 1] * inv);
}

// Cluster kernel: the n_splits CTAs of one kv head form a thread-block cluster; partials stay in
// shared memory and are merged through DISTRIBUTED shared memory after one cluster barrier —
// no global scratch, no __threadfence, no atomic ticket, no re-read (the global-merge chain cost
// 3-4 dependent memory round trips of the ~7 on this kernel's critical path).
__global__ void __launch_bounds__(kAttnThreads)
attn_cluster_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  cg::cluster_group cluster = cg::this_cluster();
  float* po = reinterpret_cast<float*>(dsm + kAttnTeamSmem);
  float* pml = po + (size_t)a.rows_pad * kHeadDim;
  const int kvh = blockIdx.x, split = blockIdx.y;          // cluster = all splits of one kv head
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  attn_partial(a, kvh, split, dsm, tid, BAR_TEAM0, po, pml);
  cluster.sync();
  const int R = a.group * a.M;
  for (int item = split * kAttnThreads + tid; item < R * 8; item += a.n_splits * kAttnThreads) {
    const int row = item >> 3, dseg = (item & 7) * 16;
    merge_splits_write(
        a, kvh, row, dseg,
        [&](int s) { return (const float*)cluster.map_shared_rank(pml, s) + row * 2; },
        [&](int s) { return (const float*)cluster.map_shared_rank(po, s) + (size_t)row * kHeadDim + dseg; });
  }
  cluster.sync();      // nobody leaves while a peer may still read its partials
}

// Push variant of the cluster kernel (opt-in, LSK_ATTN_PUSH=1).  Same partials, same merge
// arithmetic in the same order (merge_splits_write), but the exchange is turned around: instead of
// every merging thread PULLING 1 + 4 dependent remote reads out of its peers' shared memory and
// the cluster synchronising twice (once before the reads, once so nobody exits while being read),
// every CTA PUSHES the slices of its partial to the CTA that will merge them (fire-and-forget
// distributed-shared-memory stores), ONE cluster barrier publishes them, and the merge reads local
// shared memory only; nobody touches a peer afterwards, so a CTA may exit right after its merge.
// Items (row, 16-dim segment) are dealt round-robin over the cluster: item -> CTA item % n_splits.
constexpr int kInboxStride = 20;    // floats per (split, item): 16 of O, m, l, 2 pad (16 B aligned)

__host__ __device__ inline size_t attn_push_smem_bytes(int rows_pad) {
  return (size_t)kAttnTeamSmem + (size_t)rows_pad * (kHeadDim + 2) * 4 + ((size_t)rows_pad * 8 + 8) * kInboxStride * 4;   // + 8: cap is rounded up per CTA
}

__global__ void __launch_bounds__(kAttnThreads)
attn_cluster_push_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  cg::cluster_group cluster = cg::this_cluster();
  float* po = reinterpret_cast<float*>(dsm + kAttnTeamSmem);
  float* pml = po + (size_t)a.rows_pad * kHeadDim;
  float* inbox = pml + (size_t)a.rows_pad * 2;             // [n_splits][cap][kInboxStride]
  const int kvh = blockIdx.x, split = blockIdx.y;
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  attn_partial(a, kvh, split, dsm, tid, BAR_TEAM0, po, pml);
  const int R = a.group * a.M;
  const int n_items = R * 8;
  const int cap = (n_items + a.n_splits - 1) / a.n_splits;   // items per merging CTA (upper bound)
  // push: item -> (destination CTA, slot in its inbox), my partial goes to row `split` of that inbox
  for (int item = tid; item < n_items; item += kAttnThreads) {
    const int row = item >> 3, dseg = (item & 7) * 16;
    const int dest = item % a.n_splits, li = item / a.n_splits;
    float* dst = cluster.map_shared_rank(inbox, dest) + ((size_t)split * cap + li) * kInboxStride;
    const float4* src = reinterpret_cast<const float4*>(po + (size_t)row * kHeadDim + dseg);
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = src[i];
    *reinterpret_cast<float2*>(dst + 16) = make_float2(pml[row * 2], pml[row * 2 + 1]);
  }
  cluster.sync();      // release / acquire: every push has landed
  for (int li = tid; li < cap; li += kAttnThreads) {
    const int item = li * a.n_splits + split;
    if (item >= n_items) break;
    const int row = item >> 3, dseg = (item & 7) * 16;
    merge_splits_write(
        a, kvh, row, dseg,
        [&](int s) { return (const float*)(inbox + ((size_t)s * cap + li) * kInboxStride + 16); },
        [&](int s) { return (const float*)(inbox + ((size_t)s * cap + li) * kInboxStride); });
  }
}

__global__ void __launch_bounds__(kAttnThreads)
attn_splitkv_kernel(const AttnArgs a) {
  __shared__ __align__(128) unsigned char sm_raw[kAttnTeamSmem];
  pdl_launch_dependents();
  pdl_wait();
  attn_team(a, blockIdx.x, blockIdx.y, sm_raw, threadIdx.x, BAR_TEAM0);
}

}  // namespace lsk
