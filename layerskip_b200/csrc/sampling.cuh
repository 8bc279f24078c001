// sampling.cuh — the reference's sampling branch on the device:
//   decode_next_token  (llama_model_utils.py:109-131): logits / T -> top-k (if > 0) -> top-p
//                      (always) -> softmax -> multinomial
//   rejection test     (self_speculation_generator.py:191-199) + max_fn residual (:27-29)
// Distributions are what must agree with the reference (its RNG stream cannot: torch's CPU
// generator vs a counter-based Philox here), so tests compare the filtered probability rows
// exactly and acceptance statistics within binomial error.
#pragma once
#include "common.cuh"
#include "misc_kernels.cuh"

namespace lsk {

constexpr int kSampleThreads = 1024;

// ---- Philox4x32-10 (Salmon et al.), counter-based: (seed, step, row, purpose) -> 4 x u32
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

enum { RNG_DRAFT = 1, RNG_VERIFY = 2, RNG_ACCEPT = 3, RNG_RESID = 4 };

__device__ __forceinline__ float rng_uniform(const GenParams& gp, int step, int row, int purpose) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)row, (uint32_t)purpose, 0x4c534bu),
                                make_uint2((uint32_t)gp.seed, (uint32_t)(gp.seed >> 32)));
  return u01(r.x);
}

// ---- block-wide helpers (blockDim.x == kSampleThreads), deterministic order
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (kSampleThreads >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (kSampleThreads >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// order-preserving key of a float (for radix selection)
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Inverse-CDF draw over weights w[0..V) in index order (torch.multinomial's distribution):
// returns the smallest j with sum_{i<=j} w_i > target * total.  Contiguous chunk per thread.
__device__ int block_sample_index(const float* __restrict__ w, int V, float u, float* red,
                                  int* s_pick) {
  const int chunk = (V + kSampleThreads - 1) / kSampleThreads;
  const int lo = threadIdx.x * chunk, hi = min(V, lo + chunk);
  float local = 0.f;
  for (int j = lo; j < hi; ++j) local += w[j];
  // exclusive scan of `local` over threads: warp scan + scan of warp totals
  const int wp = threadIdx.x >> 5, l = threadIdx.x & 31;
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (l >= o) incl += n;
  }
  __syncthreads();
  if (l == 31) red[wp] = incl;
  __syncthreads();
  if (wp == 0) {
    float t = red[l];
    float ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, ti, o);
      if (l >= o) ti += n;
    }
    red[64 + l] = ti - t;              // exclusive prefix of warp totals
    if (l == 31) red[32] = ti;         // grand total
  }
  if (threadIdx.x == 0) *s_pick = -1;
  __syncthreads();
  const float total = red[32];
  const float target = u * total;
  const float before = red[64 + wp] + (incl - local);
  if (local > 0.f && target >= before && target < before + local) {
    float run = before;
    int pick = hi - 1;
    for (int j = lo; j < hi; ++j) {
      run += w[j];
      if (target < run) { pick = j; break; }
    }
    *s_pick = pick;
  }
  __syncthreads();
  if (*s_pick < 0) {                    // rounding at the very end of the CDF: last positive weight
    if (threadIdx.x == 0) {
      int pick = 0;
      for (int j = V - 1; j >= 0; --j)
        if (w[j] > 0.f) { pick = j; break; }
      *s_pick = pick;
    }
    __syncthreads();
  }
  return *s_pick;
}

// One row per CTA: logits -> warped probabilities (written out) -> sampled token.
//   grid.x rows; row r reads logits + r*ld, writes probs + r*V and tok_out[r].
__global__ void __launch_bounds__(kSampleThreads)
warp_and_sample_kernel(const float* __restrict__ logits, int ld, int V,
                       const GenParams* __restrict__ gpp, const DevState* __restrict__ st,
                       float* __restrict__ probs, int* __restrict__ tok_out, int purpose,
                       int row_base) {
  __shared__ float red[96 + 32];
  __shared__ float hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ float s_g;
  __shared__ int s_pick;
  pdl_launch_dependents();
  pdl_wait();
  const GenParams gp = *gpp;
  const int row = blockIdx.x;
  const float* lg = logits + (size_t)row * ld;
  float* pr = probs + (size_t)row * V;
  const float inv_t = 1.0f / gp.temperature;

  // ---- top-k threshold (HF TopKLogitsWarper: drop scores < k-th largest)
  float kth = -INFINITY;
  if (gp.top_k > 0 && gp.top_k < V) {
    __shared__ int cnt[256];
    uint32_t prefix = 0;
    int need = gp.top_k;                 // how many keys >= threshold still to take
    for (int level = 3; level >= 0; --level) {
      for (int i = threadIdx.x; i < 256; i += kSampleThreads) cnt[i] = 0;
      __syncthreads();
      const uint32_t hi_mask = level == 3 ? 0u : (0xffffffffu << (8 * (level + 1)));
      for (int j = threadIdx.x; j < V; j += kSampleThreads) {
        const uint32_t k = fkey(lg[j] * inv_t);
        if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&cnt[(k >> (8 * level)) & 255], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int b = 255, acc = 0;
        for (; b > 0; --b) {
          if (acc + cnt[b] >= need) break;
          acc += cnt[b];
        }
        need -= acc;
        s_prefix = prefix | ((uint32_t)b << (8 * level));
      }
      __syncthreads();
      prefix = s_prefix;
      if (threadIdx.x == 0) { /* need is thread-0 private; broadcast via smem */ s_g = (float)need; }
      __syncthreads();
      need = (int)s_g;
    }
    const uint32_t kb = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
    kth = __uint_as_float(kb);
  }

  // ---- softmax numerator of the (top-k filtered) row
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += kSampleThreads) {
    const float v = lg[j] * inv_t;
    if (v >= kth) mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  float z = 0.f;
  for (int j = threadIdx.x; j < V; j += kSampleThreads) {
    const float v = lg[j] * inv_t;
    const float e = (v >= kth) ? __expf(v - mx) : 0.f;
    pr[j] = e;
    z += e;
  }
  z = block_sum(z, red);
  const float inv_z = 1.0f / z;

  // ---- nucleus: keep token i iff the mass of strictly larger tokens is < top_p
  uint32_t kstar = 0;                    // keep keys >= kstar
  if (gp.top_p >= 0.f && gp.top_p < 1.0f) {
    uint32_t prefix = 0;
    float G = 0.f;                       // mass above the bucket being refined
    for (int level = 3; level >= 0; --level) {
      for (int i = threadIdx.x; i < 256; i += kSampleThreads) hist[i] = 0.f;
      __syncthreads();
      const uint32_t hi_mask = level == 3 ? 0u : (0xffffffffu << (8 * (level + 1)));
      for (int j = threadIdx.x; j < V; j += kSampleThreads) {
        const float p = pr[j] * inv_z;
        const uint32_t k = __float_as_uint(p);       // p >= 0: bit order == value order
        if (p > 0.f && (k & hi_mask) == (prefix & hi_mask))
          atomicAdd(&hist[(k >> (8 * level)) & 255], p);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int b = 255;
        float g = G;
        for (; b > 0; --b) {
          if (g + hist[b] >= gp.top_p) break;
          g += hist[b];
        }
        s_g = g;
        s_prefix = prefix | ((uint32_t)b << (8 * level));
      }
      __syncthreads();
      prefix = s_prefix;
      G = s_g;
    }
    kstar = prefix;
  }
  {  // min_tokens_to_keep = 1: the arg-max (numerator exp(0) = 1) always survives
    const uint32_t kmax = __float_as_uint(inv_z);
    if (kstar > kmax) kstar = kmax;
  }
  float z2 = 0.f;
  for (int j = threadIdx.x; j < V; j += kSampleThreads) {
    const float e = pr[j];
    const bool keep = e > 0.f && __float_as_uint(e * inv_z) >= kstar;
    z2 += keep ? e : 0.f;
  }
  z2 = block_sum(z2, red);
  const float inv_z2 = 1.0f / z2;
  for (int j = threadIdx.x; j < V; j += kSampleThreads) {
    const float e = pr[j];
    const bool keep = e > 0.f && __float_as_uint(e * inv_z) >= kstar;
    pr[j] = keep ? e * inv_z2 : 0.f;
  }
  __syncthreads();

  // ---- multinomial draw
  const float u = rng_uniform(gp, st->step_count, row_base + row, purpose);
  const int tok = block_sample_index(pr, V, u, red, &s_pick);
  if (threadIdx.x == 0) tok_out[row] = tok;
}

// Rejection test + commit for the sampling path (self_speculation_generator.py:191-221).
__global__ void __launch_bounds__(kSampleThreads)
accept_sample_kernel(const float* __restrict__ p_draft, const float* __restrict__ p_verify, int V,
                     int d, DevState* __restrict__ st, const GenParams* __restrict__ gpp,
                     RoundResult* __restrict__ res, float* __restrict__ scratch, int seq,
                     int* __restrict__ hist) {
  __shared__ float red[96 + 32];
  __shared__ int s_pick;
  __shared__ int s_n, s_dact, s_reject;
  pdl_launch_dependents();
  pdl_wait();
  const GenParams gp = *gpp;
  if (threadIdx.x == 0) {
    int d_act = d;
    for (int i = 0; i < d; ++i)
      if (is_eos(gp, st->tok[1 + i])) { d_act = i + 1; break; }
    int n = 0, reject = -1;
    for (int i = 0; i < d_act; ++i) {
      const int t = st->tok[1 + i];
      const float pv = p_verify[(size_t)i * V + t], pd = p_draft[(size_t)i * V + t];
      const float u = rng_uniform(gp, st->step_count, i, RNG_ACCEPT);
      if (u < fminf(1.0f, pv / pd)) ++n;
      else { reject = i; break; }
    }
    s_n = n; s_dact = d_act; s_reject = reject;
  }
  __syncthreads();
  const int n = s_n, d_act = s_dact, reject = s_reject;
  int bonus;
  if (reject >= 0) {
    // resample from norm(max(p_v - p_d, 0))   (max_fn, :27-29)
    const float* pv = p_verify + (size_t)reject * V;
    const float* pd = p_draft + (size_t)reject * V;
    for (int j = threadIdx.x; j < V; j += kSampleThreads) scratch[j] = fmaxf(pv[j] - pd[j], 0.f);
    __syncthreads();
    const float u = rng_uniform(gp, st->step_count, reject, RNG_RESID);
    bonus = block_sample_index(scratch, V, u, red, &s_pick);
  } else {
    bonus = st->verified[n];             // the verifier's own draw at row n (= d_act)
  }
  if (threadIdx.x == 0) {
    res->n_drafted = d_act;
    res->n_matches = n;
    res->n_emitted = n + 1;
    for (int i = 0; i < d_act; ++i) res->draft_ids[i] = st->tok[1 + i];
    for (int i = 0; i <= d_act; ++i) res->verified_ids[i] = st->verified[i];
    res->verified_ids[n] = bonus;
    for (int i = 0; i < n; ++i) res->emitted_ids[i] = st->tok[1 + i];
    res->emitted_ids[n] = bonus;
    if (hist != nullptr)
      for (int i = 0; i <= n; ++i) hist[st->n_prompt + st->n_out + i] = res->emitted_ids[i];
    st->len += n + 1;
    st->n_out += n + 1;
    st->tok[0] = bonus;
    st->step_count += 1;
    res->kv_len = st->len;
    __threadfence_system();
    *reinterpret_cast<volatile int*>(&res->seq) = seq;
  }
}

// AR commit when the token was sampled into st->verified[0].
__global__ void ar_commit_sampled_kernel(DevState* __restrict__ st, RoundResult* __restrict__ res,
                                         int seq, int* __restrict__ hist) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    const int tok = st->verified[0];
    if (hist != nullptr) hist[st->n_prompt + st->n_out] = tok;
    st->tok[0] = tok;
    st->len += 1;
    st->n_out += 1;
    st->step_count += 1;
    res->n_drafted = 0; res->n_matches = 0; res->n_emitted = 1;
    res->emitted_ids[0] = tok; res->verified_ids[0] = tok;
    res->kv_len = st->len;
    __threadfence_system();
    *reinterpret_cast<volatile int*>(&res->seq) = seq;
  }
}

}  // namespace lsk
