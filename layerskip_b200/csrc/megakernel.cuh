// megakernel.cuh — one persistent kernel per decode step / speculation round.
//
// Why: the stand-alone GEMM kernel streams at HBM speed in steady state, but every kernel
// boundary costs ~6-9 us during which HBM is idle (launch, first-byte latency of the ring,
// tile-quantisation tail, reduction + drain): measured slope 6.5 TB/s, intercept ~9 us per GEMM
// launch, 363 GEMM launches per 7B round.  Here the SAME device functions (gemm_producer /
// gemm_work / attn_team) run as stages of ONE cooperative kernel, one CTA per SM:
//   * the producer warp walks the weight streams of ALL GEMM stages back to back and never waits
//     for a grid barrier (weights do not depend on activations) — the TMA ring stays full while
//     consumers finish a stage, synchronise and run the next prologue;
//   * consumer + epilogue warps interpret the stage program, separated by software grid barriers
//     (one atomic counter, monotonically increasing target).
// Numerics are bit-identical to the multi-kernel path (same device code, same orders).
#pragma once
#include "attention.cuh"
#include "gemm_skinny.cuh"
#include "misc_kernels.cuh"

namespace lsk {

enum { ST_GEMM = 0, ST_ATTN = 1, ST_EMBED = 2, ST_FINALIZE = 3, ST_ACCEPT = 4, ST_AR_COMMIT = 5 };
constexpr int kAttnTeams = 3;    // 3 x 34.9 KiB of K/V staging fit the scratch region

struct MiscArgs {
  // EMBED: rows[r] = embed[ids[r]] for r < n_rows
  const __nv_bfloat16* embed;
  int hidden;
  const int* ids;
  float* rows;
  int row_ld;
  int n_rows;
  // FINALIZE / ACCEPT / AR_COMMIT
  const float* cand_val;
  const int* cand_idx;
  int n_cand;
  DevState* st;
  int slot;            // FINALIZE: st->tok[slot] = arg-max
  float* dst_row;      // FINALIZE: embedding of that token
  int d;               // ACCEPT
  const GenParams* gp;
  RoundResult* res;
};

struct StageDesc {
  int kind;
  int nt, pro, epi;    // ST_GEMM template selectors
  int barrier_before;  // grid barrier before this stage
  int pad_;
  GemmArgs g;
  AttnArgs a;
  MiscArgs m;
};

// Software grid barrier over the consumer + epilogue threads of all CTAs (the producer warps do
// not take part).  `target` = arrivals expected so far; the counter only grows.
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int target, int wtid,
                                          unsigned long long* phase_clk = nullptr) {
  bar_sync(BAR_WORK, kWorkThreads);            // this CTA's global writes are issued
  if (phase_clk != nullptr && wtid == 0) phase_clk[3] = clock64();      // debug: CTA done
  if (wtid == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  bar_sync(BAR_WORK, kWorkThreads);
}

// NB: `a` must be a LOCAL copy of the stage's arguments.  Reading them through a reference into
// global memory makes every `asm volatile(... "memory")` (mbarrier waits, named barriers) reload
// the fields inside the hot loops (measured: the first megakernel was 20 % slower than the
// multi-kernel path for exactly this reason).
template <int NT>
__device__ __forceinline__ void gemm_dispatch(int pro, int epi, const GemmArgs& a, const GemmCtx& c,
                                              uint32_t& q, int tid, int warp, int lane,
                                              unsigned long long* pc) {
  if (pro == PRO_RMS) {
    if (epi == EPI_QKV) gemm_work<NT, PRO_RMS, EPI_QKV>(a, c, q, tid, warp, lane, pc);
    else if (epi == EPI_SILU) gemm_work<NT, PRO_RMS, EPI_SILU>(a, c, q, tid, warp, lane, pc);
    else gemm_work<NT, PRO_RMS, EPI_LMHEAD>(a, c, q, tid, warp, lane, pc);
  } else {
    if (epi == EPI_RESID) gemm_work<NT, PRO_BF16, EPI_RESID>(a, c, q, tid, warp, lane, pc);
    else gemm_work<NT, PRO_BF16, EPI_STORE>(a, c, q, tid, warp, lane, pc);
  }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
step_megakernel(const StageDesc* __restrict__ prog, int n_stages_prog, unsigned int* grid_counter,
                int ring_stages, unsigned long long* timeline) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmCtx c = make_ctx(smem, ring_stages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) ctx_init_barriers(c);
  __syncthreads();

  uint32_t q = 0;
  if (warp == kProducerWarp) {
    // ============================================================ PRODUCER: all weight streams
    if (lane == 0) {
      for (int s = 0; s < n_stages_prog; ++s)
        if (prog[s].kind == ST_GEMM) {
          const GemmArgs ga = prog[s].g;      // local copy: see gemm_dispatch
          gemm_producer(ga, c, q);
        }
    }
    return;
  }

  const int wtid = (warp < kGemmWarps) ? tid : tid - 32;
  unsigned int n_bar = 0;
  // debug phase clocks (SM clock) for 4 sample CTAs: [cta_slot][stage][4] behind the timeline
  int cslot = -1;
  if (timeline != nullptr) {
    if (blockIdx.x == 0) cslot = 0;
    else if (blockIdx.x == 27) cslot = 1;
    else if (blockIdx.x == 100) cslot = 2;
    else if (blockIdx.x == gridDim.x - 1) cslot = 3;
  }
  unsigned long long* pclk = cslot >= 0 ? timeline + 4097 + (size_t)cslot * 4096 * 4 : nullptr;
  for (int s = 0; s < n_stages_prog; ++s) {
    const int kind = prog[s].kind, nt = prog[s].nt, pro = prog[s].pro, epi = prog[s].epi;
    unsigned long long* pc = (pclk != nullptr && s < 4095) ? pclk + (size_t)s * 4 : nullptr;
    if (prog[s].barrier_before) {
      ++n_bar;
      grid_sync(grid_counter, n_bar * gridDim.x, wtid, (pc != nullptr && s > 0) ? pc - 4 : nullptr);
    }
    if (pc != nullptr && wtid == 0) pc[0] = clock64();                 // debug: stage start
    if (timeline != nullptr && blockIdx.x == 0 && wtid == 0) {   // optional stage timeline (debug)
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      timeline[s] = t;
    }
    switch (kind) {
      case ST_GEMM: {
        const GemmArgs ga = prog[s].g;        // local copy: see gemm_dispatch
        if (nt == 1) gemm_dispatch<1>(pro, epi, ga, c, q, tid, warp, lane, pc);
        else gemm_dispatch<2>(pro, epi, ga, c, q, tid, warp, lane, pc);
        break;
      }
      case ST_ATTN: {
        // work items (kv head, split) dealt round-robin to (CTA, team); 12 of the 16 consumer
        // warps form 3 teams, each with its own K/V staging area in the scratch region
        if (warp < kAttnTeams * 4) {
          const AttnArgs aa = prog[s].a;
          const int team = warp >> 2;
          const int n_items = aa.n_kv_heads * aa.n_splits;
          unsigned char* tsm = c.scratch + (size_t)team * kAttnTeamSmem;
          for (int item = team * gridDim.x + blockIdx.x; item < n_items; item += kAttnTeams * gridDim.x)
            attn_team(aa, item / aa.n_splits, item % aa.n_splits, tsm, tid & 127, BAR_TEAM0 + team);
        }
        break;
      }
      case ST_EMBED: {
        const MiscArgs m = prog[s].m;
        for (int r = blockIdx.x; r < m.n_rows; r += gridDim.x)
          embed_row(m.embed, m.hidden, m.ids[r], m.rows + (size_t)r * m.row_ld, wtid, kWorkThreads);
        break;
      }
      case ST_FINALIZE: {
        // arg-max of the LM-head candidates -> next draft token -> its embedding row
        const MiscArgs m = prog[s].m;
        if (blockIdx.x == 0) {
          int* s_tok = reinterpret_cast<int*>(c.scratch);
          if (warp == 0) {
            const int tok = reduce_candidates(m.cand_val, m.cand_idx, m.n_cand, 0, lane);
            if (lane == 0) { *s_tok = tok; m.st->tok[m.slot] = tok; }
          }
          bar_sync(BAR_WORK, kWorkThreads);
          embed_row(m.embed, m.hidden, *s_tok, m.dst_row, wtid, kWorkThreads);
        }
        break;
      }
      case ST_ACCEPT: {
        const MiscArgs m = prog[s].m;
        if (blockIdx.x == 0) {
          int* s_ver = reinterpret_cast<int*>(c.scratch);
          if (warp < kGemmWarps)
            for (int row = warp; row <= m.d; row += kGemmWarps) {
              const int tok = reduce_candidates(m.cand_val, m.cand_idx, m.n_cand, row, lane);
              if (lane == 0) s_ver[row] = tok;
            }
          bar_sync(BAR_WORK, kWorkThreads);
          if (wtid == 0) accept_commit(s_ver, m.d, m.st, *m.gp, m.res, 0);
        }
        break;
      }
      case ST_AR_COMMIT: {
        const MiscArgs m = prog[s].m;
        if (blockIdx.x == 0 && warp == 0) {
          const int tok = reduce_candidates(m.cand_val, m.cand_idx, m.n_cand, 0, lane);
          if (lane == 0) ar_commit(tok, m.st, m.res, 0);
        }
        break;
      }
      default: break;
    }
  }
  if (timeline != nullptr && blockIdx.x == 0 && wtid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    timeline[n_stages_prog] = t;
  }
}

}  // namespace lsk
