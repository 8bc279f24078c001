/*
 * lsk.h — C ABI of the B200-native LayerSkip self-speculative decoding engine (liblsk.so).
 *
 * The reference (facebookresearch/LayerSkip) has NO FFI: its hot path is Python calling
 * HuggingFace modules.  This ABI is what a binding for that path would bind; each entry point
 * names the reference code it replaces (paths relative to the reference root).  Plain C types
 * only — device pointers travel as `const void*`, streams are owned by the engine.  Every call
 * returns 0 on success or a negative lsk_status; the text is available from lsk_last_error().
 * Not re-entrant: one host thread drives one engine (the reference is single-threaded too,
 * self_speculation/generator_base.py:97-130).
 */
#ifndef LSK_H_
#define LSK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSK_ABI_VERSION 2
#define LSK_MAX_SPEC 15      /* D_max: verify handles up to 16 rows (D+1)                    */
#define LSK_MAX_EOS 8

typedef enum {
  LSK_OK = 0,
  LSK_ERR_INVALID = -1,      /* bad argument / unsupported shape                             */
  LSK_ERR_CUDA = -2,         /* CUDA runtime error (message has the CUDA string)             */
  LSK_ERR_STATE = -3,        /* call order violated (e.g. round before prefill)              */
  LSK_ERR_NCCL = -4,
  LSK_ERR_NOMEM = -5,
  LSK_ERR_CTX = -6           /* sequence would exceed max_ctx                                */
} lsk_status;

/* flags */
#define LSK_FLAG_KEEP_LOGITS 1u  /* also store fp32 logits (needed for sampling / debug reads) */
#define LSK_FLAG_NO_PDL 2u       /* disable programmatic dependent launch                    */
#define LSK_FLAG_NO_GRAPH 4u     /* launch kernels eagerly instead of replaying CUDA graphs  */
#define LSK_FLAG_NO_PREFILL_TC 8u /* keep only the decode-layout weights: the prompt pass runs 16 rows at a time on the decode kernels instead of 128 rows at a time on tcgen05 (saves the second, canonical-layout weight copy) */
#define LSK_FLAG_TP_NCCL 16u     /* tp_size > 1: use NCCL all-reduce instead of the one-shot kernels over peer-mapped HBM */

/* Llama architecture + engine sizing.  Replaces what the reference reads off the HF model
 * object (`model.config`, generate.py:54-67). */
typedef struct {
  int32_t vocab, hidden, inter, n_layers, n_heads, n_kv_heads, head_dim;
  float rms_eps, rope_theta;
  int32_t max_ctx;           /* prompt + generated tokens the KV pool must hold              */
  int32_t tp_rank, tp_size;  /* tensor-parallel shard of this process (1 process per GPU)    */
  int32_t attn_splits;       /* split-KV factor (0 = default)                                */
  uint32_t flags;
  /* RoPE frequency scaling (HF `rope_scaling` / `rope_parameters`, transformers
   * modeling_rope_utils.py: _compute_linear_scaling_rope_parameters, _compute_llama3_parameters):
   * 0 default, 1 linear (inv_freq / factor), 2 llama3 (Llama-3.1 / 3.2 checkpoints such as
   * facebook/layerskip-llama3.2-1B, the reference's own test model, tests/tests_constants.py:9). */
  int32_t rope_scaling;
  float rope_factor, rope_low_freq_factor, rope_high_freq_factor;
  int32_t rope_original_max_pos;
} lsk_config;
#define LSK_ROPE_DEFAULT 0
#define LSK_ROPE_LINEAR 1
#define LSK_ROPE_LLAMA3 2

/* Which HF tensor a weight descriptor carries (names as in
 * transformers LlamaForCausalLM.state_dict(); call sites llama_model_utils.py:182,193,204-205). */
typedef enum {
  LSK_W_EMBED = 0,           /* model.embed_tokens.weight            [vocab, hidden]          */
  LSK_W_FINAL_NORM = 1,      /* model.norm.weight                    [hidden]                 */
  LSK_W_LM_HEAD = 2,         /* lm_head.weight                       [vocab, hidden]          */
  LSK_W_LN1 = 3,             /* layers.i.input_layernorm.weight      [hidden]                 */
  LSK_W_Q = 4, LSK_W_K = 5, LSK_W_V = 6, LSK_W_O = 7,   /* self_attn.{q,k,v,o}_proj.weight     */
  LSK_W_LN2 = 8,             /* layers.i.post_attention_layernorm.weight                      */
  LSK_W_GATE = 9, LSK_W_UP = 10, LSK_W_DOWN = 11         /* mlp.{gate,up,down}_proj.weight      */
} lsk_weight_role;

/* One full (unsharded) bf16 tensor in DEVICE memory, row-major [rows, cols] as HF stores it.
 * The engine slices its tensor-parallel shard, repacks it into its own HBM layout and does not
 * keep the pointer: the caller may free the tensor when lsk_load_weights returns. */
typedef struct {
  int32_t role;              /* lsk_weight_role                                              */
  int32_t layer;             /* decoder layer index, ignored for embed / final norm / head   */
  const void* data;          /* device pointer, bf16                                         */
  int64_t rows, cols;
} lsk_weight_desc;

/* Per-generation settings: `GenerationConfig` (self_speculation/generator_base.py:33-49) plus the
 * eos list built at generator_base.py:106. */
typedef struct {
  int32_t exit_layer;        /* E; <= 0 means "all layers" for lsk_ar_step                   */
  int32_t max_steps;
  int32_t n_eos;
  int32_t eos_ids[LSK_MAX_EOS];
  int32_t sample;            /* 0 greedy (arg-max), 1 sampling                               */
  float temperature;
  int32_t top_k;
  float top_p;
  int32_t no_repeat_ngram_size; /* > 0: NoRepeatNGramLogitsProcessor on the device (generator_base.py:77-85;
                              * transformers logits_process.py _calc_banned_ngram_tokens) over
                              * prompt + output + drafts; 0 = off                              */
  uint64_t seed;             /* counter-based RNG seed for the sampling path                 */
} lsk_generation;

/* What one speculation round produced — everything
 * SelfSpeculativeGenerationStrategy.single_step_speculation returns or streams
 * (self_speculation_generator.py:102-229): the draft ids (for SpeculativeTextStreamer, :158-161),
 * number_of_matches (:185-199), the tokens appended to output_ids (:203-205). */
typedef struct {
  int32_t n_drafted;                     /* D_actual (EOS can end the draft loop early)       */
  int32_t n_matches;
  int32_t n_emitted;                     /* n_matches + 1                                     */
  int32_t kv_len;                        /* committed context after the round                 */
  int32_t draft_ids[LSK_MAX_SPEC + 1];
  int32_t emitted_ids[LSK_MAX_SPEC + 1]; /* draft[:n] + [verified[n]]                         */
  int32_t verified_ids[LSK_MAX_SPEC + 1];
} lsk_round_out;

typedef struct lsk_engine lsk_engine;

int lsk_abi_version(void);
const char* lsk_last_error(void);

/* Engine lifetime.  Allocates packed-weight storage, the paged KV pool, scratch and streams on
 * the CURRENT CUDA device.  Replaces model placement in generate.py:54-67. */
int lsk_create(const lsk_config* cfg, lsk_engine** out);
void lsk_destroy(lsk_engine* e);

/* Tensor-parallel wiring (configs with tp_size > 1): rank 0 calls lsk_comm_unique_id, the host
 * side broadcasts the 128 bytes, every rank calls lsk_comm_init.  The reference has no
 * equivalent (generate.py:50-52 exits on non-zero ranks). */
int lsk_comm_unique_id(uint8_t id_out[128]);
int lsk_comm_init(lsk_engine* e, const uint8_t id[128]);

/* Weight ingest: HF tensors -> packed / sharded HBM layout.  Synchronous. */
int lsk_load_weights(lsk_engine* e, const lsk_weight_desc* descs, int32_t n);
/* 1 when every tensor the architecture needs has been loaded. */
int lsk_weights_complete(const lsk_engine* e);

/* Start a generation: reset lengths, store E / eos / sampling.  Replaces the state reset at
 * self_speculation_generator.py:41-50. */
int lsk_begin(lsk_engine* e, const lsk_generation* gen);

/* Prompt ingestion from HOST memory (ids[n], n >= 1).  Runs ids[0..n-2] through all layers (the
 * work the reference does inside its first forward_early + forward_remainder,
 * llama_model_utils.py:251-261, 363-383) so that afterwards every round has the steady-state
 * shape: one pending input token (ids[n-1]) and kv_len == n-1 in every layer. */
int lsk_prefill(lsk_engine* e, const int32_t* ids, int32_t n);

/* One draft / verify / accept / commit round with d_req speculations
 * (self_speculation_generator.py:102-229; the caller applies the max_steps clamp of :63-66).
 * d_req == 0 is the reference's tail round.  Blocks until the round's result is on the host. */
int lsk_round(lsk_engine* e, int32_t d_req, lsk_round_out* out);

/* One autoregressive step on the same engine (autoregressive_generator.py:44-67): all layers, or
 * layers < E when the generation's exit_layer > 0.  Returns the chosen token; the caller decides
 * about EOS exactly as the reference does (:66-67). */
int lsk_ar_step(lsk_engine* e, int32_t* token_out);

/* Queries / debugging (parity tests). */
int lsk_kv_len(const lsk_engine* e, int32_t* len_out);
/* Teacher-forced block: the m given ids as one block at positions kv_len .. kv_len+m-1 through
 * every layer and the LM head — `forward` (llama_model_utils.py:155-209) on top of the committed
 * context; logits of the m rows are then readable through LSK_DBG_LOGITS (needs
 * LSK_FLAG_KEEP_LOGITS).  Nothing is committed. */
int lsk_debug_forward_rows(lsk_engine* e, const int32_t* ids_host, int32_t m);
typedef enum {
  LSK_DBG_HIDDEN = 0,        /* fp32 [16, hidden] residual-stream rows of the last launch     */
  LSK_DBG_LOGITS = 1,        /* fp32 [16, vocab_local] (needs LSK_FLAG_KEEP_LOGITS)           */
  LSK_DBG_KROW = 2,          /* bf16->fp32 K cache row [head_dim]: layer, index = kv_head*max_ctx + pos */
  LSK_DBG_VROW = 3,
  LSK_DBG_PROBS_DRAFT = 4,   /* fp32 [16, vocab] warped (T, top-k, top-p) draft distributions     */
  LSK_DBG_PROBS_VERIFY = 5,  /* fp32 [16, vocab] warped verifier distributions of the last round */
  LSK_DBG_RESIDUAL = 6       /* fp32 [vocab] max(p_verify - p_draft, 0) of the last rejected draft (unnormalised; self_speculation_generator.py:27-29 max_fn before its division) */
} lsk_debug_what;
int lsk_debug_read(lsk_engine* e, int32_t what, int32_t layer, int64_t index,
                   float* dst_host, int64_t n_floats);
/* Replace the (identity) logical->physical KV page map with a permutation: proves the paged
 * indirection.  Only valid before lsk_prefill. */
int lsk_debug_set_page_table(lsk_engine* e, const int32_t* pages, int32_t n_pages);

/* Bytes of HBM the engine streams for one (d, ctx) round / AR step — the algorithmic-bytes
 * model of SURVEY.md §8(d), per GPU (tensor-parallel shards included). */
int lsk_round_bytes(const lsk_engine* e, int32_t d, int32_t ctx, double* bytes_out);
int lsk_ar_bytes(const lsk_engine* e, int32_t ctx, double* bytes_out);
/* Kernels launched (or replayed through graphs) since lsk_create; device time of the last
 * lsk_round / lsk_ar_step / lsk_prefill measured with CUDA events on the engine's stream. */
int lsk_launch_count(const lsk_engine* e, int64_t* count_out);
int lsk_last_device_ms(const lsk_engine* e, float* ms_out);

/* Same work as lsk_round, launched eagerly with a CUDA-event pair around every kernel so the
 * device time can be attributed per kernel class: 0 qkv, 1 attention, 2 o-proj, 3 gate/up,
 * 4 down, 5 lm-head, 6 small kernels, 7 collectives.  class_ms / class_launches have 8 entries.
 * Measurement aid for bench.py's roofline section; the numbers include launch gaps that graph
 * replay + programmatic dependent launch hide in lsk_round. */
int lsk_profile_round(lsk_engine* e, int32_t d_req, lsk_round_out* out, float* class_ms,
                      int64_t* class_launches, float* total_ms);

/* Host-side launch schedule of one weight-streaming GEMM (pure host logic; works without a GPU):
 * how many activation columns are resident at a time, tiles accumulated side by side, TMA ring
 * depth, grid.  pro: 0 RMSNorm prologue, 1 bf16 copy; epi: 0 qkv/rope, 1 residual add, 2 store,
 * 3 silu*up, 4 lm-head arg-max. */
typedef struct {
  int32_t ok, nt, tiles_per_pass, n_chunks, chunk_cols, ring_stages, stage_bytes, grid, block;
  int32_t n_tiles;
  int64_t smem_bytes, smem_limit;
} lsk_gemm_plan;
int lsk_plan_gemm(int64_t n_rows, int64_t k, int32_t m, int32_t pro, int32_t epi, int32_t sm_count,
                  lsk_gemm_plan* out);

/* Host-side launch plan of the attention kernel (pure host logic): split-KV factor (an engine
 * constant: results are batch-invariant only for a fixed partition), K/V ring depth, grid, shared
 * memory incl. the one-CTA-per-SM floor, 16-row blocks per CTA.  n_heads / n_kv_heads_local are the
 * tensor-parallel shard's head counts with the same GQA ratio as the model. */
typedef struct {
  int32_t ok, n_splits, ring_stages, grid, block, row_blocks, kv_refetched_per_row_block;
  int64_t smem_bytes, smem_limit;
} lsk_attn_plan;
int lsk_plan_attention(int32_t head_dim, int32_t n_heads, int32_t n_kv_heads_local, int32_t m,
                       int32_t sm_count, lsk_attn_plan* out);

/* Stand-alone kernel entry points used by the micro-benchmarks and unit tests: run the skinny
 * GEMM (y[m, n] = x[m, k] . W[n, k]^T, fp32 out) on packed weights / the split-KV attention on
 * caller-provided device buffers. */
int lsk_test_pack(const void* w_bf16_dev, int64_t n, int64_t k, void* packed_out_dev);
int lsk_test_gemm(const void* packed_dev, int64_t n, int64_t k, const void* x_bf16_dev,
                  int32_t m, float* y_dev, int32_t iters, float* avg_ms_out);
/* Paged split-KV attention alone: m query rows at positions ctx-m .. ctx-1 attend causally to keys
 * 0 .. ctx-1 (modeling_llama.py:187-221 eager attention).  q / out: [m][n_heads*head_dim] bf16,
 * k / v: natural [n_kv_heads][ctx][head_dim] bf16 (k already rotated) — the entry point builds the
 * engine's paged, swizzled pool from them; page_perm_host (nullable) permutes logical->physical
 * pages.  All other pointers are device pointers. */
int lsk_test_attn(const void* q_dev, const void* k_dev, const void* v_dev, int32_t n_heads,
                  int32_t n_kv_heads, int32_t head_dim, int32_t ctx, int32_t m, int32_t n_splits,
                  const int32_t* page_perm_host, void* out_dev, int32_t iters, float* avg_ms_out);
/* tcgen05 LM head (csrc/lmhead_tc.cuh, opt-in): logits[m, n] = rmsnorm(x)[m, :] . W[n, :] with
 * W natural bf16 [n, k], x fp32 [m, k], norm_w bf16 [k]; writes fp32 logits [m, n] and per row the
 * arg-max (lowest index wins).  All pointers are device pointers. */
int lsk_test_lmhead_tc(const void* w_bf16_dev, int64_t n, int64_t k, const float* x_f32_dev,
                       const void* norm_w_bf16_dev, float eps, int32_t m, float* logits_dev,
                       float* best_val_dev, int32_t* best_idx_dev, int32_t iters, float* avg_ms_out);

#ifdef __cplusplus
}
#endif
#endif /* LSK_H_ */
