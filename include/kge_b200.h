/*
 * kge_b200.h -- C ABI of the B200-native KGE hot path (train step + filtered ranking).
 *
 * The reference (bi-graph/Emgraph 1.0.0-rc1) has no native boundary: its hot path is Python over
 * TensorFlow ops.  Each entry point below therefore cites the reference *Python* interface it
 * replaces (paths relative to the reference root).  The Python package `emgraph_b200` binds these
 * with ctypes (emgraph_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would
 * add.
 *
 * Conventions
 *  - plain C types only; every call returns 0 on success, <0 on error; kge_last_error() returns a
 *    thread-local message for the last failing call.
 *  - the CALLER owns every tensor (device pointers, e.g. torch allocations); the library owns only
 *    the opaque kge_ctx (per device: workspace, sort scratch, filter index).
 *  - all work is enqueued on the caller-supplied cudaStream_t (passed as void*); no hidden syncs
 *    except workspace growth (first call at a larger size) and the *_sync getters.
 *  - a ctx is not thread-safe; one ctx per GPU/process.  There is NO CPU fallback.
 *  - embeddings are fp32 row-major [rows, K]; ids are int32; K = k (TransE, DistMult) or 2k
 *    (ComplEx, HolE: row = [re(k) | im(k)], reference models/ComplEx.py:224).
 */
#ifndef KGE_B200_H
#define KGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGE_ABI_VERSION 6
#define KGE_MAX_SHARDS 8

typedef struct kge_ctx kge_ctx;

/* scoring functions: models/TransE.py:190-216, DistMult.py:181-201, ComplEx.py:267-298, HolE.py:169-189 */
enum { KGE_TRANSE_L1 = 0, KGE_TRANSE_L2 = 1, KGE_DISTMULT = 2, KGE_COMPLEX = 3, KGE_HOLE = 4 };
/* losses: losses/pairwise.py:54-70, nll.py:43-59, nll_multiclass.py:57-81, absolute_margin.py:54-70,
 * self_adversarial.py:78-112 */
enum { KGE_LOSS_PAIRWISE = 0, KGE_LOSS_NLL = 1, KGE_LOSS_MULTICLASS_NLL = 2, KGE_LOSS_ABSOLUTE_MARGIN = 3,
       KGE_LOSS_SELF_ADVERSARIAL = 4 };
/* optimizers: training/adam.py:31-48, adagrad.py:30-46, momentum.py:51-69, sgd.py:79-125 */
enum { KGE_OPT_ADAM = 0, KGE_OPT_ADAGRAD = 1, KGE_OPT_MOMENTUM = 2, KGE_OPT_SGD = 3 };
/* training corruption side: evaluation/protocol.py:587-608 ('s,o' == 's+o': per-negative coin) */
enum { KGE_SIDE_SO = 0, KGE_SIDE_S = 1, KGE_SIDE_O = 2 };
/* ranking: corrupt_side (evaluation/protocol.py:883-888), strategy (models/EmbeddingModel.py:1989-2033) */
enum { KGE_RANK_S_O = 0, KGE_RANK_SPO = 1, KGE_RANK_S = 2, KGE_RANK_O = 3 };
enum { KGE_STRAT_WORST = 0, KGE_STRAT_BEST = 1, KGE_STRAT_MIDDLE = 2 };

/* embedding_model_params['non_linearity'] on the scores (models/EmbeddingModel.py:679-689, :801-812, :1868-1881) */
enum { KGE_NL_LINEAR = 0, KGE_NL_TANH = 1, KGE_NL_SIGMOID = 2, KGE_NL_SOFTPLUS = 3 };

/* train-step flags */
#define KGE_F_RESET_STATE 1u /* reference-faithful: optimizer state re-created every batch (training/adam.py:45-46) */
#define KGE_F_NO_UPDATE   2u /* compute loss/grads only (parity tests) */
/* kge_train_step / kge_train_step_host(_async) only: software-pipeline consecutive steps.  The corruption
 * generator and the radix sort of a step depend on the batch and on the (seed, step) counters, never on the
 * parameters, so they run on the library's side stream into the second of two buffer sets and overlap the
 * forward/backward/reduction of the PREVIOUS step; results are bit-identical to the in-order step.  For the
 * device-batch entry the caller promises that a->pos (and neg_entities / repl / keep_subj when given) were
 * resident before the previous step was submitted, i.e. are not produced by work still pending on `stream`
 * (the reference's batches are slices of one resident array, datasets/numpy_adapter.py:105-111). */
#define KGE_F_PIPELINE    4u

/* An embedding table, optionally split by contiguous row range over up to 8 GPUs of one NVSwitch
 * domain.  shard[r] is a device pointer valid on THIS device (local memory for r == own rank, a
 * peer mapping otherwise) holding rows [r*rows_per_shard, min(rows,(r+1)*rows_per_shard)).
 * Replaces the reference's single tf.Variable ent_emb[E,K] (models/EmbeddingModel.py:547-601) and
 * its host-paged "large graph" mode (models/EmbeddingModel.py:645-666, :1070-1097). */
typedef struct kge_table {
    float*  shard[KGE_MAX_SHARDS];
    int64_t rows;
    int64_t rows_per_shard;
    int32_t n_shards;
    int32_t K;
} kge_table;

/* One optimisation step on one batch == EmbeddingModel._get_model_loss + optimizer.minimize
 * (models/EmbeddingModel.py:614-822, :1415-1418). */
typedef struct kge_train_args {
    int32_t  model, loss, opt, side;
    uint32_t flags;
    int32_t  k;            /* user embedding size */
    int32_t  eta;          /* negatives per positive */
    float    margin;       /* pairwise margin (losses/pairwise.py:66) */
    float    lr, beta1, beta2, eps, momentum;
    uint64_t seed;         /* Philox key of the in-kernel corruption generator */
    uint64_t step;         /* 1-based global optimizer step; also the Philox counter high word */
    uint64_t neg_index_base; /* global index of this rank's first negative (multi-GPU streams) */
    kge_table ent, ent_m, ent_v;       /* weights + per-row optimizer state (same sharding) */
    float   *rel, *rel_m, *rel_v;      /* [R,K] replicated */
    int64_t  R;
    const int32_t* pos;    /* [n_pos,3] device */
    int64_t  n_pos;
    const int32_t* repl;       /* optional [eta*n_pos] supplied replacement ids (parity input) */
    const uint8_t* keep_subj;  /* optional [eta*n_pos]: 1 = keep subject, replace object; 0 = the reverse; >= 2 =
                                * decide by `side`.  May be given WITHOUT repl (in-kernel replacements, fixed
                                * sides): a list-valued corrupt_side (models/EmbeddingModel.py:780-816) is run as
                                * one batch holding the positives once per side, see DESIGN.md section 3.1 */
    float*   loss_out;     /* device float[1]: batch loss (overwritten) */
    float*   dbg_scores;   /* optional device [n_pos*(1+eta)]: positives then negatives (row j*n+i) */
    float*   dbg_grad_ent; /* optional device dense [E,K]: summed row gradients (must be zeroed) */
    float*   dbg_grad_rel; /* optional device dense [R,K] */
    /* row-sharded multi-GPU exchange (all optional, NULL on one GPU):
     * stage       local [(2+eta)*n_pos, K] copy of the entity row of every entity slot of this rank's batch
     *             (slot order: subjects, objects, replacements), filled by the owners' kge_train_push_rows;
     *             kge_train_fwd_bwd then reads entity rows from it instead of a->ent (no peer loads).
     * grad_tails  local buffer holding every rank's gradient-buffer tail [Qo | Qs | coef | keep] back to
     *             back (rank r at r*grad_tail_stride floats, all-gathered by the host); kge_train_apply then
     *             reads queries and coefficients locally and only the gs/go/gp rows through `grads`. */
    float*   stage;
    float*   grad_tails;
    int64_t  grad_tail_stride;
    float    alpha;        /* self_adversarial sampling temperature (losses/self_adversarial.py:75) */
    /* LP regulariser (regularizers/lp.py:81-113): loss += lambda_ent*sum|ent|^p + lambda_rel*sum|rel|^p over the
     * WHOLE tables (reg_p = 0: off).  Every row then has a gradient: touched rows get it added in the
     * reduction, the others are updated by a dense pass; the penalty is added to loss_out (by
     * kge_train_apply for the rows of [row_begin,row_end), relations counted where row_begin == 0). */
    int32_t  reg_p;
    float    reg_lambda_ent, reg_lambda_rel;
    /* embedding_model_params['negative_corruption_entities'] (models/EmbeddingModel.py:732-777,
     * evaluation/protocol.py:610-641): in-kernel replacements are drawn uniformly from neg_entities
     * [neg_entities_n] (device int32 ids: a supplied list, or the batch's unique entities), or from the first
     * neg_entities_n entities when neg_entities is NULL and neg_entities_n > 0; 0 / NULL = all entities. */
    const int32_t* neg_entities;
    int64_t  neg_entities_n;
    int32_t  non_linearity; /* KGE_NL_*: applied to positive and negative scores before the loss */
    /* dimension-sharded multi-GPU step: the tables passed are this rank's COLUMN slice, k the columns per half of the
     * slice, k_model the k of the whole model (HolE's 2/k score scale, models/HolE.py:189); 0 = k */
    int32_t  k_model;
} kge_train_args;

int         kge_abi_version(void);
/* 1 if this build carries the tcgen05 3xTF32 ranking sweep (use_tensor_cores=1 in kge_rank_counts) */
int         kge_has_tensor_core_rank(void);
/* floats in the caller-owned gradient buffer kge_train_fwd_bwd writes for a batch of n_pos positives:
 * 5 rows per positive (grad s, grad o, grad p, query Qo, query Qs) + eta coefficients + eta side flags */
int64_t     kge_train_grad_floats(int eta, int64_t n_pos, int K);
/* floats of that buffer before its [Qo | Qs | coef | keep] tail (= 3*n_pos*K: the gs, go, gp rows) */
int64_t     kge_train_grad_head_floats(int eta, int64_t n_pos, int K);
const char* kge_last_error(void);

/* per-device context (replaces the reference's implicit TF runtime state) */
int kge_ctx_create(int device, kge_ctx** out);
int kge_ctx_destroy(kge_ctx* ctx);
/* Bench instrumentation: when on, kge_train_step brackets its phases with CUDA events on the caller's
 * stream; kge_ctx_get_timing returns the average ms per step of {emit, fwd_bwd, sort-wait + reduce_apply,
 * span/hub reduction, end of the side-stream radix sort measured from the end of emit} since timing was
 * switched on (synchronises). */
int kge_ctx_set_timing(kge_ctx* ctx, int on);
int kge_ctx_get_timing(kge_ctx* ctx, float* ms_out5, int* steps_out);
/* the same with the wait split off the reduction: n_out >= 6, ms_out = {emit, fwd_bwd, wait for the sort after fwd_bwd,
 * level-1 reduction kernel alone, span/hub reduction, end of the sort measured from the end of emit} */
int kge_ctx_get_timing_ex(kge_ctx* ctx, float* ms_out, int n_out, int* steps_out);
/* Test hook: stable sort of n (key << 32 | slot) entries by key (every key < n_keys), the order the segmented reduction
 * walks.  algo 0: radix sort on the key bits; 1: the single-launch two-pass radix sort used for small batches over small key
 * ranges (csrc/kge_sort_small.cu; error when the size is outside its range).  Both give the same output. */
int kge_sort_entries(kge_ctx* ctx, const uint64_t* in, int64_t n, int64_t n_keys, int algo, uint64_t* out, void* stream);
/* bytes of device workspace currently held by the ctx */
int64_t kge_ctx_workspace_bytes(kge_ctx* ctx);

/* EmbeddingModel.predict / _lookup_embeddings + _fn  (models/EmbeddingModel.py:2101-2147, :490-533) */
int kge_score(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
              const int32_t* triples, int64_t n, float* out, void* stream);
/* the same with embedding_model_params['non_linearity'] applied to the scores, as EmbeddingModel.predict returns them
 * (models/EmbeddingModel.py:2135-2147; KGE_NL_*).  Ids are not range-checked: the caller validates 0 <= s,o < E, 0 <= p < R. */
int kge_predict(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                const int32_t* triples, int64_t n, int non_linearity, float* out, void* stream);

/* Whole step on one GPU (n_shards == 1): corruption generation -> fused score/loss/backward ->
 * duplicate-row segmented reduction -> sparse row-wise optimizer. */
int kge_train_step(kge_ctx* ctx, const kge_train_args* a, void* stream);

/* The same step in three phases, for the row-sharded multi-GPU path where the host places
 * NCCL collectives (key all-gather / barriers) between them:
 *  1. kge_train_emit: draw corruptions, write this rank's sort keys (entity id, or E + relation id)
 *     into keys_out[n_slots] with n_slots = (3+eta)*n_pos; slot layout documented in DESIGN.md.
 *  2. kge_train_fwd_bwd: fused forward + loss + backward into grad_buf (kge_train_grad_floats floats,
 *     caller-owned so that peers can map it): the gradient rows of the positive's s, o, p, its two
 *     folded queries and one coefficient per negative -- the gradient row of a replacement entity is
 *     re-materialised from (coefficient, query row, current row) by the reduction, never stored.
 *  3. kge_train_apply: sort all ranks' keys, segmented-reduce duplicate rows reading through `grads`
 *     (shard r = rank r's grad_buf, rows_per_shard = n_slots per rank, every rank the same n_pos) and
 *     apply the optimizer to rows in [row_begin,row_end) and to every relation row. */
int kge_train_emit(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, void* stream);
int kge_train_fwd_bwd(kge_ctx* ctx, const kge_train_args* a, float* grad_buf, void* stream);
int kge_train_apply(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                    const kge_table* grads, int64_t row_begin, int64_t row_end, void* stream);
/* Owner-side push of the rows the other ranks' batches need (replaces fine-grained peer loads: a
 * peer's random 1 KiB row reads through an IPC mapping of a multi-GB shard run at ~130 GB/s on B200,
 * contiguous peer stores at ~710 GB/s).  For every entity slot t of keys_all (all ranks' keys, n_keys =
 * n_ranks * (3+eta)*n_pos) whose key lies in [row_begin,row_end): copy row `key` of the local shard of
 * a->ent into stage->shard[t / slots_per_rank] at row (t % slots_per_rank).  stage->shard[r] = rank r's
 * staging buffer (local or peer mapping), stage->rows_per_shard = (2+eta)*n_pos.  The caller places a
 * cross-rank barrier between this and kge_train_fwd_bwd with a->stage set. */
int kge_train_push_rows(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                        const kge_table* stage, int64_t row_begin, int64_t row_end, void* stream);
/* Optional, between the key all-gather and kge_train_fwd_bwd: start selecting the slots this rank
 * will reduce (keys in [row_begin,row_end) + relation keys) so that their count reaches the host while
 * the forward/backward kernel runs; kge_train_apply with the same arguments then does not stall. */
int kge_train_select(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                     int64_t row_begin, int64_t row_end, void* stream);

/* Dimension-sharded ("column-parallel") multi-GPU step: the table of a model too large or too slow for one GPU is
 * split by COLUMN range over the GPUs of one NVSwitch domain -- rank r holds ent[E, Kc], rel[R, Kc] and the
 * optimizer state of columns [r*Kc, (r+1)*Kc) of every row (for ComplEx / HolE: the same range of the real and of
 * the imaginary half, stored [re | im]) -- and every rank processes the WHOLE global batch on its slice.  Every
 * scoring function of models/{TransE,DistMult,ComplEx,HolE}.py is a sum over columns (TransE norm 2: the squared
 * distance is), so the ranks exchange one sum per scored triple and nothing else:
 *   kge_train_partial   draws the corruptions of the global batch (same Philox stream on every rank), starts the
 *                       sort of the slot keys, writes the slice's raw partial sums of positives [i_begin,i_end):
 *                       sums[0,nc) positives, sums[nc + j*nc + (i-i_begin)] negative (j,i), nc = i_end-i_begin
 *   (caller)            all-reduce(sum) of `sums` over the ranks (NCCL / peer memory); 4*(1+eta) bytes per positive
 *   kge_train_backward  scores from the totals, loss terms, dL/dscore, gradient rows of the slice (ctx-owned buffer)
 *   kge_train_reduce    batch loss (identical on every rank) -> a->loss_out; duplicate-row segmented reduction +
 *                       sparse optimizer on the slice: the single-GPU kernels, unchanged
 * Chunks [i_begin,i_end) of one step must be submitted in order and cover [0,n_pos); i_begin == 0 starts a step.
 * a->ent.n_shards == 1, a->k = columns per half of the slice (a multiple of 4), a->k_model = the model's k.
 * Replaces the reference's host-paged "large graph" mode (models/EmbeddingModel.py:645-666, :1070-1097, :1251-1281). */
int kge_train_partial(kge_ctx* ctx, const kge_train_args* a, int64_t i_begin, int64_t i_end, float* sums, void* stream);
/* kge_train_partial for the whole batch at once with the entity rows streamed in SORTED order (each row once, in address
 * order, instead of (3+eta) random 4*Kc-byte reads per positive).  `sums` takes the pieces of n_chunks consecutive positive
 * ranges back to back -- piece c covers positives [lo_c, lo_c + nc_c) with nc_c = n_pos / n_chunks (+1 for the first
 * n_pos % n_chunks pieces) and starts at float (1+eta)*lo_c, laid out as kge_train_partial lays out one range -- so that the
 * caller can all-reduce piece c and hand it to kge_train_backward(lo_c, lo_c + nc_c) while later pieces are in flight. */
int kge_train_partial_sorted(kge_ctx* ctx, const kge_train_args* a, int n_chunks, float* sums, void* stream);
int kge_train_backward(kge_ctx* ctx, const kge_train_args* a, int64_t i_begin, int64_t i_end, const float* sums, void* stream);
int kge_train_reduce(kge_ctx* ctx, const kge_train_args* a, void* stream);

/* The all-reduce(sum) of the step above over peer memory instead of NCCL: floats [off, off+len) of every rank's `sums` buffer
 * are summed in rank order and the totals written into every rank's `totals` buffer.  sums / totals / flags: shard[p] = rank p's
 * buffer (own memory or a CUDA-IPC mapping, kge_ipc_open); flags: 16 uint32 words per rank, zeroed once before the first call;
 * seq: a number that grows with every call (the same on every rank).  Every rank must call it with the same range and seq; the
 * kernel returns (on the stream) when this rank's totals are complete.  off and len are multiples of 4. */
int kge_allreduce_p2p(kge_ctx* ctx, const kge_table* sums, const kge_table* totals, const kge_table* flags, int rank,
                      int64_t off, int64_t len, uint32_t seq, void* stream);

/* Host-buffer form of kge_train_step: what a reference-side caller binds.  The reference feeds every
 * batch from host numpy through tf.data (models/EmbeddingModel.py:1329-1337, :1044-1111) and reads
 * the batch loss back with .numpy() (:1421).  pos_host [n_pos,3] int32 (pinned for an async copy) is
 * copied to a ctx-owned device buffer (a->pos is ignored), the step runs, the loss is copied to
 * *loss_host and the stream is synchronised before returning. */
int kge_train_step_host(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                        void* stream);
/* The same call split in two so that a training loop can keep the next step queued while it reads the loss of
 * the previous one (the copies and the step are enqueued exactly as above, nothing is skipped):
 * kge_train_step_host_async enqueues [H2D of the batch, the step, D2H of the loss into *loss_host] and returns
 * a ticket; kge_train_host_wait(ticket) blocks until that step, including its loss copy, has finished.  Up to
 * 3 steps may be in flight (the 4th call waits for the oldest); pos_host / loss_host of an in-flight step must
 * stay valid and distinct (pinned). */
int kge_train_step_host_async(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                              void* stream, int* ticket_out);
int kge_train_host_wait(kge_ctx* ctx, int ticket);

/* optional post-step row renormalisation (models/EmbeddingModel.py:1434-1439, clip_by_norm axes=1) */
int kge_normalize_rows(kge_ctx* ctx, float* emb, int64_t rows, int K, void* stream);

/* Known-triple filter: replaces SQLiteAdapter (datasets/sqlite_adapter.py:54-98, :234-262) and the
 * per-test-triple queries of get_participating_entities (:449-508). Device-resident sorted
 * (s,p)->objects and (p,o)->subjects indexes, deduplicated. */
int kge_filter_build(kge_ctx* ctx, const int32_t* triples, int64_t F, int64_t E, int64_t R, void* stream);
int kge_filter_clear(kge_ctx* ctx);
/* number of distinct filter triples (host sync) */
int64_t kge_filter_size_sync(kge_ctx* ctx);

/* evaluate_performance / get_ranks (evaluation/protocol.py:726-979, models/EmbeddingModel.py:2046-2099,
 * :1845-1986): all-entity sweep for every test triple with in-kernel x1e5 int quantisation (F7),
 * filter applied in-kernel, rank count fused.  Sweeps candidate rows [row_begin,row_end) of the
 * LOCAL shard `ent_local` (row 0 of ent_local == global row row_begin) and writes
 * counts[T,2,4] int32 (side {0:object sweep, 1:subject sweep} x {gt, eq, gt_filtered, eq_filtered});
 * counts from all shards are summed by the caller (NCCL all-reduce) before kge_rank_finalize.
 * side (KGE_RANK_*) selects which sweeps run (KGE_RANK_S / KGE_RANK_O run one).
 * use_tensor_cores: 1 = tcgen05 3xTF32 path (DistMult/ComplEx/HolE), 0 = fp32 CUDA-core sweep.
 * non_linearity (KGE_NL_*) is applied to every score before the x1e5 quantisation (models/EmbeddingModel.py:1868-1881). */
int kge_rank_counts(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                    const float* ent_local, int64_t row_begin, int64_t row_end,
                    const int32_t* test, int64_t T, int side, int filtered, int use_tensor_cores,
                    int non_linearity, int32_t* counts, void* stream);
/* The same sweep for a table whose rows this process cannot address (column-sharded training, row-range shards for
 * ranking): the caller supplies the subject and object rows of the test triples (s_rows, o_rows: [T,K] each, gathered
 * from their owners) and E, the number of entities of the whole table; everything else as kge_rank_counts. */
int kge_rank_counts_rows(kge_ctx* ctx, int model, int k, int64_t E, const float* rel, int64_t R,
                         const float* s_rows, const float* o_rows, const float* ent_local, int64_t row_begin,
                         int64_t row_end, const int32_t* test, int64_t T, int side, int filtered,
                         int use_tensor_cores, int non_linearity, int32_t* counts, void* stream);
/* ranks_out: [T,2] (col 0 subject, col 1 object) for KGE_RANK_S_O, else [T].
 * self_is_candidate: optional device uint8 [T,2] (col 0 subject, col 1 object): 0 when the test triple's own
 * entity was NOT among the swept candidates (entities_subset ranking, models/EmbeddingModel.py:1845-1857,
 * :1898-1940); NULL = every test entity is a candidate (all-entity sweep). */
int kge_rank_finalize(kge_ctx* ctx, const int32_t* counts, int64_t T, int side, int strategy,
                      int filtered, const uint8_t* self_is_candidate, int32_t* ranks_out, void* stream);

/* Host-buffer form of evaluate_performance's device work for a single-GPU table: test_host [T,3]
 * int32 in, ranks_host ([T,2] for KGE_RANK_S_O else [T]) out; synchronises the stream. */
int kge_rank_host(kge_ctx* ctx, int model, int k, const kge_table* ent, const float* rel, int64_t R,
                  const int32_t* test_host, int64_t T, int side, int strategy, int filtered,
                  int use_tensor_cores, int non_linearity, int32_t* ranks_host, void* stream);

/* Device memory that peers can map (plain cudaMalloc) + CUDA IPC helpers for mapping peer shards
 * (one process per GPU). handle: 64 bytes. */
int kge_dev_alloc(int64_t bytes, void** out);
int kge_dev_free(void* p);
int kge_ipc_export(void* dev_ptr, void* handle_out64);
int kge_ipc_open(const void* handle64, void** dev_ptr_out);
int kge_ipc_close(void* dev_ptr);
int kge_enable_peer_access(int device, int peer);

#ifdef __cplusplus
}
#endif
#endif /* KGE_B200_H */
