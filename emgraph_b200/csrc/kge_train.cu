// KGE training step for sm_100a: corruption generation, sort keys, segmented reduction of duplicate
// rows and the sparse row-wise optimizer.  The fused forward/loss/backward kernel lives in
// kge_train_fwd.cuh (one translation unit per scoring model).
//
//   emit        : counter-based corruption generation (Philox4x32-10) + per-slot sort keys
//   fwd_bwd     : kge_train_fwd.cuh
//   reduce_apply: radix sort of (row id, slot) + atomics-free two-level segmented reduction:
//                 every warp owns KGE_CH consecutive sorted slots; runs that live inside one chunk
//                 are reduced and fed straight into Adam/Adagrad/momentum/SGD for that row; runs that
//                 cross chunk borders (hub entities) leave one partial row per chunk and are finished
//                 by kge_span_apply_kernel in chunk order.  Summation order is fixed by the sort =>
//                 bit-reproducible; only touched rows are read or written.
//
// Replaces reference evaluation/protocol.py:531-659 (generate_corruptions_for_fit) and
// training/{adam,adagrad,momentum,sgd}.py (+ the Keras OptimizerV2 sparse apply with its duplicate
// index summation).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "kge_train_fwd.cuh"

// ------------------------------------------------------------------------------------------------
// slot layout for one batch of n positives (S = (3+eta)*n slots)
//   [0,n)            subject row of positive i          key = s_i         gradient: gbuf row i      (gs)
//   [n,2n)           object row of positive i           key = o_i         gradient: gbuf row n+i    (go)
//   [2n,2n+eta*n)    replacement row of negative (j,i)  key = repl[j*n+i] gradient: F(coef, Q row, r)
//   [2n+eta*n,S)     relation row of positive i         key = E + p_i     gradient: gbuf row 2n+i   (gp)
// ------------------------------------------------------------------------------------------------

__global__ void kge_emit_kernel(const int32_t* __restrict__ pos, int64_t n, int eta, int64_t E, int side,
                                const int32_t* __restrict__ repl_in, const uint8_t* __restrict__ keep_in,
                                uint64_t seed, uint64_t step, uint64_t neg_base,
                                int32_t* __restrict__ repl_out, uint8_t* __restrict__ keep_out,
                                int32_t* __restrict__ keys, uint64_t* __restrict__ packed) {
    int64_t S = (int64_t)(3 + eta) * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < S; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t key;
        if (t < n) {
            key = pos[3 * t + 0];
        } else if (t < 2 * n) {
            key = pos[3 * (t - n) + 2];
        } else if (t < 2 * n + (int64_t)eta * n) {
            int64_t q = t - 2 * n;  // j*n + i
            int32_t r;
            uint8_t ks;
            if (repl_in != nullptr) {
                r = repl_in[q];
                ks = keep_in != nullptr ? keep_in[q] : (side == KGE_SIDE_O ? 1 : 0);
            } else {
                uint64_t g = neg_base + (uint64_t)q;
                uint32_t o[4];
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32), o);
                // replacement ~ U{0..E-1} (evaluation/protocol.py:616-619); multiply-shift mapping
                r = (int32_t)(((uint64_t)o[0] * (uint64_t)E) >> 32);
                ks = side == KGE_SIDE_SO ? (uint8_t)(o[1] >> 31) : (side == KGE_SIDE_O ? 1 : 0);
            }
            repl_out[q] = r;
            keep_out[q] = ks;
            key = r;
        } else {
            key = (int32_t)E + pos[3 * (t - 2 * n - (int64_t)eta * n) + 1];
        }
        if (keys != nullptr) keys[t] = key;
        if (packed != nullptr) packed[t] = ((uint64_t)(uint32_t)key << 32) | (uint64_t)(uint32_t)t;
    }
}

// deterministic fixed-order reduction of the per-positive loss terms
__global__ void kge_loss_reduce_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out) {
    __shared__ double sm[1024];
    double acc = 0.0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) acc += (double)part[t];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)sm[0];
}

// ------------------------------------------------------------------------------------------------
// segmented reduction of duplicate rows + sparse row-wise optimizer
// ------------------------------------------------------------------------------------------------
#define KGE_CH 16  // sorted slots per warp

struct GradView {
    float*  base[KGE_MAX_SHARDS];  // rank r's gradient buffer (local or peer mapping)
    int64_t S;                     // slots per rank
    int64_t n;                     // positives per rank
    int     eta, K, n_ranks;
};

struct ApplyParams {
    const uint64_t* ks;    // sorted (key << 32 | global slot id), global slot id = rank*S + local slot
    int64_t n_keys;
    GradView G;
    TableView ent, ent_m, ent_v;
    float *rel, *rel_m, *rel_v;
    int64_t E, R;
    int64_t row_begin, row_end;  // owned entity rows
    int opt;
    bool has_m, has_v;     // entity optimizer-state tables present
    uint32_t flags;
    float lr, lr_t, beta1, beta2, eps, momentum;
    float* partial;        // [2*n_chunks][K]
    uint8_t* span_head;    // [n_chunks]
    float* dbg_grad_ent;
    float* dbg_grad_rel;
};

struct SlotMeta {
    const float* row;
    float c;
    int mode;  // 0: add row ; 1: replacement row of a negative, F(c, Q, r)
};

__device__ __forceinline__ SlotMeta decode_slot(const GradView& G, int32_t slot) {
    int rr = 0;
    int64_t t = slot;
    if (G.n_ranks > 1) {
        rr = (int)(t / G.S);
        t -= (int64_t)rr * G.S;
    }
    float* base = G.base[rr];
    const int64_t n = G.n;
    SlotMeta m;
    m.c = 1.f;
    m.mode = 0;
    if (t < 2 * n) {
        m.row = base + t * G.K;
    } else if (t < 2 * n + (int64_t)G.eta * n) {
        const int64_t q = t - 2 * n;
        const int64_t i = q % n;
        const float* coef = gbuf_coef(base, n, G.K);
        const uint8_t* keep = gbuf_keep(base, G.eta, n, G.K);
        m.c = coef[q];
        m.row = base + ((keep[q] ? 3 : 4) * n + i) * G.K;
        m.mode = 1;
    } else {
        m.row = base + (2 * n + (t - 2 * n - (int64_t)G.eta * n)) * G.K;
    }
    return m;
}

struct RowPtrs {
    float *w, *m, *v;
    bool is_rel, owned;
    int64_t row;
};

__device__ __forceinline__ RowPtrs resolve_row(const ApplyParams& P, int32_t key) {
    RowPtrs r;
    r.is_rel = key >= P.E;
    r.row = r.is_rel ? key - P.E : key;
    r.owned = r.is_rel || (r.row >= P.row_begin && r.row < P.row_end);
    r.m = r.v = nullptr;
    const int K = P.ent.K;
    if (r.is_rel) {
        r.w = P.rel + (size_t)r.row * K;
        if (P.rel_m) r.m = P.rel_m + (size_t)r.row * K;
        if (P.rel_v) r.v = P.rel_v + (size_t)r.row * K;
    } else {
        r.w = table_row(P.ent, r.row);
        if (P.has_m) r.m = table_row(P.ent_m, r.row);
        if (P.has_v) r.v = table_row(P.ent_v, r.row);
    }
    return r;
}

// optimizer update of V consecutive columns of one row given the summed gradient g
template <int V>
__device__ __forceinline__ void opt_update(const ApplyParams& P, const RowPtrs& r, int c0, const float (&g)[V], const float (&w_in)[V]) {
    if (r.is_rel ? (P.dbg_grad_rel != nullptr) : (P.dbg_grad_ent != nullptr)) {
        float* d = (r.is_rel ? P.dbg_grad_rel : P.dbg_grad_ent) + (size_t)r.row * P.ent.K + c0;
        st_vec<V>(d, g);
    }
    if (P.flags & KGE_F_NO_UPDATE) return;
    const bool reset = (P.flags & KGE_F_RESET_STATE) != 0;
    float wv[V], mv[V], vv[V];
#pragma unroll
    for (int x = 0; x < V; ++x) wv[x] = w_in[x];
    if (P.opt == KGE_OPT_ADAM) {
        // Keras Adam (beta1 .9, beta2 .999, eps 1e-7): var -= lr_t * m / (sqrt(v) + eps)
        if (reset) {
#pragma unroll
            for (int x = 0; x < V; ++x) mv[x] = vv[x] = 0.f;
        } else {
            ld_vec<V>(mv, r.m + c0);
            ld_vec<V>(vv, r.v + c0);
        }
#pragma unroll
        for (int x = 0; x < V; ++x) {
            mv[x] = P.beta1 * mv[x] + (1.f - P.beta1) * g[x];
            vv[x] = P.beta2 * vv[x] + (1.f - P.beta2) * g[x] * g[x];
            wv[x] = wv[x] - P.lr_t * mv[x] / (sqrtf(vv[x]) + P.eps);
        }
        if (r.m) st_vec<V>(r.m + c0, mv);
        if (r.v) st_vec<V>(r.v + c0, vv);
    } else if (P.opt == KGE_OPT_ADAGRAD) {
        // Keras Adagrad: accumulator starts at 0.1; var -= lr * g / (sqrt(acc) + eps)
        if (reset) {
#pragma unroll
            for (int x = 0; x < V; ++x) mv[x] = 0.1f;
        } else {
            ld_vec<V>(mv, r.m + c0);
        }
#pragma unroll
        for (int x = 0; x < V; ++x) {
            mv[x] = mv[x] + g[x] * g[x];
            wv[x] = wv[x] - P.lr * g[x] / (sqrtf(mv[x]) + P.eps);
        }
        if (r.m) st_vec<V>(r.m + c0, mv);
    } else if (P.opt == KGE_OPT_MOMENTUM) {
        // Keras SGD momentum: vel = mu*vel - lr*g ; var += vel
        if (reset) {
#pragma unroll
            for (int x = 0; x < V; ++x) mv[x] = 0.f;
        } else {
            ld_vec<V>(mv, r.m + c0);
        }
#pragma unroll
        for (int x = 0; x < V; ++x) {
            mv[x] = P.momentum * mv[x] - P.lr * g[x];
            wv[x] = wv[x] + mv[x];
        }
        if (r.m) st_vec<V>(r.m + c0, mv);
    } else {
#pragma unroll
        for (int x = 0; x < V; ++x) wv[x] = wv[x] - P.lr * g[x];
    }
    st_vec<V>(r.w + c0, wv);
}

// contribution of one slot to V columns of the gradient; rc = current value of the row being updated
template <int V, int TMODE>
__device__ __forceinline__ void add_slot(float (&g)[V], const float (&a)[V], float c, int mode, const float (&rc)[V]) {
#pragma unroll
    for (int x = 0; x < V; ++x) {
        if (mode == 0) {
            g[x] += a[x];
        } else if (TMODE == 0) {
            g[x] = fmaf(c, a[x], g[x]);  // DistMult / ComplEx / HolE: c*Q
        } else if (TMODE == 1) {
            float d = a[x] - rc[x];  // TransE L1: c*sign(Q-r)
            g[x] += d > 0.f ? c : (d < 0.f ? -c : 0.f);
        } else {
            g[x] = fmaf(c, a[x] - rc[x], g[x]);  // TransE L2: c*(Q-r)
        }
    }
}

// Level 1: one warp per chunk of KGE_CH sorted slots.
template <int V, int TMODE>
__global__ void __launch_bounds__(256) kge_reduce_apply_kernel(ApplyParams P) {
    __shared__ SlotMeta meta[8][KGE_CH];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * 8 + wib;
    const int64_t b0 = w * KGE_CH;
    if (b0 >= P.n_keys) return;
    const int cnt = (int)min((int64_t)KGE_CH, P.n_keys - b0);
    const int K = P.ent.K;

    int32_t key = -2;
    if (lane < cnt) {
        const uint64_t kv = P.ks[b0 + lane];
        key = (int32_t)(kv >> 32);
        meta[wib][lane] = decode_slot(P.G, (int32_t)(kv & 0xffffffffu));
    }
    int32_t key_prev = -1, key_next = -1;
    if (lane == 0 && b0 > 0) key_prev = (int32_t)(P.ks[b0 - 1] >> 32);
    if (lane == 1 && b0 + cnt < P.n_keys) key_next = (int32_t)(P.ks[b0 + cnt] >> 32);
    key_prev = __shfl_sync(0xffffffffu, key_prev, 0);
    key_next = __shfl_sync(0xffffffffu, key_next, 1);
    const int32_t left = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane < cnt && (lane == 0 || key != left);
    unsigned heads = __ballot_sync(0xffffffffu, head);
    __syncwarp();

    bool span_head = false;
    while (heads) {
        const int a = __ffs(heads) - 1;
        heads &= heads - 1;
        const int b = heads ? (__ffs(heads) - 1) : cnt;
        const int32_t skey = __shfl_sync(0xffffffffu, key, a);
        const bool open_start = (a == 0) && (skey == key_prev);
        const bool open_end = (b == cnt) && (skey == key_next);
        const RowPtrs r = resolve_row(P, skey);
        if (!r.owned) continue;
        if (!open_start && open_end) span_head = true;
        const bool complete = !open_start && !open_end;
        float* part = P.partial + ((size_t)(2 * w + (open_start ? 0 : 1))) * K;
        for (int c0 = lane * V; c0 < K; c0 += 32 * V) {
            float g[V], rc[V];
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] = rc[x] = 0.f;
            // the row's current value: needed by the optimizer and by the TransE slot gradients
            if (complete || TMODE != 0) ld_vec<V>(rc, r.w + c0);
            int u = a;
            for (; u + 4 <= b; u += 4) {
                float v4[4][V];
#pragma unroll
                for (int q = 0; q < 4; ++q) ld_vec<V>(v4[q], meta[wib][u + q].row + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) add_slot<V, TMODE>(g, v4[q], meta[wib][u + q].c, meta[wib][u + q].mode, rc);
            }
            for (; u < b; ++u) {
                float v1[V];
                ld_vec<V>(v1, meta[wib][u].row + c0);
                add_slot<V, TMODE>(g, v1, meta[wib][u].c, meta[wib][u].mode, rc);
            }
            if (complete) opt_update<V>(P, r, c0, g, rc);
            else st_vec<V>(part + c0, g);
        }
    }
    if (lane == 0) P.span_head[w] = span_head ? 1 : 0;
}

// Level 2: one CTA per chunk that starts a run crossing chunk borders; threads own columns and add
// the per-chunk partial rows in chunk order.
template <int V>
__global__ void __launch_bounds__(128) kge_span_apply_kernel(ApplyParams P) {
    const int64_t w = blockIdx.x;
    if (!P.span_head[w]) return;
    const int lane = threadIdx.x & 31;
    const int K = P.ent.K;
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    const int32_t key = (int32_t)(P.ks[w * KGE_CH + KGE_CH - 1] >> 32);
    // last chunk whose first slot still carries `key` (32 chunks probed per round)
    int64_t last = w;
    for (;;) {
        const int64_t c = last + 1 + lane;
        const bool same = c < n_chunks && (int32_t)(P.ks[c * KGE_CH] >> 32) == key;
        const unsigned mk = __ballot_sync(0xffffffffu, same);
        if (mk == 0xffffffffu) {
            last += 32;
            continue;
        }
        last += __ffs(~mk) - 1;
        break;
    }
    const RowPtrs r = resolve_row(P, key);
    for (int c0 = threadIdx.x * V; c0 < K; c0 += 128 * V) {
        float g[V];
        ld_vec<V>(g, P.partial + (size_t)(2 * w + 1) * K + c0);
        int64_t c = w + 1;
        for (; c + 8 <= last + 1; c += 8) {
            float v8[8][V];
#pragma unroll
            for (int q = 0; q < 8; ++q) ld_vec<V>(v8[q], P.partial + (size_t)(2 * (c + q)) * K + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int x = 0; x < V; ++x) g[x] += v8[q][x];
        }
        for (; c <= last; ++c) {
            float v1[V];
            ld_vec<V>(v1, P.partial + (size_t)(2 * c) * K + c0);
#pragma unroll
            for (int x = 0; x < V; ++x) g[x] += v1[x];
        }
        float rc[V];
        ld_vec<V>(rc, r.w + c0);
        opt_update<V>(P, r, c0, g, rc);
    }
}

__global__ void kge_normalize_rows_kernel(float* emb, int64_t rows, int K) {
    // tf.clip_by_norm(ent_emb, clip_norm=1, axes=1)  (models/EmbeddingModel.py:1434-1439)
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float* p = emb + (size_t)r * K;
    float acc = 0.f;
    for (int c = lane; c < K; c += 32) acc = fmaf(p[c], p[c], acc);
    acc = warp_sum(acc);
    float nrm = sqrtf(acc);
    if (nrm > 1.f) {
        float inv = 1.f / nrm;
        for (int c = lane; c < K; c += 32) p[c] *= inv;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int validate_train(const kge_train_args* a) {
    KGE_REQUIRE(a != nullptr, "kge_train: null args");
    KGE_REQUIRE(a->model >= KGE_TRANSE_L1 && a->model <= KGE_HOLE, "kge_train: unknown model %d", a->model);
    KGE_REQUIRE(a->loss >= KGE_LOSS_PAIRWISE && a->loss <= KGE_LOSS_MULTICLASS_NLL, "Unsupported loss function: %d", a->loss);
    KGE_REQUIRE(a->opt >= KGE_OPT_ADAM && a->opt <= KGE_OPT_SGD, "Unsupported optimizer: %d", a->opt);
    KGE_REQUIRE(a->side >= KGE_SIDE_SO && a->side <= KGE_SIDE_O, "Invalid corruption side %d", a->side);
    KGE_REQUIRE(a->k > 0 && a->eta > 0, "kge_train: k and eta must be positive");
    KGE_REQUIRE(a->ent.K == model_row_width(a->model, a->k), "kge_train: table width %d != internal_k %d", a->ent.K,
                model_row_width(a->model, a->k));
    KGE_REQUIRE(a->ent.n_shards >= 1 && a->ent.n_shards <= KGE_MAX_SHARDS, "kge_train: bad shard count");
    KGE_REQUIRE(a->ent.rows + a->R < (int64_t)INT32_MAX, "kge_train: E+R must fit int32 sort keys");
    KGE_REQUIRE(a->n_pos >= 0 && (a->pos != nullptr || a->n_pos == 0), "kge_train: positives missing");
    return 0;
}

static int ensure_train_ws(kge_ctx* ctx, const kge_train_args* a) {
    int64_t n = a->n_pos;
    if (ctx->repl.reserve((size_t)a->eta * n * sizeof(int32_t))) return -2;
    if (ctx->keep.reserve((size_t)a->eta * n)) return -2;
    if (ctx->loss_part.reserve((size_t)(n + 1) * sizeof(float))) return -2;
    return 0;
}

extern "C" int64_t kge_train_grad_floats(int eta, int64_t n_pos, int K) { return gbuf_floats(eta, n_pos, K); }

static int emit_impl(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, uint64_t* packed_out, cudaStream_t st) {
    if (int rc = ensure_train_ws(ctx, a)) return rc;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    int threads = 256;
    int blocks = (int)std::min<int64_t>((S + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_emit_kernel<<<blocks, threads, 0, st>>>(a->pos, a->n_pos, a->eta, a->ent.rows, a->side, a->repl, a->keep_subj,
                                                a->seed, a->step, a->neg_index_base, ctx->repl.as<int32_t>(),
                                                ctx->keep.as<uint8_t>(), keys_out, packed_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int kge_train_emit(kge_ctx* ctx, const kge_train_args* a, int32_t* keys_out, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_emit: null ctx");
    if (int rc = validate_train(a)) return rc;
    if (a->n_pos == 0) return 0;
    KGE_REQUIRE(keys_out != nullptr, "kge_train_emit: keys_out missing");
    ctx->sel_valid = false;
    return emit_impl(ctx, a, keys_out, nullptr, (cudaStream_t)stream);
}

extern "C" int kge_train_fwd_bwd(kge_ctx* ctx, const kge_train_args* a, float* grad_buf, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_fwd_bwd: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(grad_buf != nullptr, "kge_train_fwd_bwd: grad_buf missing");
    KGE_REQUIRE(a->loss_out != nullptr, "kge_train_fwd_bwd: loss_out missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_pos == 0) {
        KGE_CUDA_CHECK(cudaMemsetAsync(a->loss_out, 0, sizeof(float), st));
        return 0;
    }
    FwdBwdParams P;
    P.ent = make_view(a->ent);
    P.rel = a->rel;
    P.pos = a->pos;
    P.repl = ctx->repl.as<int32_t>();
    P.keep = ctx->keep.as<uint8_t>();
    P.n = a->n_pos;
    P.eta = a->eta;
    P.k = a->k;
    P.loss = a->loss;
    P.margin = a->margin;
    P.scale = a->model == KGE_HOLE ? 2.0f / (float)a->k : 1.0f;
    P.gbuf = grad_buf;
    P.loss_part = ctx->loss_part.as<float>();
    P.dbg_scores = a->dbg_scores;
    int rc;
    switch (a->model) {
        case KGE_TRANSE_L1: rc = kge_launch_fwd_bwd_m0(P, ctx->sm_count, st); break;
        case KGE_TRANSE_L2: rc = kge_launch_fwd_bwd_m1(P, ctx->sm_count, st); break;
        case KGE_DISTMULT: rc = kge_launch_fwd_bwd_m2(P, ctx->sm_count, st); break;
        default: rc = kge_launch_fwd_bwd_m3(P, ctx->sm_count, st); break;
    }
    if (rc) return rc;
    kge_loss_reduce_kernel<<<1, 1024, 0, st>>>(ctx->loss_part.as<float>(), a->n_pos, a->loss_out);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// owner-side selection of the slots a rank must reduce: keys of its row range + every relation key
__global__ void kge_select_flag_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t E, int64_t row_begin, int64_t row_end,
                                       uint64_t* __restrict__ packed, uint8_t* __restrict__ flags) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t key = keys[t];
        packed[t] = ((uint64_t)(uint32_t)key << 32) | (uint64_t)(uint32_t)t;
        flags[t] = (key >= E || (key >= row_begin && key < row_end)) ? 1 : 0;
    }
}

template <int V>
static int launch_apply(const ApplyParams& P, int tmode, cudaStream_t st) {
    const int64_t n_chunks = (P.n_keys + KGE_CH - 1) / KGE_CH;
    dim3 grid((unsigned)((n_chunks + 7) / 8)), block(256);
    if (tmode == 0) kge_reduce_apply_kernel<V, 0><<<grid, block, 0, st>>>(P);
    else if (tmode == 1) kge_reduce_apply_kernel<V, 1><<<grid, block, 0, st>>>(P);
    else kge_reduce_apply_kernel<V, 2><<<grid, block, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    kge_span_apply_kernel<V><<<(unsigned)n_chunks, 128, 0, st>>>(P);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// packed_in: n_items (key << 32 | global slot) entries, unsorted; sorted by key (stable) into ctx->ks_sorted
static int apply_impl(kge_ctx* ctx, const kge_train_args* a, const uint64_t* packed_in, int64_t n_items, const kge_table* grads,
                      int64_t row_begin, int64_t row_end, cudaStream_t st) {
    KGE_REQUIRE(n_items < (int64_t)INT32_MAX, "kge_train_apply: too many slots");
    if (n_items == 0) return 0;
    const int K = a->ent.K;
    const int64_t n_chunks = (n_items + KGE_CH - 1) / KGE_CH;
    if (ctx->ks_sorted.reserve((size_t)n_items * 8)) return -2;
    if (ctx->partial.reserve((size_t)2 * n_chunks * K * sizeof(float))) return -2;
    if (ctx->span_head.reserve((size_t)n_chunks)) return -2;
    int64_t E = a->ent.rows;
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < E + a->R) ++end_bit;
    size_t tmp_bytes = 0;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, packed_in, ctx->ks_sorted.as<uint64_t>(), (int)n_items, 32,
                                                  32 + end_bit, st));
    if (ctx->sort_tmp.reserve(tmp_bytes)) return -2;
    KGE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp_bytes, packed_in, ctx->ks_sorted.as<uint64_t>(), (int)n_items,
                                                  32, 32 + end_bit, st));
    ApplyParams P;
    P.ks = ctx->ks_sorted.as<uint64_t>();
    P.n_keys = n_items;
    for (int i = 0; i < KGE_MAX_SHARDS; ++i) P.G.base[i] = grads->shard[i];
    P.G.n_ranks = grads->n_shards;
    P.G.S = grads->rows_per_shard;
    P.G.n = grads->rows_per_shard / (3 + a->eta);
    P.G.eta = a->eta;
    P.G.K = K;
    P.ent = make_view(a->ent);
    P.ent_m = make_view(a->ent_m);
    P.ent_v = make_view(a->ent_v);
    P.has_m = table_present(a->ent_m);
    P.has_v = table_present(a->ent_v);
    P.rel = a->rel;
    P.rel_m = a->rel_m;
    P.rel_v = a->rel_v;
    P.E = E;
    P.R = a->R;
    P.row_begin = row_begin;
    P.row_end = row_end;
    P.opt = a->opt;
    P.flags = a->flags;
    P.lr = a->lr;
    P.beta1 = a->beta1;
    P.beta2 = a->beta2;
    P.eps = a->eps;
    P.momentum = a->momentum;
    P.partial = ctx->partial.as<float>();
    P.span_head = ctx->span_head.as<uint8_t>();
    P.dbg_grad_ent = a->dbg_grad_ent;
    P.dbg_grad_rel = a->dbg_grad_rel;
    const bool reset = (a->flags & KGE_F_RESET_STATE) != 0;
    const bool no_update = (a->flags & KGE_F_NO_UPDATE) != 0;
    double t = reset ? 1.0 : (double)(a->step < 1 ? 1 : a->step);
    P.lr_t = (float)((double)a->lr * sqrt(1.0 - pow((double)a->beta2, t)) / (1.0 - pow((double)a->beta1, t)));
    if (!no_update && !reset) {
        if (a->opt == KGE_OPT_ADAM)
            KGE_REQUIRE(P.has_m && P.has_v && a->rel_m && a->rel_v, "kge_train: adam state (m,v) missing");
        if (a->opt == KGE_OPT_ADAGRAD || a->opt == KGE_OPT_MOMENTUM)
            KGE_REQUIRE(P.has_m && a->rel_m, "kge_train: optimizer state missing");
    }
    const int tmode = a->model == KGE_TRANSE_L1 ? 1 : (a->model == KGE_TRANSE_L2 ? 2 : 0);
    if (K % 4 == 0) return launch_apply<4>(P, tmode, st);
    return launch_apply<1>(P, tmode, st);
}

// Selection of the slots this rank reduces (keys in [row_begin,row_end) or relation keys), started
// early so that the count reaches the host while kge_train_fwd_bwd is still running.
extern "C" int kge_train_select(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys, int64_t row_begin,
                                int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_select: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(keys_all != nullptr || n_keys == 0, "kge_train_select: keys missing");
    KGE_REQUIRE(n_keys < (int64_t)INT32_MAX, "kge_train_select: too many slots");
    cudaStream_t st = (cudaStream_t)stream;
    ctx->sel_valid = false;
    if (n_keys == 0) return 0;
    if (ctx->ks_in.reserve((size_t)n_keys * 8) || ctx->ks_sel.reserve((size_t)n_keys * 8) || ctx->sel_flags.reserve((size_t)n_keys) ||
        ctx->sel_count.reserve(sizeof(int)))
        return -2;
    if (ctx->h_count == nullptr) {
        KGE_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_count, sizeof(int)));
        KGE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_count, cudaEventDisableTiming));
    }
    int threads = 256;
    int blocks = (int)std::min<int64_t>((n_keys + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
    kge_select_flag_kernel<<<blocks, threads, 0, st>>>(keys_all, n_keys, a->ent.rows, row_begin, row_end, ctx->ks_in.as<uint64_t>(),
                                                       ctx->sel_flags.as<uint8_t>());
    KGE_CUDA_CHECK(cudaGetLastError());
    size_t tb = 0;
    KGE_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, tb, ctx->ks_in.as<uint64_t>(), ctx->sel_flags.as<uint8_t>(),
                                              ctx->ks_sel.as<uint64_t>(), ctx->sel_count.as<int>(), (int)n_keys, st));
    if (ctx->sort_tmp.reserve(tb)) return -2;
    KGE_CUDA_CHECK(cub::DeviceSelect::Flagged(ctx->sort_tmp.p, tb, ctx->ks_in.as<uint64_t>(), ctx->sel_flags.as<uint8_t>(),
                                              ctx->ks_sel.as<uint64_t>(), ctx->sel_count.as<int>(), (int)n_keys, st));
    KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_count, ctx->sel_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaEventRecord(ctx->ev_count, st));
    ctx->sel_valid = true;
    ctx->sel_keys = keys_all;
    ctx->sel_n = n_keys;
    ctx->sel_begin = row_begin;
    ctx->sel_end = row_end;
    return 0;
}

extern "C" int kge_train_apply(kge_ctx* ctx, const kge_train_args* a, const int32_t* keys_all, int64_t n_keys,
                               const kge_table* grads, int64_t row_begin, int64_t row_end, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_apply: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(grads != nullptr && keys_all != nullptr, "kge_train_apply: missing keys/grads");
    if (n_keys == 0) return 0;
    KGE_REQUIRE(grads->n_shards >= 1 && grads->rows_per_shard > 0 && grads->rows_per_shard % (3 + a->eta) == 0,
                "kge_train_apply: grads.rows_per_shard must be the slots per rank, (3+eta)*n_pos");
    KGE_REQUIRE(n_keys == grads->rows_per_shard * grads->n_shards, "kge_train_apply: n_keys != n_shards * slots per rank");
    if (!(ctx->sel_valid && ctx->sel_keys == keys_all && ctx->sel_n == n_keys && ctx->sel_begin == row_begin && ctx->sel_end == row_end)) {
        if (int rc = kge_train_select(ctx, a, keys_all, n_keys, row_begin, row_end, stream)) return rc;
    }
    KGE_CUDA_CHECK(cudaEventSynchronize(ctx->ev_count));
    ctx->sel_valid = false;
    const int64_t m = *ctx->h_count;
    return apply_impl(ctx, a, ctx->ks_sel.as<uint64_t>(), m, grads, row_begin, row_end, (cudaStream_t)stream);
}

extern "C" int kge_train_step(kge_ctx* ctx, const kge_train_args* a, void* stream) {
    KGE_REQUIRE(ctx != nullptr, "kge_train_step: null ctx");
    if (int rc = validate_train(a)) return rc;
    KGE_REQUIRE(a->ent.n_shards == 1, "kge_train_step is the single-GPU entry; use the phased calls when sharded");
    if (a->n_pos == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = a->ent.K;
    int64_t S = (int64_t)(3 + a->eta) * a->n_pos;
    if (ctx->ks_in.reserve((size_t)S * 8)) return -2;
    if (ctx->grad_rows.reserve((size_t)gbuf_floats(a->eta, a->n_pos, K) * sizeof(float))) return -2;
    if (int rc = emit_impl(ctx, a, nullptr, ctx->ks_in.as<uint64_t>(), st)) return rc;
    if (int rc = kge_train_fwd_bwd(ctx, a, ctx->grad_rows.as<float>(), stream)) return rc;
    kge_table g;
    memset(&g, 0, sizeof(g));
    g.shard[0] = ctx->grad_rows.as<float>();
    g.rows = S;
    g.rows_per_shard = S;
    g.n_shards = 1;
    g.K = K;
    return apply_impl(ctx, a, ctx->ks_in.as<uint64_t>(), S, &g, 0, a->ent.rows, st);
}

extern "C" int kge_train_step_host(kge_ctx* ctx, const kge_train_args* a, const int32_t* pos_host, float* loss_host,
                                   void* stream) {
    KGE_REQUIRE(ctx != nullptr && a != nullptr, "kge_train_step_host: null argument");
    KGE_REQUIRE(a->n_pos >= 0 && (a->n_pos == 0 || pos_host != nullptr), "kge_train_step_host: positives missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_pos == 0) {
        if (loss_host) *loss_host = 0.f;
        return 0;
    }
    if (ctx->h_pos.reserve((size_t)a->n_pos * 3 * sizeof(int32_t)) || ctx->h_loss.reserve(sizeof(float))) return -2;
    KGE_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pos.p, pos_host, (size_t)a->n_pos * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    kge_train_args b = *a;
    b.pos = ctx->h_pos.as<int32_t>();
    if (b.loss_out == nullptr) b.loss_out = ctx->h_loss.as<float>();
    if (int rc = kge_train_step(ctx, &b, stream)) return rc;
    if (loss_host) KGE_CUDA_CHECK(cudaMemcpyAsync(loss_host, b.loss_out, sizeof(float), cudaMemcpyDeviceToHost, st));
    KGE_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int kge_normalize_rows(kge_ctx* ctx, float* emb, int64_t rows, int K, void* stream) {
    KGE_REQUIRE(ctx != nullptr && emb != nullptr, "kge_normalize_rows: null argument");
    if (rows == 0) return 0;
    const int warps = 8;
    kge_normalize_rows_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(emb, rows, K);
    KGE_CUDA_CHECK(cudaGetLastError());
    return 0;
}
