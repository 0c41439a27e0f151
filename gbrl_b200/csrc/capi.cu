// capi.cu -- C-ABI (include/gbrl_b200.h): model lifetime, step / fit / predict orchestration, getters.
//
// Mirrors the dispatcher of the reference, class GBRL (gbrl/src/cpp/gbrl.cpp): same validation order and
// error conditions for the calls on the hot path; the device work is delegated to the kernels in this
// directory.  There is no CPU path: every entry point requires a CUDA device.
#include "engine.cuh"
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include <string.h>
#include <algorithm>
#include <numeric>
#include <random>
#include <map>
#include <mutex>

namespace gb {

std::atomic<long long> g_kernel_launches{0};
static thread_local std::string g_last_error;

void dist_unique_id(uint8_t id[128]);
void dist_init(Model &m, const uint8_t id[128], int rank, int world);
void dist_shutdown(Model &m);
double microbench(int which, int iters);
void diag_chain_sums(const float *host_mat, long long ne, int D, int T, int mode, const float *host_mean, float *host_partial,
                     float *host_centered, int impl, double *info);

void ensure_dyn_smem_impl(const void *func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(func, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return;
    GB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done[key] = bytes;
}

int replay_grid_mult(bool side_stream) {
    static const int env = getenv("GBRL_B200_SIDE_GRID") ? atoi(getenv("GBRL_B200_SIDE_GRID")) : 0;
    if (!side_stream) return 8;
    return env > 0 ? env : 8;      // measured r02 (C2 / J3 / C3 / C5): 8, 32 and 128 CTAs per SM give the same iteration time
}

// ---------------------------------------------------------------- DevBuf
void DevBuf::ensure(size_t n, bool keep, cudaStream_t s) {
    if (n <= bytes && p) return;
    size_t nb = n;
    if (nb < 256) nb = 256;
    void *np = nullptr;
    GB_CUDA(cudaMalloc(&np, nb));
    if (keep && p && bytes) {
        GB_CUDA(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, s));
        GB_CUDA(cudaMemsetAsync((char *)np + bytes, 0, nb - bytes, s));
        GB_CUDA(cudaStreamSynchronize(s));
    } else if (keep) {
        GB_CUDA(cudaMemsetAsync(np, 0, nb, s));
    }
    if (p) cudaFree(p);
    p = np; bytes = nb;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
}

// ---------------------------------------------------------------- workspace
static void prepare_workspace(Model &m, int N, int F, cudaStream_t s) {
    Workspace &ws = m.ws;
    const int D = m.cfg.output_dim, md = m.cfg.max_depth, B = m.cfg.n_bins;
    ws.N = N; ws.F = F; ws.D = D; ws.depth = md; ws.B = B;
    ws.nT = ceil_div(F, FT);
    ws.MAXN = (2 << md) - 1;
    // 2-D sharding of the histogram work: G_t tile groups x G_r row groups, G_t * G_r == world.
    // Tile group tg owns a contiguous block of feature tiles; inside a tile group the 8192-row chunks of every
    // node go round-robin to the G_r row groups.  The per-level all-reduce (integer sum) merges both dimensions.
    int gt = m.world < ws.nT ? m.world : ws.nT;
    while (gt > 1 && m.world % gt != 0) --gt;
    const int gr = m.world / gt;
    const int tg = m.rank % gt;
    ws.row_groups = gr; ws.row_group = m.rank / gt;
    ws.tile_lo = (int)((long long)ws.nT * tg / gt);
    ws.tile_hi = (int)((long long)ws.nT * (tg + 1) / gt);
    const size_t n1 = (size_t)(N > 0 ? N : 1);
    ws.order[0].ensure(n1 * sizeof(int)); ws.order[1].ensure(n1 * sizeof(int));
    ws.order_p[0] = ws.order[0].as<int>(); ws.order_p[1] = ws.order[1].as<int>();
    ws.pnode[0].ensure(n1 * sizeof(int)); ws.pnode[1].ensure(n1 * sizeof(int));
    ws.pnode_p[0] = ws.pnode[0].as<int>(); ws.pnode_p[1] = ws.pnode[1].as<int>();
    ws.nid.ensure(n1 * sizeof(int)); ws.rflag.ensure(n1 * sizeof(uint16_t) + 16);
    {   // chunk totals / prefixes of the partition + its "CTAs done" counter (kept zero between launches)
        const int cap = ceil_div((int)n1, 2048) + 1;
        if (cap > ws.chunk_cap || !ws.chunk_sums.p) {
            ws.chunk_cap = cap;
            ws.chunk_sums.ensure((size_t)(cap + 1) * sizeof(int));
            GB_CUDA(cudaMemsetAsync(ws.chunk_sums.p, 0, (size_t)(cap + 1) * sizeof(int), s));
        }
    }
    const int lv = md > 0 ? md - 1 : 0;
    const size_t slot_bytes = (size_t)ws.nT * NB * FT * (1 + D) * sizeof(long long);
    ws.hist[0].ensure(slot_bytes << lv); ws.hist[1].ensure(slot_bytes << lv);
    ws.hist_p[0] = ws.hist[0].as<long long>(); ws.hist_p[1] = ws.hist[1].as<long long>();
    const size_t C = (size_t)F * B;
    ws.scores.ensure((C << lv) * sizeof(float));
    ws.cand_flags.ensure(C << lv);
    ws.obl_tot.ensure(2 * C * sizeof(float));
    size_t tb = (size_t)F << lv;
    if (tb < C / 256 + 1) tb = C / 256 + 1;
    ws.tile_best.ensure(tb * sizeof(float2));
    const int nTl = ws.tile_hi - ws.tile_lo > 0 ? ws.tile_hi - ws.tile_lo : 1;
    ws.items_cap = (N / 1536 + (1 << md) + 1) * nTl;      // smallest item size any histogram variant asks for (hist_item_rows)
    if (!ws.n_sms) {
        cudaDeviceProp prop;
        GB_CUDA(cudaGetDeviceProperties(&prop, m.device));
        ws.n_sms = prop.multiProcessorCount;
    }
    ws.items.ensure((size_t)ws.items_cap * sizeof(Item));
    ws.replay_cap = 1 << 18;
    if ((size_t)ws.replay_cap < 4 * ((size_t)1 << md)) ws.replay_cap = 4 << md;
    ws.rbits_words = (long long)((n1 + 255) / 256) * 8 * 64 + 4096 + 256 * 2048;      // + the 32-group alignment slack of up to 2048 items
    ws.rwide_groups = ws.rbits_words / 8 + 1;
    const size_t rwide_bytes = (size_t)ws.rwide_groups * ((size_t)2 * D * (sizeof(double) + sizeof(int4) + 2 * sizeof(float)) + sizeof(int)) +
                               (size_t)ws.replay_cap * 8 * sizeof(float) + 64 +
                               ((size_t)ws.rwide_groups / 32 + 2) * 2 * D * (sizeof(int4) + sizeof(float));      // + window tables
    auto ensure_replay = [&](DevBuf &replay, DevBuf &scores, DevBuf &rgrad, DevBuf &rbits, DevBuf &rmeta, DevBuf &rwide) {
        replay.ensure((size_t)ws.replay_cap * (sizeof(ReplayItem) + sizeof(int)));
        scores.ensure((size_t)2 * ws.replay_cap * sizeof(float));      // replayed scores, then the exact-tier scores of the same items
        if (m.cfg.tie_replay && D <= 4) {
            // replay streams (split.cu): planes for 64 full-size candidates, i.e. all items of a level unless the near-tie
            // band is unusually crowded (those items are gathered directly by their chain CTA)
            rgrad.ensure(n1 * D * sizeof(float));
            rbits.ensure((size_t)ws.rbits_words * sizeof(unsigned int));
            rmeta.ensure(((size_t)3 * ws.replay_cap + 2) * sizeof(int));
            if (D <= 2) rwide.ensure(rwide_bytes);
        }
    };
    ensure_replay(ws.replay, ws.replay_scores, ws.rgrad, ws.rbits, ws.rmeta, ws.rwide);
    // Speculative levels (tree.cu): every level keeps its own replay buffers, row order and histograms, so that the chains of a
    // level can be walked while the deeper levels are grown, and a level can be returned to.  GBRL_B200_SPEC=0 turns it off.
    {
        static const bool env_off = getenv("GBRL_B200_SPEC") != nullptr && getenv("GBRL_B200_SPEC")[0] == '0';
        size_t extra = 0;
        if (md > 0) {
            const size_t per_slot = (size_t)ws.replay_cap * 20 + n1 * D * 4 + (size_t)ws.rbits_words * 4 + (D <= 2 ? rwide_bytes : 0);
            extra = (size_t)md * per_slot + (slot_bytes << md) + (size_t)(md + 1) * n1 * 8;
        }
        ws.spec = m.cfg.tie_replay && md > 0 && !(m.cfg.replay_variant & 2) && !env_off && extra <= ((size_t)24 << 30);
        if (ws.spec) {
            for (int l = 0; l < md; ++l) {
                ReplaySlot &sl = ws.slots[l];
                ensure_replay(sl.replay, sl.replay_scores, sl.rgrad, sl.rbits, sl.rmeta, sl.rwide);
                sl.ctl_snap.ensure(sizeof(Ctl));
                ws.hist_lv[l].ensure(slot_bytes << l);
                ws.order_lv[l].ensure(n1 * sizeof(int));
                ws.pnode_lv[l].ensure(n1 * sizeof(int));
                if (!sl.stream) {
                    int lo = 0, hi = 0;
                    GB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority: the side work never delays the tree
                    GB_CUDA(cudaStreamCreateWithPriority(&sl.stream, cudaStreamNonBlocking, lo));
                    GB_CUDA(cudaEventCreateWithFlags(&sl.ev_sel, cudaEventDisableTiming));
                    GB_CUDA(cudaEventCreateWithFlags(&sl.ev_dec, cudaEventDisableTiming));
                    GB_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
                }
            }
            ws.order_lv[md].ensure(n1 * sizeof(int));
            ws.pnode_lv[md].ensure(n1 * sizeof(int));
            ws.state_snap.ensure((size_t)ws.MAXN * sizeof(int));
            ws.spec_flag.ensure(sizeof(unsigned int));
            if (!ws.h_spec_flag) GB_CUDA(cudaMallocHost(&ws.h_spec_flag, sizeof(unsigned int)));
        }
    }
    // node arrays carved from one allocation
    const size_t MN = ws.MAXN;
    const size_t bytes = MN * D * sizeof(long long) + MN * 10 * sizeof(int) + MN * 4 * sizeof(float);
    ws.nodes.ensure(bytes);
    char *p = ws.nodes.as<char>();
    ws.na.tot_sum = (long long *)p; p += MN * D * sizeof(long long);
    int **ip[] = {&ws.na.seg_start, &ws.na.seg_len, &ws.na.state, &ws.na.split_f, &ws.na.split_j,
                  &ws.na.direct, &ws.na.rep_begin, &ws.na.rep_count, &ws.na.best_idx, &ws.na.leaf_index};
    for (auto q : ip) { *q = (int *)p; p += MN * sizeof(int); }
    float **fp[] = {&ws.na.split_thr, &ws.na.best_gain, &ws.na.parent_score, &ws.na.band};
    for (auto q : fp) { *q = (float *)p; p += MN * sizeof(float); }
    (void)s;
}

static const float *stage_in(DevBuf &buf, const float *ptr, int is_dev, size_t count, cudaStream_t s) {
    if (is_dev) return ptr;
    buf.ensure((count > 0 ? count : 1) * sizeof(float));
    if (count) GB_CUDA(cudaMemcpyAsync(buf.p, ptr, count * sizeof(float), cudaMemcpyHostToDevice, s));
    return buf.as<float>();
}

static void sync_ctl(Model &m, cudaStream_t s) {
    Ctl h;
    GB_CUDA(cudaMemcpyAsync(&h, m.ws.ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    m.ens.n_leaves = h.n_leaves;
    m.ens.n_leaves_ub = h.n_leaves;
    GB_CHECK(h.n_trees == m.ens.n_trees, "internal error: device/host tree count mismatch");
    m.replay_items = h.stat_replay_items; m.replay_nodes = h.stat_replay_nodes;
    m.nodes_evaluated = h.stat_nodes_evaluated;
    if (h.replay_overflow > m.replay_overflow && m.replay_overflow == 0) {
        // not silent: a node (or level) whose near-tie list did not fit was decided on the exact tier alone, which can differ from
        // the reference where the reference's own rounding noise decides (two tiers)
        fprintf(stderr, "gbrl_b200: warning: the near-tie replay list overflowed (%d node(s)); those decisions use exact-arithmetic "
                        "scores only. See get_stats()['replay_overflow'].\n", h.replay_overflow);
    }
    m.replay_overflow = h.replay_overflow;
    m.hist_rows = h.stat_hist_rows;
    m.chain_fast = h.stat_chain_fast; m.chain_slow = h.stat_chain_slow; m.chain_seq = h.stat_chain_seq; m.replay_flips = h.stat_replay_flips;
    m.max_noise = __builtin_bit_cast(float, h.stat_max_noise);
}

static void check_features(Model &m, int n_features) {
    // gbrl.cpp:946-957: the feature split is latched at iteration 0
    if (m.iteration == 0) { m.n_num_features = n_features; m.n_cat_features = 0; }
    GB_CHECK(n_features == m.n_num_features && m.n_cat_features == 0,
             "Incompatible dataset: feature count differs from the one the ensemble was started with");
    GB_CHECK(n_features == m.cfg.input_dim, "Incompatible dataset: n_num_features + n_cat_features != input_dim");
}

static void do_step(Model &m, const float *obs, int obs_dev, const float *grads, int grads_dev, int N, int F, cudaStream_t s) {
    GB_CHECK(N > 0, "step: n_samples must be positive");
    check_features(m, F);
    GB_CUDA(cudaSetDevice(m.device));
    Workspace &ws = m.ws;
    const int D = m.cfg.output_dim;
    const float *X = stage_in(ws.xstage, obs, obs_dev, (size_t)N * F, s);
    const float *G = stage_in(ws.gstage, grads, grads_dev, (size_t)N * D, s);
    prepare_workspace(m, N, F, s);
    { ProfScope ps(m, P_CAND, s); compute_thresholds(m, X, N, F, s); }   // fitter.cpp:72-90: candidates from the current observations
    ws.codes_rows = N; ws.row_offset = 0;
    { ProfScope ps(m, P_BIN, s); bin_features(m, X, N, F, s); }
    { ProfScope ps(m, P_PRE, s); build_grads(m, G, N, s); }               // fitter.cpp:57-64
    grow_tree(m, X, G, N, F, s);                // fitter.cpp:98-102
    sync_ctl(m, s);
    m.iteration++;                              // fitter.cpp:114
}

__global__ void fill_rows_kernel(float *dst, const float *bias, long long n, int D) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += (long long)gridDim.x * blockDim.x)
        dst[i] = bias[i % D];
}
__global__ void gather_rows_kernel(const float *src, const int *perm, float *dst, int N, int W) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * W; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[(size_t)perm[i / W] * W + i % W];
}

// ---------------------------------------------------------------- fit as a session
// fit == fit_begin + fit_iterate(iterations) + fit_end.  bench.py uses the three pieces to time exactly K boosting
// iterations with the inputs already resident in HBM; gbrl_b200_fit() is the reference-facing call.
static void fit_begin(Model &m, const float *obs, int obs_dev, const float *targets, int targets_dev, int N, int F, int shuffle,
                      cudaStream_t s) {
    GB_CHECK(N > 0, "fit: n_samples must be positive");
    check_features(m, F);
    GB_CHECK(!m.opts.empty(), "fit: no optimizers set");
    GB_CUDA(cudaSetDevice(m.device));
    Workspace &ws = m.ws;
    FitSession &fs = m.fs;
    const int D = m.cfg.output_dim, bs = m.cfg.batch_size;
    GB_CHECK(bs > 0, "fit: batch_size must be positive");
    const float *X = stage_in(ws.xstage, obs, obs_dev, (size_t)N * F, s);
    const float *Tg = stage_in(ws.tstage, targets, targets_dev, (size_t)N * D, s);
    if (shuffle) {
        // gbrl.cpp:1017-1026: the reference shuffles with std::random_device (not reproducible); we do the same
        std::vector<int> perm(N);
        std::iota(perm.begin(), perm.end(), 0);
        // every rank of a multi-GPU job must see the same row order (the histogram items and the replay chains are
        // defined on positions of `order`): rank 0 draws the seed, an integer all-reduce hands it to the others
        unsigned long long seed = 0;
        if (m.rank == 0) { std::random_device rd; seed = ((unsigned long long)rd() << 32) | rd(); }
        if (m.world > 1) {
            DevBuf sb;
            sb.ensure(sizeof(long long));
            long long hs = (long long)(seed >> 1);
            GB_CUDA(cudaMemcpyAsync(sb.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
            dist_allreduce_hist(m, sb.as<long long>(), 1, s);
            GB_CUDA(cudaMemcpyAsync(&hs, sb.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
            GB_CUDA(cudaStreamSynchronize(s));
            seed = (unsigned long long)hs;
        }
        std::mt19937_64 g(seed);
        std::shuffle(perm.begin(), perm.end(), g);
        DevBuf permbuf;
        permbuf.ensure((size_t)N * sizeof(int)); m.fit_x.ensure((size_t)N * F * sizeof(float)); m.fit_t.ensure((size_t)N * D * sizeof(float));
        GB_CUDA(cudaMemcpyAsync(permbuf.p, perm.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, s));
        GB_LAUNCH(gather_rows_kernel, 1184, 256, 0, s, X, permbuf.as<int>(), m.fit_x.as<float>(), N, F);
        GB_LAUNCH(gather_rows_kernel, 1184, 256, 0, s, Tg, permbuf.as<int>(), m.fit_t.as<float>(), N, D);
        GB_CUDA(cudaStreamSynchronize(s));
        X = m.fit_x.as<float>(); Tg = m.fit_t.as<float>();
    }
    prepare_workspace(m, N, F, s);
    // gbrl.cpp:1076-1078: bias := column mean of the targets (reference thread partition emulated)
    { ProfScope ps(m, P_PRE, s); column_mean_ref(m, Tg, N, D, m.bias.as<float>(), s); }
    m.h_bias.resize(D);
    GB_CUDA(cudaMemcpyAsync(m.h_bias.data(), m.bias.p, D * sizeof(float), cudaMemcpyDeviceToHost, s));
    // fitter.cpp:134-151: candidates ONCE on the full data
    { ProfScope ps(m, P_CAND, s); compute_thresholds(m, X, N, F, s); }
    ws.codes_rows = N; ws.row_offset = 0;
    { ProfScope ps(m, P_BIN, s); bin_features(m, X, N, F, s); }
    fs.n_trees0 = m.ens.n_trees;
    fs.incremental = (fs.n_trees0 == 0);
    ws.preds_full.ensure((size_t)N * D * sizeof(float));
    // two gradient buffers like the reference (regular / last batch), zero-initialised once (fitter.cpp:125-130)
    const size_t gsz = (size_t)(bs < N ? bs : N) * D;
    ws.grads_fit.ensure((gsz + (size_t)(N % bs) * D + 2) * sizeof(float));
    GB_CUDA(cudaMemsetAsync(ws.grads_fit.p, 0, ws.grads_fit.bytes, s));
    fs.g_regular = ws.grads_fit.as<float>(); fs.g_last = fs.g_regular + gsz;
    if (fs.incremental) GB_LAUNCH(fill_rows_kernel, 1184, 256, 0, s, ws.preds_full.as<float>(), m.bias.as<float>(), (long long)N, D);
    fs.X = X; fs.T = Tg; fs.N = N; fs.F = F; fs.it = 0;
    fs.batch_start = 0;
    fs.batch_n = fs.batch_start + bs < N ? bs : N - fs.batch_start;
    fs.active = true;
}

static void fit_iterate(Model &m, int iterations, cudaStream_t s) {
    FitSession &fs = m.fs;
    GB_CHECK(fs.active, "fit_iterate without fit_begin");
    GB_CUDA(cudaSetDevice(m.device));
    Workspace &ws = m.ws;
    const int D = m.cfg.output_dim, bs = m.cfg.batch_size, N = fs.N, F = fs.F;
    for (int k = 0; k < iterations; ++k, ++fs.it) {
        const float *bX = fs.X + (size_t)fs.batch_start * F;
        const float *bT = fs.T + (size_t)fs.batch_start * D;
        const bool is_last = fs.batch_start + bs > N;
        float *grads = is_last ? fs.g_last : fs.g_regular;
        const float *preds;
        if (fs.incremental) {
            preds = ws.preds_full.as<float>() + (size_t)fs.batch_start * D;
        } else {
            // fitter.cpp:191: predict_cpu(batch, 0, i): stop index 0 means "all trees"
            ProfScope ps(m, P_PRED, s);
            m.fit_batch_preds.ensure((size_t)fs.batch_n * D * sizeof(float));
            const int stop = (fs.it == 0) ? m.ens.n_trees : std::min(fs.it, m.ens.n_trees);
            launch_predict(m, bX, fs.batch_n, F, 0, stop, m.fit_batch_preds.as<float>(), true, s);
            preds = m.fit_batch_preds.as<float>();
        }
        // the tree is grown on the batch rows: order/nid are batch-relative, codes are addressed with row_offset
        ws.N = fs.batch_n; ws.row_offset = fs.batch_start;
        { ProfScope ps(m, P_PRE, s);
          multirmse_grads(m, preds, bT, grads, fs.batch_n, s);      // fitter.cpp:193-195
          build_grads(m, grads, fs.batch_n, s); }                   // fitter.cpp:203-214
        grow_tree(m, bX, grads, fs.batch_n, F, s);                  // fitter.cpp:220-225
        if (fs.incremental) {
            ProfScope ps(m, P_PRED, s);
            // full batch: the rows still know the node they ended in; mini-batch: walk the new tree for all N rows
            if (fs.batch_n == N) launch_update_preds_from_nodes(m, N, ws.preds_full.as<float>(), s);
            else launch_update_preds_last_tree(m, fs.X, N, F, ws.preds_full.as<float>(), s);
        }
        fs.batch_start += fs.batch_n;                               // fitter.cpp:227-230
        if (fs.batch_start >= N) fs.batch_start = 0;
        fs.batch_n = fs.batch_start + bs < N ? bs : N - fs.batch_start;
        m.iteration++;
    }
    ws.N = N; ws.row_offset = 0;
}

static float fit_end(Model &m, cudaStream_t s) {
    FitSession &fs = m.fs;
    GB_CHECK(fs.active, "fit_end without fit_begin");
    Workspace &ws = m.ws;
    const int D = m.cfg.output_dim, N = fs.N, F = fs.F;
    // fitter.cpp:244-250: full-data loss over trees [0, iterations)
    float loss = INFINITY;
    const float *fp;
    if (fs.incremental) fp = ws.preds_full.as<float>();
    else {
        m.fit_batch_preds.ensure((size_t)N * D * sizeof(float));
        const int stop = (fs.it == 0) ? m.ens.n_trees : std::min(fs.it, m.ens.n_trees);
        launch_predict(m, fs.X, N, F, 0, stop, m.fit_batch_preds.as<float>(), true, s);
        fp = m.fit_batch_preds.as<float>();
    }
    multirmse_loss(m, fp, fs.T, N, &loss, s);
    sync_ctl(m, s);
    fs.active = false;
    return loss;
}

static float do_fit(Model &m, const float *obs, int obs_dev, const float *targets, int targets_dev, int iterations, int N, int F,
                    int shuffle, cudaStream_t s) {
    GB_CHECK(iterations >= 0, "fit: iterations must be >= 0");
    fit_begin(m, obs, obs_dev, targets, targets_dev, N, F, shuffle, s);
    fit_iterate(m, iterations, s);
    return fit_end(m, s);
}

// ---------------------------------------------------------------- profiling
static const char *const PROF_NAMES[P_NCAT] = {"gbrl_b200/candidates", "gbrl_b200/binning", "gbrl_b200/preprocess", "gbrl_b200/histogram",
                                               "gbrl_b200/exchange", "gbrl_b200/scan", "gbrl_b200/select_replay", "gbrl_b200/plan_decide",
                                               "gbrl_b200/partition", "gbrl_b200/finalize", "gbrl_b200/predict", "gbrl_b200/spec_wait"};

ProfScope::ProfScope(Model &m_, int cat_, cudaStream_t s_) : m(m_), cat(cat_), s(s_), on(m_.profile) {
    nvtxRangePushA(PROF_NAMES[cat]);          // one NVTX range per kernel class (shows up in nsys / ncu --nvtx timelines)
    if (!on) return;
    if (m.prof_used + 2 > m.prof_events.size()) {
        for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); m.prof_events.push_back(e); }
    }
    idx = m.prof_used; m.prof_used += 2;
    m.prof_cats.push_back(cat);
    l0 = g_kernel_launches.load();
    cudaEventRecord(m.prof_events[idx], s);
}
ProfScope::~ProfScope() {
    nvtxRangePop();
    if (!on) return;
    cudaEventRecord(m.prof_events[idx + 1], s);
    m.prof_launches.push_back(g_kernel_launches.load() - l0);
}
void prof_collect(Model &m, cudaStream_t s) {
    cudaStreamSynchronize(s);
    for (size_t i = 0; i < m.prof_cats.size(); ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, m.prof_events[2 * i], m.prof_events[2 * i + 1]) == cudaSuccess) {
            m.prof_ms[m.prof_cats[i]] += ms;
            m.prof_n[m.prof_cats[i]] += m.prof_launches[i];
        }
    }
    m.prof_cats.clear(); m.prof_launches.clear(); m.prof_used = 0;
}

static void do_predict(Model &m, const float *obs, int obs_dev, int N, int F, int start_tree, int stop_tree, float *preds,
                       int preds_dev, cudaStream_t s) {
    // gbrl.cpp:369-422 + binding.cpp:800-811
    GB_CHECK(N > 0, "predict: n_samples must be positive");
    if (m.iteration == 0) { m.n_num_features = F; m.n_cat_features = 0; }
    GB_CHECK(F == m.cfg.input_dim, "Incompatible dataset: n_num_features + n_cat_features != input_dim");
    GB_CHECK(F == m.n_num_features, "Incompatible dataset: feature layout differs from the ensemble's");
    GB_CHECK(start_tree >= 0 && stop_tree >= 0, "predict: tree indices must be non-negative");
    GB_CUDA(cudaSetDevice(m.device));
    const int D = m.cfg.output_dim;
    const int n_trees = m.ens.n_trees;
    GB_CHECK(stop_tree <= n_trees, "predict: stop_tree_idx greater than number of trees in model");
    GB_CHECK(start_tree <= n_trees, "predict: start_tree_idx greater than number of trees in model");
    if (stop_tree == 0) stop_tree = n_trees;
    Workspace &ws = m.ws;
    const float *X = stage_in(ws.xstage, obs, obs_dev, (size_t)N * F, s);
    float *out = preds;
    if (!preds_dev) { ws.pstage.ensure((size_t)N * D * sizeof(float)); out = ws.pstage.as<float>(); }
    const bool have_opts = !m.opts.empty();
    // predictor.cpp:122-140: bias is always added; trees only if there are trees and optimizers
    { ProfScope ps(m, P_PRED, s); launch_predict(m, X, N, F, start_tree, (n_trees > 0 && have_opts) ? stop_tree : start_tree, out, true, s); }
    if (!preds_dev) GB_CUDA(cudaMemcpyAsync(preds, out, (size_t)N * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace gb

// ====================================================================================================
using namespace gb;
struct gbrl_b200_model { Model m; };

#define API_BEGIN try {
#define API_END                                   \
    return 0;                                     \
    } catch (const std::exception &e) {           \
        gb::g_last_error = e.what();              \
        return 1;                                 \
    } catch (...) {                               \
        gb::g_last_error = "unknown error";       \
        return 1;                                 \
    }

extern "C" {

const char *gbrl_b200_last_error(void) { return gb::g_last_error.c_str(); }

int gbrl_b200_cuda_available(void) {
    int n = 0;
    return (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
}

int gbrl_b200_create(const gbrl_b200_config *cfg, gbrl_b200_model **out) {
    API_BEGIN
    GB_CHECK(cfg && out, "null argument");
    GB_CHECK(gbrl_b200_cuda_available(), "no CUDA device: this engine has no CPU fallback");
    GB_CHECK(cfg->input_dim > 0 && cfg->output_dim > 0, "input_dim and output_dim must be positive");
    GB_CHECK(cfg->output_dim <= 64, "output_dim > 64 is not supported by this engine yet");
    GB_CHECK(cfg->max_depth >= 0 && cfg->max_depth <= MAX_DEPTH_SUPPORTED, "max_depth must be in [0, 12]");
    GB_CHECK(cfg->n_bins >= 1 && cfg->n_bins <= NB, "n_bins must be in [1, 256]");
    GB_CHECK(cfg->split_score_func == 0 || cfg->split_score_func == 1, "invalid split_score_func");
    GB_CHECK(cfg->generator_type == 0 || cfg->generator_type == 1, "invalid generator_type");
    GB_CHECK(cfg->grow_policy == 0 || cfg->grow_policy == 1, "invalid grow_policy");
    int ndev = 0;
    GB_CUDA(cudaGetDeviceCount(&ndev));
    GB_CHECK(cfg->device_ordinal >= 0 && cfg->device_ordinal < ndev, "invalid device ordinal");
    GB_CUDA(cudaSetDevice(cfg->device_ordinal));
    auto *h = new gbrl_b200_model();
    Model &m = h->m;
    m.cfg = *cfg;
    if (m.cfg.ref_threads < 1) m.cfg.ref_threads = 1;
    if (m.cfg.par_th < 1) m.cfg.par_th = 1;
    m.device = cfg->device_ordinal;
    const int I = cfg->input_dim, D = cfg->output_dim;
    m.bias.ensure(D * sizeof(float), true); m.feature_weights.ensure(I * sizeof(float), true);
    m.rev_num_map.ensure(I * sizeof(int), true);      // zero until set_feature_mapping (types.cpp:232-234)
    m.h_bias.assign(D, 0.0f); m.h_fw.assign(I, 0.0f);
    m.h_mapping.assign(I, 0); m.h_rev_num.assign(I, 0); m.h_rev_cat.assign(I, 0); m.h_numerics.assign(I, 0);
    m.ws.ctl.ensure(sizeof(Ctl), true);
    GB_CUDA(cudaDeviceSynchronize());
    *out = h;
    API_END
}

void gbrl_b200_destroy(gbrl_b200_model *h) {
    if (!h) return;
    try { gb::dist_shutdown(h->m); } catch (...) {}
    cudaSetDevice(h->m.device);
    for (auto &sl : h->m.ws.slots) {
        if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
        if (sl.ev_sel) cudaEventDestroy(sl.ev_sel);
        if (sl.ev_dec) cudaEventDestroy(sl.ev_dec);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (h->m.ws.h_spec_flag) cudaFreeHost(h->m.ws.h_spec_flag);
    delete h;
}

int gbrl_b200_set_bias(gbrl_b200_model *h, const float *bias, int n, int dev) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(n == m.cfg.output_dim, "Incompatible dimensions: bias");
    GB_CUDA(cudaSetDevice(m.device));
    GB_CUDA(cudaMemcpy(m.bias.p, bias, n * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemcpy(m.h_bias.data(), m.bias.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    API_END
}

int gbrl_b200_set_feature_weights(gbrl_b200_model *h, const float *w, int n, int dev) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(n == m.cfg.input_dim, "Incompatible dimensions: feature_weights");
    GB_CUDA(cudaSetDevice(m.device));
    GB_CUDA(cudaMemcpy(m.feature_weights.p, w, n * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemcpy(m.h_fw.data(), m.feature_weights.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    API_END
}

int gbrl_b200_set_feature_mapping(gbrl_b200_model *h, const int *mapping, const uint8_t *numerics, int n) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(n == m.cfg.input_dim, "Incompatible dimensions: feature_mapping");
    // gbrl.cpp:280-293
    int j = 0, k = 0;
    for (int i = 0; i < n; ++i) { m.h_rev_num[i] = -1; m.h_rev_cat[i] = -1; }
    for (int i = 0; i < n; ++i) {
        m.h_mapping[i] = mapping[i]; m.h_numerics[i] = numerics[i] ? 1 : 0;
        if (numerics[i]) m.h_rev_num[j++] = i; else m.h_rev_cat[k++] = i;
    }
    GB_CUDA(cudaSetDevice(m.device));
    GB_CUDA(cudaMemcpy(m.rev_num_map.p, m.h_rev_num.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    API_END
}

int gbrl_b200_get_bias(gbrl_b200_model *h, float *out) {
    API_BEGIN
    memcpy(out, h->m.h_bias.data(), h->m.cfg.output_dim * sizeof(float));
    API_END
}
int gbrl_b200_get_feature_weights(gbrl_b200_model *h, float *out) {
    API_BEGIN
    memcpy(out, h->m.h_fw.data(), h->m.cfg.input_dim * sizeof(float));
    API_END
}
int gbrl_b200_get_feature_mapping(gbrl_b200_model *h, int *mapping, uint8_t *numerics, int *rev_num, int *rev_cat) {
    API_BEGIN
    Model &m = h->m;
    const int n = m.cfg.input_dim;
    if (mapping) memcpy(mapping, m.h_mapping.data(), n * sizeof(int));
    if (numerics) memcpy(numerics, m.h_numerics.data(), n);
    if (rev_num) memcpy(rev_num, m.h_rev_num.data(), n * sizeof(int));
    if (rev_cat) memcpy(rev_cat, m.h_rev_cat.data(), n * sizeof(int));
    API_END
}

int gbrl_b200_set_optimizer(gbrl_b200_model *h, int scheduler, float init_lr, int start_idx, int stop_idx, float stop_lr, int T) {
    API_BEGIN
    Model &m = h->m;
    // gbrl.cpp:457-471
    GB_CHECK((int)m.opts.size() < m.cfg.output_dim && (int)m.opts.size() < MAX_OPTS, "Optimizer Limit Reached");
    GB_CHECK(start_idx < stop_idx, "invalid index ranges");
    GB_CHECK(!(start_idx < 0 || stop_idx <= 0 || start_idx >= m.cfg.output_dim || stop_idx > m.cfg.output_dim), "invalid index ranges");
    GB_CHECK(scheduler == GBRL_B200_SCHED_CONST || scheduler == GBRL_B200_SCHED_LINEAR, "Unrecognized scheduler func");
    Optimizer o; o.sched = scheduler; o.init_lr = init_lr; o.start_idx = start_idx; o.stop_idx = stop_idx; o.stop_lr = stop_lr; o.T = T;
    m.opts.push_back(o);
    GB_CUDA(cudaSetDevice(m.device));
    upload_optimizers(m, 0);
    API_END
}
int gbrl_b200_n_optimizers(gbrl_b200_model *h) { return (int)h->m.opts.size(); }
int gbrl_b200_get_optimizer(gbrl_b200_model *h, int i, int *scheduler, float *init_lr, int *start_idx, int *stop_idx, float *stop_lr, int *T) {
    API_BEGIN
    GB_CHECK(i >= 0 && i < (int)h->m.opts.size(), "optimizer index out of range");
    const Optimizer &o = h->m.opts[i];
    *scheduler = o.sched; *init_lr = o.init_lr; *start_idx = o.start_idx; *stop_idx = o.stop_idx; *stop_lr = o.stop_lr; *T = o.T;
    API_END
}
int gbrl_b200_get_scheduler_lrs(gbrl_b200_model *h, float *out) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(!m.opts.empty(), "No optimizers found");
    const int t = m.ens.n_trees;
    for (size_t i = 0; i < m.opts.size(); ++i) {
        const Optimizer &o = m.opts[i];
        float lr = o.init_lr;
        if (o.sched == GBRL_B200_SCHED_LINEAR) {
            float T_ = (float)o.T, t_ = (float)t + 1;
            float pr = (T_ - t_) / T_;
            lr = o.init_lr + (1.0f - pr) * (o.stop_lr - o.init_lr);
            if (lr < o.stop_lr) lr = o.stop_lr;
        }
        out[i] = lr;
    }
    API_END
}

int gbrl_b200_step(gbrl_b200_model *h, const float *obs, int obs_dev, const float *grads, int grads_dev, int n_samples,
                   int n_features, void *stream) {
    API_BEGIN
    GB_CHECK(h && obs && grads, "null argument");
    gb::do_step(h->m, obs, obs_dev, grads, grads_dev, n_samples, n_features, (cudaStream_t)stream);
    API_END
}

int gbrl_b200_fit(gbrl_b200_model *h, const float *obs, int obs_dev, const float *targets, int targets_dev, int iterations,
                  int n_samples, int n_features, int shuffle, float *loss_out, void *stream) {
    API_BEGIN
    GB_CHECK(h && obs && targets, "null argument");
    const float l = gb::do_fit(h->m, obs, obs_dev, targets, targets_dev, iterations, n_samples, n_features, shuffle, (cudaStream_t)stream);
    if (loss_out) *loss_out = l;
    API_END
}

int gbrl_b200_fit_begin(gbrl_b200_model *h, const float *obs, int obs_dev, const float *targets, int targets_dev, int n_samples,
                        int n_features, int shuffle, void *stream) {
    API_BEGIN
    GB_CHECK(h && obs && targets, "null argument");
    gb::fit_begin(h->m, obs, obs_dev, targets, targets_dev, n_samples, n_features, shuffle, (cudaStream_t)stream);
    API_END
}
int gbrl_b200_fit_iterate(gbrl_b200_model *h, int iterations, int sync, void *stream) {
    API_BEGIN
    gb::fit_iterate(h->m, iterations, (cudaStream_t)stream);
    if (sync) GB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    API_END
}
int gbrl_b200_fit_end(gbrl_b200_model *h, float *loss_out, void *stream) {
    API_BEGIN
    const float l = gb::fit_end(h->m, (cudaStream_t)stream);
    if (loss_out) *loss_out = l;
    API_END
}

int gbrl_b200_profile(gbrl_b200_model *h, int enable) {
    API_BEGIN
    Model &m = h->m;
    if (!enable && m.profile) gb::prof_collect(m, 0);
    m.profile = enable != 0;
    if (enable) { for (int i = 0; i < gb::P_NCAT; ++i) { m.prof_ms[i] = 0; m.prof_n[i] = 0; } }
    API_END
}
int gbrl_b200_get_profile(gbrl_b200_model *h, double *ms, long long *launches, int n, long long *hist_rows) {
    API_BEGIN
    Model &m = h->m;
    GB_CUDA(cudaSetDevice(m.device));
    gb::prof_collect(m, 0);
    for (int i = 0; i < n && i < gb::P_NCAT; ++i) { if (ms) ms[i] = m.prof_ms[i]; if (launches) launches[i] = m.prof_n[i]; }
    if (hist_rows) {
        gb::Ctl c;
        GB_CUDA(cudaMemcpy(&c, m.ws.ctl.p, sizeof(gb::Ctl), cudaMemcpyDeviceToHost));
        *hist_rows = c.stat_hist_rows;
    }
    API_END
}

int gbrl_b200_predict(gbrl_b200_model *h, const float *obs, int obs_dev, int n_samples, int n_features, int start_tree_idx,
                      int stop_tree_idx, float *preds, int preds_dev, void *stream) {
    API_BEGIN
    GB_CHECK(h && obs && preds, "null argument");
    gb::do_predict(h->m, obs, obs_dev, n_samples, n_features, start_tree_idx, stop_tree_idx, preds, preds_dev, (cudaStream_t)stream);
    API_END
}

int gbrl_b200_get_metadata(gbrl_b200_model *h, gbrl_b200_metadata *o) {
    API_BEGIN
    Model &m = h->m;
    memset(o, 0, sizeof(*o));
    o->input_dim = m.cfg.input_dim; o->output_dim = m.cfg.output_dim; o->policy_dim = m.cfg.policy_dim;
    o->max_depth = m.cfg.max_depth; o->min_data_in_leaf = m.cfg.min_data_in_leaf; o->n_bins = m.cfg.n_bins;
    o->par_th = m.cfg.par_th; o->batch_size = m.cfg.batch_size; o->split_score_func = m.cfg.split_score_func;
    o->generator_type = m.cfg.generator_type; o->grow_policy = m.cfg.grow_policy; o->verbose = m.cfg.verbose;
    o->n_num_features = m.n_num_features; o->n_cat_features = m.n_cat_features; o->n_trees = m.ens.n_trees;
    o->n_leaves = m.ens.n_leaves; o->iteration = m.iteration;
    o->kernel_launches = gb::g_kernel_launches.load(); o->replay_items = m.replay_items; o->replay_nodes = m.replay_nodes;
    o->replay_overflow = m.replay_overflow; o->nodes_evaluated = m.nodes_evaluated; o->max_noise_ratio = m.max_noise;
    o->chain_blocks_fast = m.chain_fast; o->chain_blocks_slow = m.chain_slow; o->chain_lanes_seq = m.chain_seq; o->replay_flips = m.replay_flips;
    o->spec_trees = m.spec_trees; o->spec_rollbacks = m.spec_rollbacks;
    API_END
}

int gbrl_b200_get_ensemble(gbrl_b200_model *h, int *tree_indices, int *depths, float *values, int *feature_indices,
                           float *feature_values, float *edge_weights, uint8_t *inequality_directions) {
    API_BEGIN
    Model &m = h->m;
    Ensemble &e = m.ens;
    GB_CUDA(cudaSetDevice(m.device));
    const int md = m.cfg.max_depth, D = m.cfg.output_dim;
    const size_t S = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS ? e.n_trees : e.n_leaves;
    if (e.n_trees == 0) return 0;
    GB_CUDA(cudaMemcpy(tree_indices, e.tree_indices.p, e.n_trees * sizeof(int), cudaMemcpyDeviceToHost));
    GB_CUDA(cudaMemcpy(depths, e.depths.p, S * sizeof(int), cudaMemcpyDeviceToHost));
    GB_CUDA(cudaMemcpy(values, e.values.p, (size_t)e.n_leaves * D * sizeof(float), cudaMemcpyDeviceToHost));
    if (md > 0) {
        GB_CUDA(cudaMemcpy(feature_indices, e.feature_indices.p, S * md * sizeof(int), cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(feature_values, e.feature_values.p, S * md * sizeof(float), cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(edge_weights, e.edge_weights.p, (size_t)e.n_leaves * md * sizeof(float), cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(inequality_directions, e.ineq.p, (size_t)e.n_leaves * md, cudaMemcpyDeviceToHost));
    }
    API_END
}

__global__ void set_ctl_counts_kernel(Ctl *ctl, int n_trees, int n_leaves) { ctl->n_trees = n_trees; ctl->n_leaves = n_leaves; }

int gbrl_b200_set_ensemble(gbrl_b200_model *h, int n_trees, int n_leaves, const int *tree_indices, const int *depths,
                           const float *values, const int *feature_indices, const float *feature_values,
                           const float *edge_weights, const uint8_t *inequality_directions, int n_num_features) {
    API_BEGIN
    Model &m = h->m;
    Ensemble &e = m.ens;
    GB_CUDA(cudaSetDevice(m.device));
    GB_CHECK(n_trees >= 0 && n_leaves >= 0, "negative sizes");
    const int md = m.cfg.max_depth, D = m.cfg.output_dim;
    e.n_trees = 0; e.n_leaves = 0;
    // make room (capacity is computed from tree counts; leaves may exceed trees * 2^md only if inconsistent)
    GB_CHECK((long long)n_leaves <= (long long)n_trees * (1 << md), "inconsistent ensemble: too many leaves");
    // a truncated or corrupt .gbrl_model must not turn into out-of-bounds device accesses in predict / rebuild_heap
    const bool obl_ = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    if (n_trees > 0) {
        GB_CHECK(tree_indices && depths && values, "inconsistent ensemble: missing arrays");
        GB_CHECK(md == 0 || (feature_indices && feature_values && edge_weights && inequality_directions),
                 "inconsistent ensemble: missing split arrays");
        GB_CHECK(n_num_features == m.cfg.input_dim, "inconsistent ensemble: feature count differs from input_dim");
        GB_CHECK(tree_indices[0] == 0, "inconsistent ensemble: tree_indices[0] != 0");
        for (int t = 0; t < n_trees; ++t) {
            const int l0 = tree_indices[t], l1 = t + 1 < n_trees ? tree_indices[t + 1] : n_leaves;
            GB_CHECK(l0 >= 0 && l0 < l1 && l1 <= n_leaves, "inconsistent ensemble: tree_indices not increasing / out of range");
            GB_CHECK(l1 - l0 <= (1 << md), "inconsistent ensemble: a tree has more than 2^max_depth leaves");
            if (obl_) GB_CHECK(depths[t] >= 0 && depths[t] <= md && (l1 - l0) == (1 << depths[t]), "inconsistent ensemble: oblivious depth / leaf count");
        }
        const size_t S_ = obl_ ? (size_t)n_trees : (size_t)n_leaves;
        for (size_t r = 0; r < S_; ++r) {
            const int dep = depths[r];
            GB_CHECK(dep >= 0 && dep <= md, "inconsistent ensemble: depth out of range");
            for (int k = 0; k < dep; ++k) {
                const int f = feature_indices[r * md + k];
                GB_CHECK(f >= 0 && f < m.cfg.input_dim, "inconsistent ensemble: feature index out of range");
            }
        }
    }
    ensure_ensemble_capacity(m, n_trees > 0 ? n_trees : 1, 0);
    const size_t S = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS ? n_trees : n_leaves;
    if (n_trees > 0) {
        GB_CUDA(cudaMemcpy(e.tree_indices.p, tree_indices, n_trees * sizeof(int), cudaMemcpyHostToDevice));
        GB_CUDA(cudaMemcpy(e.depths.p, depths, S * sizeof(int), cudaMemcpyHostToDevice));
        GB_CUDA(cudaMemcpy(e.values.p, values, (size_t)n_leaves * D * sizeof(float), cudaMemcpyHostToDevice));
        if (md > 0) {
            GB_CUDA(cudaMemcpy(e.feature_indices.p, feature_indices, S * md * sizeof(int), cudaMemcpyHostToDevice));
            GB_CUDA(cudaMemcpy(e.feature_values.p, feature_values, S * md * sizeof(float), cudaMemcpyHostToDevice));
            GB_CUDA(cudaMemcpy(e.edge_weights.p, edge_weights, (size_t)n_leaves * md * sizeof(float), cudaMemcpyHostToDevice));
            GB_CUDA(cudaMemcpy(e.ineq.p, inequality_directions, (size_t)n_leaves * md, cudaMemcpyHostToDevice));
        }
    }
    e.n_trees = n_trees; e.n_leaves = n_leaves; e.n_leaves_ub = n_leaves;
    GB_LAUNCH(set_ctl_counts_kernel, 1, 1, 0, 0, m.ws.ctl.as<Ctl>(), n_trees, n_leaves);
    rebuild_heap_topology(m, 0);
    GB_CUDA(cudaDeviceSynchronize());
    m.n_num_features = n_num_features; m.n_cat_features = 0;
    if (n_trees > 0 && m.iteration == 0) m.iteration = n_trees;
    API_END
}

int gbrl_b200_set_iteration(gbrl_b200_model *h, int iteration) {
    API_BEGIN
    GB_CHECK(iteration >= 0, "negative iteration");
    h->m.iteration = iteration;
    API_END
}

int gbrl_b200_get_candidates(gbrl_b200_model *h, float *thresholds, int *n_candidates) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(m.have_candidates, "no candidates computed yet");
    GB_CUDA(cudaSetDevice(m.device));
    const int C = m.ws.F * m.ws.B;
    if (thresholds) GB_CUDA(cudaMemcpy(thresholds, m.ws.thr.p, (size_t)C * sizeof(float), cudaMemcpyDeviceToHost));
    if (n_candidates) *n_candidates = C;
    API_END
}

int gbrl_b200_get_root_scores(gbrl_b200_model *h, float *scores, int *n_candidates) {
    API_BEGIN
    Model &m = h->m;
    GB_CHECK(m.have_candidates, "no tree grown yet");
    GB_CUDA(cudaSetDevice(m.device));
    const int C = m.ws.F * m.ws.B;
    if (scores) GB_CUDA(cudaMemcpy(scores, m.ws.scores.p, (size_t)C * sizeof(float), cudaMemcpyDeviceToHost));
    if (n_candidates) *n_candidates = C;
    API_END
}

int gbrl_b200_dist_unique_id(uint8_t id[128]) {
    API_BEGIN
    gb::dist_unique_id(id);
    API_END
}
int gbrl_b200_dist_init(gbrl_b200_model *h, const uint8_t id[128], int rank, int world_size) {
    API_BEGIN
    GB_CHECK(world_size >= 1 && rank >= 0 && rank < world_size, "invalid rank / world size");
    GB_CUDA(cudaSetDevice(h->m.device));
    gb::dist_init(h->m, id, rank, world_size);
    API_END
}
int gbrl_b200_dist_shutdown(gbrl_b200_model *h) {
    API_BEGIN
    gb::dist_shutdown(h->m);
    API_END
}

int gbrl_b200_microbench(int which, int iters, double *result) {
    API_BEGIN
    *result = gb::microbench(which, iters);
    API_END
}

int gbrl_b200_diag_chain_sums(const float *mat, long long n_elements, int D, int T, int mode, const float *mean, float *partial,
                              float *centered, int impl, double *info) {
    API_BEGIN
    GB_CHECK(D >= 1 && T >= 1 && n_elements >= 0, "diag_chain_sums: bad arguments");
    gb::diag_chain_sums(mat, n_elements, D, T, mode, mean, partial, centered, impl, info);
    API_END
}

}  // extern "C"
