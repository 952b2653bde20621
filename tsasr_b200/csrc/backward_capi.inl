// backward_capi.inl -- tsasr_joint_bwd: chunked backward of the fused joint (included by capi.cu).
//
// Per chunk of cell tiles (a bounded workspace of operand images, aligned to whole (utterance, frame-tile) groups):
//   0. tile_activity_kernel + compact_active_kernel  (tile pruning on) ordered list of the tiles that matter
//   1. joint_gemm_kernel<MODE_GRAD>  recompute logits, emit bf16 dlogits + J operand images
//   2. dj_gemm_kernel                dJ^T = W^T * dlogits^T, act', in-register broadcast sums -> partial rows
//   3. reduce_dpre_kernel            fold partial rows into d_enc / d_dec (deterministic)
//   4. dw_gemm_kernel                dW/db split-K partials (+= across chunks, plain RMW, deterministic)
// and a final reduce_dw_kernel over the split-K partials.  All launches are chained with programmatic dependent launch.

namespace {

static constexpr size_t kDefaultChunkBytes = (size_t)3 << 30;  // 3 GiB of operand images per chunk

struct BwdPlan {
    int chunk_tiles;      // multiple of nTu
    int n_chunks;
    int NV2, NHU, NVB, NT4;
    int hu_blk[3], hu_splits[3], hu_item0[3], max_splits, dw_pairs;
    size_t dY_bytes, J_bytes, part_bytes, dW_bytes, db_bytes, prune_bytes, total;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// development: TSASR_DEBUG_TIMING=1 prints the per-kernel milliseconds of every backward call to stderr
// (synchronises; never on in production).  Uses the same event brackets as tsasr_kernel_timing_enable().
struct KernelTimer {
    bool on, was_on;
    int first;
    KernelTimer() {
        static const bool enabled = getenv("TSASR_DEBUG_TIMING") != nullptr;
        on = enabled;
        if (!on) return;
        std::lock_guard<std::mutex> lock(g_timed_mu);
        was_on = g_timing_on.load();
        first = g_n_timed;
        g_timing_on.store(true);
    }
    ~KernelTimer() {
        if (!on) return;
        std::lock_guard<std::mutex> lock(g_timed_mu);
        float total = 0.f;
        for (int i = first; i < g_n_timed; ++i) {
            float ms = 0.f;
            cudaEventSynchronize(g_timed[i].e1);
            cudaEventElapsedTime(&ms, g_timed[i].e0, g_timed[i].e1);
            total += ms;
            fprintf(stderr, "[tsasr timing] %-28s %8.3f ms\n", g_timed[i].name, ms);
        }
        fprintf(stderr, "[tsasr timing] %-28s %8.3f ms\n", "backward kernels total", total);
        if (!was_on) {  // nobody will collect these events
            for (int i = first; i < g_n_timed; ++i) { cudaEventDestroy(g_timed[i].e0); cudaEventDestroy(g_timed[i].e1); }
            g_n_timed = first;
            g_timing_on.store(false);
        }
    }
};

static void plan_bwd(const JointParams& jp, int num_sms, long long max_chunk_cells, BwdPlan* pl) {
    const int tT = 1 << jp.tT_log2, tU = 128 >> jp.tT_log2;
    const int total_tiles = jp.B * jp.nTt * jp.nTu;
    pl->NVB = (jp.V + 63) / 64;
    pl->NT4 = jp.NT * 4;
    // Default chunk: as many tiles as fit kDefaultChunkBytes of operand images (one chunk for
    // BASELINE config 2).  The images are written once and read back while still warm in the 126 MB
    // L2 by CTAs that work on the same tile at the same time; what spills goes to HBM at ~1x volume.
    const size_t tile_bytes = (size_t)(pl->NT4 + jp.KB) * kImgBytes + (size_t)(tT + tU) * jp.H * 4;
    long long want_tiles = max_chunk_cells > 0 ? max_chunk_cells / 128 : (long long)(kDefaultChunkBytes / tile_bytes);
    if (want_tiles < 1) want_tiles = 1;
    if (want_tiles > total_tiles) want_tiles = total_tiles;  // before the narrowing below: max_chunk_cells is a 64-bit count
    int groups = (int)(want_tiles / jp.nTu);
    if (groups < 1) groups = 1;
    pl->chunk_tiles = groups * jp.nTu;
    if (pl->chunk_tiles > total_tiles) pl->chunk_tiles = total_tiles;
    pl->n_chunks = (total_tiles + pl->chunk_tiles - 1) / pl->chunk_tiles;
    // balance the chunks (same count, equal sizes up to one group)
    {
        const int total_groups = total_tiles / jp.nTu;
        const int gpc = (total_groups + pl->n_chunks - 1) / pl->n_chunks;
        pl->chunk_tiles = gpc * jp.nTu;
    }
    // dW work decomposition: (256 v-rows) x (h-unit) x (split-K), one CTA pair each
    pl->NV2 = (jp.V + 255) / 256;
    {
        const int KBe = (jp.KB + 1) & ~1;
        pl->hu_blk[0] = 0;
        if (KBe <= 6) { pl->NHU = 1; pl->hu_blk[1] = jp.KB; pl->hu_blk[2] = jp.KB; }
        else { pl->NHU = 2; pl->hu_blk[1] = KBe - 2 < 8 ? KBe - 2 : 8; pl->hu_blk[2] = jp.KB; }
        // Schedule: work items (h-unit, v-tile, split) are dealt round-robin to the CTA pairs, h-unit 0 first.
        // Item cost ~ measured cycles per pipeline stage of the unit (config 2, L2-bound: ~580 + 157 per J
        // half-image slot) / split factor.  For k = 1, 2, ... items per pair, pick split factors that fill k *
        // pairs slots with near-equal items and keep the k with the smallest simulated makespan; a larger k must
        // promise >= 10 % (measured: at config 2 the model predicted 4 % for k = 2 and it ran 8 % slower; at V = 5000
        // k = 3 runs 18 % faster than k = 1).
        const int pairs = num_sms / 2;
        int w[2] = {0, 0};
        for (int u = 0; u < pl->NHU; ++u) {
            const int nblk_e = (pl->hu_blk[u + 1] - pl->hu_blk[u] + 1) & ~1;
            w[u] = 580 + 157 * (nblk_e / 2);
        }
        if (const char* env = getenv("TSASR_DEBUG_DW_WEIGHTS")) sscanf(env, "%d,%d", &w[0], &w[1]);  // development knob
        double best_span = 1e30;
        int best_sp[2] = {1, 1};
        int k_max = 4;
        if (const char* env = getenv("TSASR_DEBUG_DW_K")) k_max = atoi(env) >= 1 ? atoi(env) : 4;  // development knob
        const double drain = 3000.0 / (2.0 * pl->chunk_tiles);  // one accumulator drain, in stage-cost units per tile
        for (int k = 1; k <= k_max; ++k) {
            const int slots = k * pairs;
            int sp[2] = {1, 1};
            if (pl->NV2 * pl->NHU > slots) continue;
            for (;;) {  // grow the split factor of the unit type with the most expensive items while slots remain
                int used = 0, worst = 0;
                for (int u = 0; u < pl->NHU; ++u) {
                    used += sp[u] * pl->NV2;
                    if ((long long)w[u] * sp[worst] > (long long)w[worst] * sp[u]) worst = u;
                }
                if (used + pl->NV2 > slots || sp[worst] >= pl->chunk_tiles) break;
                ++sp[worst];
            }
            // simulate the round-robin deal
            double span = 0;
            const int n0 = sp[0] * pl->NV2, n_items = n0 + (pl->NHU > 1 ? sp[1] * pl->NV2 : 0);
            for (int c = 0; c < pairs && c < n_items; ++c) {
                double t = 0;
                for (int it = c; it < n_items; it += pairs) t += it < n0 ? (double)w[0] / sp[0] : (double)w[1] / sp[1];
                if (t > span) span = t;
            }
            span += drain * (k - 1);
            if (span < best_span * 0.90) { best_span = span; best_sp[0] = sp[0]; best_sp[1] = sp[1]; }
        }
        pl->max_splits = 1;
        pl->hu_item0[0] = 0;
        pl->hu_splits[2] = 0;
        for (int u = 0; u < 2; ++u) {
            pl->hu_splits[u] = u < pl->NHU ? best_sp[u] : 0;
            pl->hu_item0[u + 1] = pl->hu_item0[u] + pl->hu_splits[u] * pl->NV2;
            if (pl->hu_splits[u] > pl->max_splits) pl->max_splits = pl->hu_splits[u];
        }
        pl->dw_pairs = pl->hu_item0[2] < pairs ? pl->hu_item0[2] : pairs;
    }
    pl->dY_bytes = align_up((size_t)pl->chunk_tiles * pl->NT4 * kImgBytes, 1024);
    pl->J_bytes = align_up((size_t)pl->chunk_tiles * jp.KB * kImgBytes, 1024);
    pl->part_bytes = align_up((size_t)pl->chunk_tiles * (tT + tU) * jp.H * 4, 1024);
    pl->dW_bytes = align_up((size_t)pl->max_splits * pl->NV2 * 256 * jp.H * 4, 1024);
    pl->db_bytes = align_up((size_t)pl->max_splits * pl->NV2 * 256 * 4, 1024);
    // tile pruning: stats (64 B) | flags (1 B per chunk tile) | ids (4 B per chunk tile)
    pl->prune_bytes = align_up(64 + align_up((size_t)pl->chunk_tiles, 16) + 4 * (size_t)pl->chunk_tiles, 1024);
    pl->total = pl->dY_bytes + pl->J_bytes + pl->part_bytes + pl->dW_bytes + pl->db_bytes + pl->prune_bytes + 1024;
}

static int fake_params_for_plan(JointParams& p, int B, int T, int U, int H, int V) {
    if (H % 64 != 0 || H < 64 || H > 64 * kMaxKB || V < 2 || B < 1 || T < 1 || U < 1) return -1;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.U = U; p.H = H; p.V = V;
    choose_tile(T, U, &p.tT_log2);
    const int tT = 1 << p.tT_log2, tU = 128 >> p.tT_log2;
    p.nTt = (T + tT - 1) / tT;
    p.nTu = (U + tU - 1) / tU;
    p.KB = H / 64;
    p.NT = (V + kTileN - 1) / kTileN;
    return 0;
}

}  // namespace

extern "C" {

size_t tsasr_joint_bwd_workspace_bytes(int B, int T, int U, int H, int V, long long max_chunk_cells) {
    JointParams p;
    if (fake_params_for_plan(p, B, T, U, H, V) != 0) return 0;
    int sms = 148, max_smem = 0;
    if (device_info(&sms, &max_smem) != TSASR_OK) sms = 148;  // sizing only; the launch re-checks the device
    BwdPlan pl;
    plan_bwd(p, sms, max_chunk_cells, &pl);
    return pl.total;
}

size_t tsasr_joint_bwd_stats_offset(int B, int T, int U, int H, int V, long long max_chunk_cells) {
    JointParams p;
    if (fake_params_for_plan(p, B, T, U, H, V) != 0) return 0;
    int sms = 148, max_smem = 0;
    if (device_info(&sms, &max_smem) != TSASR_OK) sms = 148;
    BwdPlan pl;
    plan_bwd(p, sms, max_chunk_cells, &pl);
    return pl.dY_bytes + pl.J_bytes + pl.part_bytes + pl.dW_bytes + pl.db_bytes;  // relative to the 1024-aligned workspace base
}

int tsasr_joint_bwd(const void* enc, const void* dec, const void* W, const float* bias, const int32_t* targets,
                    const int32_t* logit_lengths, const int32_t* target_lengths, int B, int T, int U, int H, int V,
                    int blank, int act_kind, float act_param, const float* lat2, const float* logz,
                    const float* alpha, const float* beta, const float* cost, const float* dcost, void* workspace,
                    size_t workspace_bytes, long long max_chunk_cells, float prune_log2_eps, float clamp, float* d_enc,
                    float* d_dec, float* dW, float* db, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_joint_bwd");
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(enc && dec && W && bias && logit_lengths && target_lengths && lat2 && logz && alpha && beta && cost &&
                workspace && d_enc && d_dec && dW && db, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    JointParams jp;
    if (int rc = fill_joint_params(jp, enc, dec, bias, targets, logit_lengths, target_lengths, B, T, U, H, V, blank,
                                   act_kind, act_param, max_smem))
        return rc;
    BwdPlan pl;
    plan_bwd(jp, sms, max_chunk_cells, &pl);
    uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
    if ((size_t)(ws - static_cast<uint8_t*>(workspace)) + pl.total - 1024 > workspace_bytes)
        return fail(TSASR_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);

    jp.lat2_in = reinterpret_cast<const float2*>(lat2);
    jp.logz_in = logz;
    jp.alpha = alpha;
    jp.beta = beta;
    jp.cost = cost;
    jp.dcost = dcost;
    jp.clamp = clamp;
    jp.dY_img = reinterpret_cast<__nv_bfloat16*>(ws);
    jp.J_img = reinterpret_cast<__nv_bfloat16*>(ws + pl.dY_bytes);

    BwdParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.logit_lengths = logit_lengths;
    bp.target_lengths = target_lengths;
    bp.B = B; bp.T = T; bp.U = U; bp.H = H; bp.V = V;
    bp.act_kind = act_kind;
    bp.act_param = act_param;
    bp.tT_log2 = jp.tT_log2; bp.nTt = jp.nTt; bp.nTu = jp.nTu;
    bp.KB = jp.KB; bp.NVB = pl.NVB; bp.NT4 = pl.NT4;
    bp.dY_img = jp.dY_img;
    bp.J_img = jp.J_img;
    bp.dpre_part = reinterpret_cast<float*>(ws + pl.dY_bytes + pl.J_bytes);
    bp.dW_part = reinterpret_cast<float*>(ws + pl.dY_bytes + pl.J_bytes + pl.part_bytes);
    bp.db_part = reinterpret_cast<float*>(ws + pl.dY_bytes + pl.J_bytes + pl.part_bytes + pl.dW_bytes);
    uint8_t* prune_base = ws + pl.dY_bytes + pl.J_bytes + pl.part_bytes + pl.dW_bytes + pl.db_bytes;
    int* prune_stats = reinterpret_cast<int*>(prune_base);
    uint8_t* prune_flags = prune_base + 64;
    int* prune_ids = reinterpret_cast<int*>(prune_base + 64 + align_up((size_t)pl.chunk_tiles, 16));
    const bool prune = prune_log2_eps < 0.f;
    bp.NV2 = pl.NV2; bp.NHU = pl.NHU;
    for (int i = 0; i < 3; ++i) { bp.hu_blk[i] = pl.hu_blk[i]; bp.hu_item0[i] = pl.hu_item0[i]; }
    bp.hu_splits[0] = pl.hu_splits[0]; bp.hu_splits[1] = pl.hu_splits[1];
    bp.enc = reinterpret_cast<const __nv_bfloat16*>(enc);
    bp.dec = reinterpret_cast<const __nv_bfloat16*>(dec);
    bp.NHC = (H + kDjChunkH - 1) / kDjChunkH;

    JointMaps maps;
    if (int rc = make_joint_maps(&maps, jp, enc, dec, W)) return rc;
    CUtensorMap tmap_dj;
    if (int rc = make_tmap_2d_bf16(&tmap_dj, W, (uint64_t)V, (uint64_t)H, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;

    cudaError_t e;
    e = cudaMemsetAsync(d_enc, 0, sizeof(float) * (size_t)B * T * H, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(d_enc)");
    e = cudaMemsetAsync(d_dec, 0, sizeof(float) * (size_t)B * U * H, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(d_dec)");
    e = cudaMemsetAsync(prune_stats, 0, 64, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(prune stats)");

    // operand images as 2-D tensors [images * 128 rows, 64]: an un-swizzled box copies image rows verbatim
    CUtensorMap tmap_dy, tmap_dy_half, tmap_j_half;
    if (int rc = make_tmap_2d_bf16(&tmap_dy, jp.dY_img, (uint64_t)pl.chunk_tiles * pl.NT4 * 128, 64, 64, 128, CU_TENSOR_MAP_SWIZZLE_NONE))
        return rc;
    if (int rc = make_tmap_2d_bf16(&tmap_dy_half, jp.dY_img, (uint64_t)pl.chunk_tiles * pl.NT4 * 128, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE))
        return rc;
    if (int rc = make_tmap_2d_bf16(&tmap_j_half, jp.J_img, (uint64_t)pl.chunk_tiles * jp.KB * 128, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE))
        return rc;

    const DjSmem djL = dj_smem_layout();
    const DwSmem dwL = dw_smem_layout();
    auto dj_kern = jp.tT_log2 == 3 ? dj_gemm_kernel<3> : (jp.tT_log2 == 4 ? dj_gemm_kernel<4> : dj_gemm_kernel<5>);
    e = cudaFuncSetAttribute(dj_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)djL.total);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dj_gemm_kernel)");
    e = cudaFuncSetAttribute(dw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dwL.total);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dw_gemm_kernel)");

    const int total_tiles = B * jp.nTt * jp.nTu;
    const int tU = 128 >> jp.tT_log2;
    int chunk_idx = 0;
    KernelTimer timer;
    for (int t0 = 0; t0 < total_tiles; t0 += pl.chunk_tiles, ++chunk_idx) {
        const int t1 = t0 + pl.chunk_tiles < total_tiles ? t0 + pl.chunk_tiles : total_tiles;
        jp.tile_begin = bp.tile_begin = t0;
        jp.tile_end = bp.tile_end = t1;
        if (prune) {
            // which tiles of the chunk carry a non-negligible share of the alignment posterior?
            ScopedTiming tm("tile_activity+compact", st);
            const int n = t1 - t0;
            e = launch_pdl(tile_activity_kernel, dim3((n + 7) / 8), dim3(256), 0, st, bp, alpha, beta, cost,
                           prune_log2_eps * 0.6931471805599453f, prune_flags);
            if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "tile_activity_kernel launch");
            e = launch_pdl(compact_active_kernel, dim3(1), dim3(1024), 0, st, bp, (const uint8_t*)prune_flags, prune_ids, prune_stats);
            if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "compact_active_kernel launch");
            g_launches += 2;
            jp.active_ids = bp.active_ids = prune_ids;
            jp.active_count = bp.active_count = prune_stats;
            bp.tile_flags = prune_flags;
        }
        {
            ScopedTiming tm("joint_gemm_kernel<GRAD>", st);
            if (int rc = clamp > 0.f ? launch_joint<MODE_GRAD_CLAMP>(maps, jp, sms, st) : launch_joint<MODE_GRAD>(maps, jp, sms, st)) return rc;
        }

        {
            const int dj_cs = 2;
            const int n_units = ((t1 - t0 + 1) / 2) * bp.NHC;  // (tile pair, chunk of 256 h-rows)
            const int max_clusters = sms / dj_cs;
            const int n_clusters = n_units < max_clusters ? n_units : max_clusters;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(n_clusters * dj_cs);
            cfg.blockDim = dim3(kDjThreads);
            cfg.dynamicSmemBytes = djL.total;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = dj_cs;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_launch_attr(attr, 1);
            static const bool prof_on = getenv("TSASR_DEBUG_PROF") != nullptr;
            long long* d_prof = nullptr;
            BwdParams bpp = bp;
            if (prof_on) {
                cudaMalloc(&d_prof, sizeof(long long) * 4 * n_clusters * dj_cs);
                cudaMemset(d_prof, 0, sizeof(long long) * 4 * n_clusters * dj_cs);
                bpp.prof = d_prof;
            }
            {
                ScopedTiming tm("dj_gemm_kernel", st);
                e = cudaLaunchKernelEx(&cfg, dj_kern, tmap_dj, tmap_dy, bpp);
            }
            if (prof_on) {
                cudaStreamSynchronize(st);
                const int n = n_clusters * dj_cs;
                long long* h = new long long[4 * n];
                cudaMemcpy(h, d_prof, sizeof(long long) * 4 * n, cudaMemcpyDeviceToHost);
                double tot = 0, acc = 0, full = 0, units = 0;
                int nl = 0;
                for (int i = 0; i < n; ++i)
                    if (h[4 * i] > 0) { tot += h[4 * i]; acc += h[4 * i + 1]; full += h[4 * i + 2]; units += h[4 * i + 3]; ++nl; }
                fprintf(stderr, "[tsasr prof] dj issuing ctas=%d units/cta=%.1f cycles/cta=%.0f wait: acc_empty=%.1f%% full=%.1f%% other=%.1f%% cycles/unit=%.0f\n",
                        nl, units / nl, tot / nl, 100 * acc / tot, 100 * full / tot, 100 * (tot - acc - full) / tot, tot / units);
                delete[] h;
                cudaFree(d_prof);
            }
            ++g_launches;
            if (e != cudaSuccess) return cuda_fail(e, "dj_gemm_kernel launch");
            if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "dj_gemm_kernel launch");
        }

        {
            const int tT = 1 << jp.tT_log2;
            const int g0 = t0 / jp.nTu, g1 = t1 / jp.nTu;
            const int b0 = g0 / jp.nTt, b1 = (g1 - 1) / jp.nTt + 1;
            const int rows = (g1 - g0) * tT + (b1 - b0) * U;
            ScopedTiming tm("reduce_dpre_kernel", st);
            e = launch_pdl(reduce_dpre_kernel, dim3(rows), dim3(160), 0, st, bp, d_enc, d_dec);
            ++g_launches;
            if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "reduce_dpre_kernel launch");
        }

        bp.accumulate = chunk_idx > 0;
        {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(2 * pl.dw_pairs);
            cfg.blockDim = dim3(kBwdThreads);
            cfg.dynamicSmemBytes = dwL.total;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_launch_attr(attr, 1);
            static const bool prof_on = getenv("TSASR_DEBUG_PROF") != nullptr;
            long long* d_prof = nullptr;
            BwdParams bpp = bp;
            const int n = 2 * pl.dw_pairs;
            if (prof_on) {
                cudaMalloc(&d_prof, sizeof(long long) * 4 * n);
                cudaMemset(d_prof, 0, sizeof(long long) * 4 * n);
                bpp.prof = d_prof;
            }
            {
                ScopedTiming tm("dw_gemm_kernel", st);
                e = cudaLaunchKernelEx(&cfg, dw_gemm_kernel, tmap_dy_half, tmap_j_half, bpp);
            }
            if (prof_on) {
                cudaStreamSynchronize(st);
                long long* h = new long long[4 * n];
                cudaMemcpy(h, d_prof, sizeof(long long) * 4 * n, cudaMemcpyDeviceToHost);
                {
                    double tot = 0, full = 0, stages = 0, mx = 0;
                    int nl = 0;
                    for (int i = 0; i < n; ++i)
                        if (h[4 * i] > 0) { tot += h[4 * i]; full += h[4 * i + 2]; stages += h[4 * i + 3]; ++nl; if (h[4 * i] > mx) mx = h[4 * i]; }
                    if (nl) fprintf(stderr, "[tsasr prof] dw splits=(%d,%d) items=%d pairs=%d: cycles/pair avg=%.0f max=%.0f cycles/stage=%.0f wait full=%.1f%%\n",
                                    pl.hu_splits[0], pl.hu_splits[1], pl.hu_item0[2], nl, tot / nl, mx, tot / stages, 100 * full / tot);
                }
                delete[] h;
                cudaFree(d_prof);
            }
            if (e != cudaSuccess) return cuda_fail(e, "dw_gemm_kernel launch");
        }
        ++g_launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "dw_gemm_kernel launch");
    }
    (void)tU;
    {
        ScopedTiming tm("reduce_dw_kernel", st);
        const int dw_blocks = (int)(((size_t)V * H / 4 + 255) / 256);
        e = launch_pdl(reduce_dw_kernel, dim3(dw_blocks < sms * 8 ? dw_blocks : sms * 8), dim3(256), 0, st, bp, dW, db);
    }
    ++g_launches;
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "reduce_dw_kernel launch");
    return TSASR_OK;
}

}  // extern "C"
