// linear_capi.inl -- tsasr_linear_fwd / tsasr_linear_bwd (included by capi.cu): the projections either side of the joint.

namespace {

static constexpr int kLinTargetCtas = 296;  // one wave of a B200: 148 SMs x 2 resident CTAs
static constexpr int kLinMaxSplits = 32;

struct LinSplit { int tiles_m, tiles_n, splits, kb_per_split; };

// split-K plan of the dW product (M = N_out rows, N = K_in columns, contraction over the R rows); shape-only, so that
// tsasr_linear_bwd_workspace_bytes and tsasr_linear_bwd agree without looking at the device
static LinSplit lin_dw_split(int R, int K, int N) {
    LinSplit s;
    s.tiles_m = (N + kLinBM - 1) / kLinBM;
    s.tiles_n = (K + kLinBN - 1) / kLinBN;
    const int kb_total = (R + kLinBK - 1) / kLinBK;
    int want = kLinTargetCtas / (s.tiles_m * s.tiles_n);
    if (want < 1) want = 1;
    if (want > kLinMaxSplits) want = kLinMaxSplits;
    if (want > kb_total) want = kb_total;
    s.kb_per_split = (kb_total + want - 1) / want;
    s.splits = (kb_total + s.kb_per_split - 1) / s.kb_per_split;  // no empty slice
    return s;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool AK, bool BK>
static int launch_linear(const LinParams& p, dim3 grid, cudaStream_t st, const char* name) {
    auto kern = linear_gemm_kernel<AK, BK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLinSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(linear_gemm_kernel)");
    ScopedTiming tm(name, st);
    kern<<<grid, kLinThreads, kLinSmemBytes, st>>>(p);
    ++g_launches;
    e = cudaGetLastError();
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, name);
}

static int check_linear_dims(int R, int K, int N) {
    REQUIRE(R >= 1 && K >= 1 && N >= 1, "R, K, N must be >= 1 (got %d %d %d)", R, K, N);
    REQUIRE((long long)R * K < (1ll << 40) && (long long)R * N < (1ll << 40), "operand too large");
    return TSASR_OK;
}

}  // namespace

extern "C" {

size_t tsasr_linear_bwd_workspace_bytes(int R, int K, int N) {
    if (R < 1 || K < 1 || N < 1) return 0;
    const LinSplit s = lin_dw_split(R, K, N);
    if (s.splits == 1) return 256;
    return ((size_t)s.splits * ((size_t)N * K + (size_t)N) * sizeof(float) + 255) / 256 * 256;
}

int tsasr_linear_fwd(const float* X, const float* W, const float* bias, int R, int K, int N, float* Y, void* Y_bf16,
                     tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_linear_fwd");
    if (int rc = check_linear_dims(R, K, N)) return rc;
    REQUIRE(X && W && (Y || Y_bf16), "null pointer argument");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    LinParams p;
    memset(&p, 0, sizeof(p));
    p.A = X; p.a_ld_mn = K; p.a_ld_k = 1;
    p.B = W; p.b_ld_mn = K; p.b_ld_k = 1;
    p.bias = bias;
    p.C32 = Y; p.C16 = reinterpret_cast<__nv_bfloat16*>(Y_bf16); p.c_ld = N;
    p.M = R; p.N = N; p.K = K;
    p.kb_per_split = (K + kLinBK - 1) / kLinBK;
    p.a_vec = (K % 4 == 0) && aligned16(X);
    p.b_vec = (K % 4 == 0) && aligned16(W);
    p.c_vec = (N % 4 == 0) && (!Y || aligned16(Y)) && (!Y_bf16 || (reinterpret_cast<uintptr_t>(Y_bf16) & 7) == 0) && (!bias || aligned16(bias));
    const dim3 grid((R + kLinBM - 1) / kLinBM, (N + kLinBN - 1) / kLinBN, 1);
    return launch_linear<true, true>(p, grid, static_cast<cudaStream_t>(stream), "linear_gemm_kernel<fwd>");
}

int tsasr_linear_bwd(const float* dY, const float* X, const float* W, int R, int K, int N, float* dX, float* dW, float* db,
                     void* workspace, size_t workspace_bytes, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_linear_bwd");
    if (int rc = check_linear_dims(R, K, N)) return rc;
    REQUIRE(dY && (dX || dW), "null pointer argument");
    REQUIRE(!dX || W, "dX needs W");
    REQUIRE(!db || dW, "db is produced by the dW product: pass dW too");
    REQUIRE(!dW || X, "dW needs X");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dX) {  // dX[r, k] = sum_n dY[r, n] W[n, k]
        LinParams p;
        memset(&p, 0, sizeof(p));
        p.A = dY; p.a_ld_mn = N; p.a_ld_k = 1;
        p.B = W; p.b_ld_mn = 1; p.b_ld_k = K;
        p.C32 = dX; p.c_ld = K;
        p.M = R; p.N = K; p.K = N;
        p.kb_per_split = (N + kLinBK - 1) / kLinBK;
        p.a_vec = (N % 4 == 0) && aligned16(dY);
        p.c_vec = (K % 4 == 0) && aligned16(dX);
        const dim3 grid((R + kLinBM - 1) / kLinBM, (K + kLinBN - 1) / kLinBN, 1);
        if (int rc = launch_linear<true, false>(p, grid, st, "linear_gemm_kernel<dX>")) return rc;
    }
    if (dW) {  // dW[n, k] = sum_r dY[r, n] X[r, k];  db[n] = sum_r dY[r, n] (the row of ones)
        const LinSplit s = lin_dw_split(R, K, N);
        const size_t need = tsasr_linear_bwd_workspace_bytes(R, K, N);
        float* ws = static_cast<float*>(workspace);
        if (s.splits > 1) {
            REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
            if (workspace_bytes < need) return fail(TSASR_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        }
        LinParams p;
        memset(&p, 0, sizeof(p));
        p.A = dY; p.a_ld_mn = 1; p.a_ld_k = N;
        p.B = X; p.b_ld_mn = 1; p.b_ld_k = K;
        p.M = N; p.N = K; p.K = R;
        p.c_ld = K;
        p.kb_per_split = s.kb_per_split;
        float* part = ws;                                    // [splits][N*K]
        float* ones_part = (ws && db) ? ws + (size_t)s.splits * N * K : nullptr;  // [splits][N]
        if (s.splits == 1) {
            p.C32 = dW;
            p.ones_out = db;
        } else {
            p.C32 = part;
            p.c_split_stride = (long long)N * K;
            p.ones_out = ones_part;
        }
        p.c_vec = (K % 4 == 0) && aligned16(p.C32);
        const dim3 grid(s.tiles_m, s.tiles_n, s.splits);
        if (int rc = launch_linear<false, false>(p, grid, st, "linear_gemm_kernel<dW>")) return rc;
        if (s.splits > 1) {
            const long long n = (long long)N * K;
            long long blocks = (n / 4 + 255) / 256;
            if (blocks > sms * 8) blocks = sms * 8;
            if (blocks < 1) blocks = 1;
            ScopedTiming tm("linear_fold_kernel", st);
            linear_fold_kernel<<<(unsigned)blocks, 256, 0, st>>>(part, (long long)N * K, s.splits, n, dW, ones_part, N, db);
            ++g_launches;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "linear_fold_kernel");
        }
    }
    return TSASR_OK;
}

}  // extern "C"
