// predictor_capi.inl -- tsasr_lstm_fwd / tsasr_lstm_bwd / tsasr_onehot_dw (included by capi.cu): the prediction network.

namespace {

static int check_lstm_dims(int B, int U, int Hd) {
    REQUIRE(B >= 1 && U >= 1, "B and U must be >= 1 (got %d %d)", B, U);
    if (Hd != 128 && Hd != 256 && Hd != 512)
        return fail(TSASR_E_UNSUPPORTED, "the recurrent kernels keep 4 x Hd/32 weights per gate row in registers: Hd must be 128, 256 or 512 (got %d)", Hd);
    if (B > kLstmBatchTile * kLstmMaxPasses)
        return fail(TSASR_E_UNSUPPORTED, "the recurrent kernels hold one state per (thread, batch tile): B <= %d (got %d)", kLstmBatchTile * kLstmMaxPasses, B);
    return TSASR_OK;
}

static int lstm_debug_flags() {  // development ablations, see LstmFwdParams::dbg
    static const int v = getenv("TSASR_DEBUG_LSTM") ? atoi(getenv("TSASR_DEBUG_LSTM")) : 0;
    return v;
}

// the hand-off buffer of a recurrent kernel starts as all-sentinel (0xFFFFFFFF), see predictor.cu
static int fill_sentinel(void* buf, size_t bytes, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(buf, 0xFF, bytes, st);
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "cudaMemsetAsync (lstm hand-off buffer)");
}

}  // namespace

extern "C" {

int tsasr_lstm_fwd(const void* tokens, int tokens_i64, int blank, int n_embed, const float* xw, const float* W_ih, const float* W_hh,
                   const float* b_ih, const float* b_hh, const float* rel_lengths, const int32_t* abs_lengths, int B, int U, int Hd,
                   float* out, float* hprev, float* gates, float* cells, float* h_n, float* c_n, int32_t* lengths_out,
                   tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_lstm_fwd");
    if (int rc = check_lstm_dims(B, U, Hd)) return rc;
    REQUIRE(W_hh && out && (rel_lengths || abs_lengths), "null pointer argument");
    REQUIRE((tokens != nullptr) != (xw != nullptr), "pass either tokens (one-hot input) or xw (dense input, x W_ih^T + b_ih precomputed)");
    REQUIRE(!tokens || (W_ih && n_embed >= 1), "one-hot input needs W_ih [4Hd, n_embed]");
    REQUIRE((reinterpret_cast<uintptr_t>(W_hh) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "W_hh and out must be 16-byte aligned");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    if (lstm_fwd_smem_bytes(B, U, Hd, tokens != nullptr) > (size_t)max_smem)
        return fail(TSASR_E_UNSUPPORTED, "B * U = %d token positions do not fit the shared memory of one CTA", B * U);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LstmFwdParams p;
    memset(&p, 0, sizeof(p));
    if (int rc = fill_sentinel(out, (size_t)B * U * Hd * sizeof(float), st)) return rc;
    p.tok64 = tokens && tokens_i64 ? static_cast<const long long*>(tokens) : nullptr;
    p.tok32 = tokens && !tokens_i64 ? static_cast<const int*>(tokens) : nullptr;
    p.blank = blank; p.n_embed = n_embed;
    p.xw = xw; p.W_ih = W_ih; p.W_hh = W_hh; p.b_ih = b_ih; p.b_hh = b_hh;
    p.rel_lengths = rel_lengths; p.abs_lengths = abs_lengths;
    p.B = B; p.U = U; p.Hd = Hd;
    p.dbg = lstm_debug_flags();
    p.out = out; p.hprev = hprev; p.gates = gates; p.cells = cells; p.h_n = h_n; p.c_n = c_n; p.lengths_out = lengths_out;
    ScopedTiming tm("lstm_seq_fwd_kernel", st);
    cudaError_t e = launch_lstm_fwd(p, st);
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "lstm_seq_fwd_kernel");
}

int tsasr_lstm_bwd(const float* d_out, const float* d_hn, const float* d_cn, const float* W_hh, const float* gates, const float* cells,
                   const int32_t* lengths, int B, int U, int Hd, float* dG, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_lstm_bwd");
    if (int rc = check_lstm_dims(B, U, Hd)) return rc;
    REQUIRE(d_out && W_hh && gates && cells && lengths && dG, "null pointer argument");
    REQUIRE((reinterpret_cast<uintptr_t>(W_hh) & 15) == 0 && (reinterpret_cast<uintptr_t>(dG) & 15) == 0, "W_hh and dG must be 16-byte aligned");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LstmBwdParams p;
    memset(&p, 0, sizeof(p));
    if (int rc = fill_sentinel(dG, (size_t)B * U * 4 * Hd * sizeof(float), st)) return rc;
    p.d_out = d_out; p.d_hn = d_hn; p.d_cn = d_cn; p.W_hh = W_hh; p.gates = gates; p.cells = cells; p.lengths = lengths;
    p.B = B; p.U = U; p.Hd = Hd; p.dG = dG;
    p.dbg = lstm_debug_flags();
    ScopedTiming tm("lstm_seq_bwd_kernel", st);
    cudaError_t e = launch_lstm_bwd(p, st);
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "lstm_seq_bwd_kernel");
}

int tsasr_onehot_dw(const void* tokens, int tokens_i64, int blank, int n_embed, const float* dG, int n_pos, int G, float* dW_ih,
                    tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_onehot_dw");
    REQUIRE(tokens && dG && dW_ih, "null pointer argument");
    REQUIRE(n_embed >= 1 && n_pos >= 1 && G >= 1, "n_embed, n_pos, G must be >= 1 (got %d %d %d)", n_embed, n_pos, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScopedTiming tm("onehot_dw_kernel", st);
    cudaError_t e = launch_onehot_dw(tokens, tokens_i64, blank, n_embed, dG, n_pos, G, dW_ih, st);
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "onehot_dw_kernel");
}

}  // extern "C"
