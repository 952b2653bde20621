/*
 * tsasr_b200.h -- C ABI of the B200-native joint + RNN-T loss hot path (libtsasr_b200.so).
 *
 * The reference (lucadellalib/ts-asr) has no native boundary of its own: it is 100 % Python and
 * reaches its kernels through torch.ops.torchaudio.rnnt_loss_forward and Numba @cuda.jit launches.
 * Each entry point below cites the reference interface it replaces (paths relative to the
 * reference root; SB = vendor/speechbrain/speechbrain).  INTEGRATION.md shows the ctypes binding a
 * SpeechBrain maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless noted; nothing is allocated inside the library:
 *     the caller owns all buffers (torch tensors) and passes a workspace where one is needed;
 *   - `stream` is a cudaStream_t (the caller's current stream); no call synchronises the host;
 *   - every function returns 0 on success or a negative TSASR_E_* code; tsasr_last_error() gives
 *     the message (thread-local); nothing throws;
 *   - B utterances, T = max frames, U = lattice width = max labels + 1, V = vocabulary incl. blank,
 *     H = joint dimension;  logit_lengths[b] = T_b (1..T), target_lengths[b] = label count (0..U-1);
 *     targets is int32 [B, U-1];
 *   - lattice-sized arrays (lat2, den/logz, alpha, beta) hold tsasr_lattice_elems(B,T,U) elements in
 *     the skewed layout: cell (b,t,u) at ((b*(T+U-1) + t+u) * U + u).
 */
#ifndef TSASR_B200_H
#define TSASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSASR_ABI_VERSION 4

enum {
    TSASR_OK = 0,
    TSASR_E_INVALID = -1,     /* bad argument (shape, dtype, blank, alignment) */
    TSASR_E_UNSUPPORTED = -2, /* shape outside what the sm_100a kernels implement */
    TSASR_E_CUDA = -3,        /* CUDA runtime / driver error, see tsasr_last_error() */
    TSASR_E_WORKSPACE = -4    /* workspace too small */
};

/* logits dtype codes */
enum { TSASR_F32 = 0, TSASR_F16 = 1, TSASR_BF16 = 2 };

/* activation codes: Transducer_joint(nonlinearity=...)  (SB/nnet/transducer/transducer_joint.py:40-46) */
enum { TSASR_ACT_LEAKY_RELU = 0, TSASR_ACT_RELU = 1, TSASR_ACT_TANH = 2, TSASR_ACT_IDENTITY = 3 };

typedef void* tsasr_stream_t; /* cudaStream_t */

int tsasr_abi_version(void);
const char* tsasr_last_error(void);

/* Number of elements of one lattice-sized array: B * (T+U-1) * U. */
size_t tsasr_lattice_elems(int B, int T, int U);

/* ---- compat path: materialised logits in, dense dlogits out ------------------------------------
 * Replaces torchaudio.functional.rnnt_loss as called at SB/nnet/losses.py:72-79
 * (torch.ops.torchaudio.rnnt_loss_forward: ReduceMax2D / ReduceLogSumExpGivenMax2D / ComputeLogProbs),
 * and `logits.log_softmax(-1)` + the label/blank gathers of the Numba kernels
 * (SB/nnet/losses.py:84, SB/nnet/loss/transducer_loss.py:81-90,160-166).
 * normalized != 0: rows are already log-probs (Transducer.apply input), den is set to 0.
 * lat2: float2 {lp_blank, lp_emit} per cell; den: log-sum-exp per cell. */
int tsasr_logits_to_lattice(const void* logits, int logits_dtype, const int32_t* targets,
                            const int32_t* logit_lengths, const int32_t* target_lengths, int B, int T, int U, int V,
                            int blank, int normalized, float* lat2, float* den, tsasr_stream_t stream);

/* Anti-diagonal wavefront forward/backward DP.  Replaces cu_kernel_forward / cu_kernel_backward
 * (SB/nnet/loss/transducer_loss.py:31-106,109-180) and torchaudio's ComputeAlphasBetasCosts.
 * Outputs alpha, beta (lattice-sized), cost[b] = -log P(y_b|x_b) (from beta(0,0)), and the two
 * log-likelihoods ll_alpha[b], ll_beta[b] (each B floats) for consistency checks.
 * U <= 1024: one thread per lattice column, cp.async prefetch ring (the tuned kernel); 1024 < U <= 8192: several
 * columns per thread, anti-diagonals double-buffered in shared memory (the reference's Numba kernels stop at 1024
 * threads = columns, torchaudio has no limit); wider lattices return TSASR_E_UNSUPPORTED. */
int tsasr_lattice_alpha_beta(const float* lat2, const int32_t* logit_lengths, const int32_t* target_lengths, int B,
                             int T, int U, float* alpha, float* beta, float* cost, float* ll_alpha, float* ll_beta,
                             tsasr_stream_t stream);

/* Dense gradient d cost_b / d logits * dcost[b] with the softmax folded in (torchaudio
 * ComputeGradients; SB/nnet/losses.py:72-79 backward).  dcost may be NULL (= 1).  clamp <= 0: off.
 * dlogits has the dtype and shape of logits; cells outside T_b x U_b are written as zeros. */
int tsasr_logits_grad(const void* logits, int logits_dtype, const int32_t* targets, const int32_t* logit_lengths,
                      const int32_t* target_lengths, int B, int T, int U, int V, int blank, const float* lat2,
                      const float* den, const float* alpha, const float* beta, const float* cost, const float* dcost,
                      float clamp, void* dlogits, tsasr_stream_t stream);

/* Sparse gradient w.r.t. log-probs, fp32 [B,T,U,V] (zero-filled here), Numba semantics:
 * cu_kernel_compute_grad, SB/nnet/loss/transducer_loss.py:183-236. */
int tsasr_logprobs_grad(const int32_t* targets, const int32_t* logit_lengths, const int32_t* target_lengths, int B,
                        int T, int U, int V, int blank, const float* lat2, const float* alpha, const float* beta,
                        const float* cost, const float* dcost, float* grads, tsasr_stream_t stream);

/* ---- fused path: joint + head + log-softmax, 4-D tensors never materialised ---------------------
 * Replaces Transducer_joint.forward (joint="sum": SB/nnet/transducer/transducer_joint.py:73-74,95),
 * Linear.forward of the transducer head (SB/nnet/linear.py:74) and the log-softmax + gathers of the
 * loss, as chained at train_librispeechmix_scratch.py:132,135,158.
 *   enc  bf16 [B,T,H]   dec bf16 [B,U,H]   W bf16 [V,H]   bias fp32 [V]
 *   out: lat2 (float2 per cell), logz (log-sum-exp per cell), skewed layout.
 * Requirements: H % 64 == 0, 64 <= H <= 640, V >= 2 (a host that has another H <= 640 zero-pads enc, dec and W to the next
 * multiple of 64, which is exact; tsasr_b200/functional.py does).  tcgen05 / TMEM / TMA kernel. */
int tsasr_joint_fwd(const void* enc, const void* dec, const void* W, const float* bias, const int32_t* targets,
                    const int32_t* logit_lengths, const int32_t* target_lengths, int B, int T, int U, int H, int V,
                    int blank, int act_kind, float act_param, float* lat2, float* logz, tsasr_stream_t stream);

/* The whole forward of the fused loss behind ONE call: input preparation (bf16 operand copies when the operands are
 * fp32, int64 -> int32 targets as SB/nnet/losses.py:74 casts them, the length conversion of SB/nnet/losses.py:58-59 and
 * the statistics of tsasr_prepare_lengths -- one launch), tsasr_joint_fwd and tsasr_lattice_alpha_beta, queued back to
 * back on `stream` (what train_librispeechmix_scratch.py:132,135,158 + torchaudio's forward amount to).
 *   operand_dtype: TSASR_F32 (enc, dec, W are fp32 and converted into scratch) or TSASR_BF16 (used as they are);
 *   targets: int32 [B,U-1], or int64 when targets_i64 != 0;  lengths: rel_* fp32 (SpeechBrain relative) or abs_* int32;
 *   scratch: 256-byte aligned, tsasr_joint_loss_fwd_layout() gives the byte offsets
 *     {enc16, dec16, W16, targets32, logit_lengths, target_lengths, stats[4], total}; the int32 lengths, the statistics
 *     and -- where a conversion happened -- the bf16 operands / int32 targets live there and feed tsasr_joint_bwd;
 *   stats_host (may be NULL): HOST pointer to 8 int32 of mapped pinned memory (cudaHostAlloc / torch pin_memory); the
 *     preparation kernel writes {max T_b, max labels, min T_b, min labels} there, a system-scope fence, then
 *     stats_host[4] = stats_seq -- the caller polls the tag and performs torchaudio's argument checks without a stream,
 *     event or copy of its own (the kernels clamp every length, so nothing depends on the outcome);
 *   cost3: 3*B floats {cost[b] = -log P, ll_alpha[b], ll_beta[b]};
 *   enc_bf16 / dec_bf16 (may be NULL; only with operand_dtype == TSASR_F32): bf16 copies of enc / dec that already exist --
 *     the second output of tsasr_linear_fwd (encoder_proj / decoder_proj) -- used as they are instead of converting that
 *     operand again; W is still converted. */
int tsasr_joint_loss_fwd_layout(int B, int T, int U, int H, int V, size_t* offsets8);
int tsasr_joint_loss_fwd(const void* enc, const void* dec, const void* W, int operand_dtype, const float* bias, const void* targets,
                         int targets_i64, const float* rel_logit_lengths, const float* rel_target_lengths,
                         const int32_t* abs_logit_lengths, const int32_t* abs_target_lengths, int B, int T, int U, int H, int V,
                         int blank, int act_kind, float act_param, void* scratch, size_t scratch_bytes, int32_t* stats_host,
                         int stats_seq, float* lat2, float* logz, float* alpha, float* beta, float* cost3, const void* enc_bf16,
                         const void* dec_bf16, tsasr_stream_t stream);

/* Workspace (bytes) tsasr_joint_bwd needs; bounded independently of B*T*U by `max_chunk_cells`
 * (0 = library default). */
size_t tsasr_joint_bwd_workspace_bytes(int B, int T, int U, int H, int V, long long max_chunk_cells);

/* Backward of the fused chain (autograd of Linear + activation + broadcast add fed by the loss
 * gradient; reference: SB/core.py:1077 loss.backward()).  Recomputes the logits tile-wise, forms
 * dlogits in bf16 chunk by chunk (never the whole [B,T,U,V]) and runs the two backward GEMMs.
 *   out: d_enc fp32 [B,T,H], d_dec fp32 [B,U,H], dW fp32 [V,H], db fp32 [V]  (all overwritten).
 * prune_log2_eps < 0: 128-cell tiles whose largest alignment posterior exp(alpha + beta - L) is below
 * 2^prune_log2_eps are left out of the backward (every term of their dlogits carries that factor; at -30 what is
 * dropped is below fp32 resolution next to the O(1) terms of the alignment band).  >= 0: every live tile is processed.
 * clamp > 0: torchaudio's rnnt_loss(clamp=...) -- the dlogits of the UNIT cost are clamped to [-clamp, clamp] before the
 * upstream factor dcost[b] multiplies them (ComputeGradients, then `grad * dy`: torchaudio/functional/functional.py:1729-1734),
 * inside the gradient pass (a separate kernel instantiation: the default path pays nothing).  <= 0: off.
 * tsasr_joint_bwd_stats_offset(): byte offset, from the workspace base rounded up to 1024, of three int32
 * {active tiles of the last chunk, active tiles, live tiles} written by the call (measurement aid). */
size_t tsasr_joint_bwd_stats_offset(int B, int T, int U, int H, int V, long long max_chunk_cells);
int tsasr_joint_bwd(const void* enc, const void* dec, const void* W, const float* bias, const int32_t* targets,
                    const int32_t* logit_lengths, const int32_t* target_lengths, int B, int T, int U, int H, int V,
                    int blank, int act_kind, float act_param, const float* lat2, const float* logz,
                    const float* alpha, const float* beta, const float* cost, const float* dcost, void* workspace,
                    size_t workspace_bytes, long long max_chunk_cells, float prune_log2_eps, float clamp, float* d_enc,
                    float* d_dec, float* dW, float* db, tsasr_stream_t stream);

/* ---- host-path helpers of the fused loss (one launch each instead of a dozen elementwise launches) -------
 * tsasr_prepare_lengths: the integer length conversion of SB/nnet/losses.py:58-59, bit-exact
 * ((rel * dim) in fp32, round-half-to-even, int32) for both vectors -- pass rel_* = NULL and abs_* instead when the
 * lengths are already absolute -- plus stats_out[4] = {max T_b, max labels, min T_b, min labels}, the numbers
 * torchaudio's argument checks compare against the tensor shapes.  *_out may be NULL with abs_* inputs. */
int tsasr_prepare_lengths(const float* rel_logit_lengths, const float* rel_target_lengths, const int32_t* abs_logit_lengths,
                          const int32_t* abs_target_lengths, int B, int T, int n_targets, int32_t* logit_lengths_out,
                          int32_t* target_lengths_out, int32_t* stats_out, tsasr_stream_t stream);
/* fp32 -> bf16 (round to nearest even) of the three GEMM operands in one launch; counts are in elements. */
int tsasr_cast_operands_bf16(const float* enc, size_t n_enc, const float* dec, size_t n_dec, const float* W, size_t n_w,
                             void* enc16, void* dec16, void* W16, tsasr_stream_t stream);

/* ---- the projections either side of the joint -------------------------------------
 * Replaces speechbrain.nnet.linear.Linear.forward (SB/nnet/linear.py:63-76) as instantiated for encoder_proj / decoder_proj
 * (hparams/LibriSpeechMix/conformer-t_scratch.yaml:172-174,187-189; train_librispeechmix_scratch.py:122,127) and the
 * autograd backward of nn.Linear.  Row-major fp32 everywhere: X [R,K], W [N,K], bias [N] or NULL, Y [R,N].
 * tcgen05 GEMMs on in-kernel bf16 (hi, lo) splits of the fp32 operands (3 MMAs per step): 16-17 bits per term
 * (max error ~1e-5 of the largest output at K = 256), no operand copies.
 *   tsasr_linear_fwd: Y = X W^T + bias written as fp32 (Y, may be NULL) and/or bf16 (Y_bf16, may be NULL: the operand
 *     image tsasr_joint_loss_fwd takes as enc_bf16 / dec_bf16) in the same pass.
 *   tsasr_linear_bwd: dX = dY W (skipped when dX is NULL), dW = dY^T X and db = column sums of dY (db may be NULL; dW may
 *     be NULL only together with db), all overwritten; dY is the fp32 d_enc / d_dec of tsasr_joint_bwd, consumed as it is.
 *     workspace: tsasr_linear_bwd_workspace_bytes(R,K,N) bytes, 256-byte aligned (split-K partials, folded in a fixed
 *     order: deterministic). */
size_t tsasr_linear_bwd_workspace_bytes(int R, int K, int N);
int tsasr_linear_fwd(const float* X, const float* W, const float* bias, int R, int K, int N, float* Y, void* Y_bf16,
                     tsasr_stream_t stream);
int tsasr_linear_bwd(const float* dY, const float* X, const float* W, int R, int K, int N, float* dX, float* dW, float* db,
                     void* workspace, size_t workspace_bytes, tsasr_stream_t stream);

/* ---- the prediction network ---------------------------------------------------------
 * Replaces speechbrain.nnet.embedding.Embedding(consider_as_one_hot=True) (SB/nnet/embedding.py:65-114) followed by
 * speechbrain.nnet.RNN.LSTM (one layer, unidirectional; SB/nnet/RNN.py:170-278, torch.nn.LSTM on a PackedSequence) as
 * chained at train_librispeechmix_scratch.py:125-126 (hparams/LibriSpeechMix/conformer-t_scratch.yaml:176-185).
 * fp32 arithmetic throughout (gate order i, f, g, o; h_0 = c_0 = 0).
 *   tsasr_lstm_fwd: the whole teacher-forced recurrence in one cooperative launch.
 *     input: EITHER tokens [B,U] (int32, or int64 when tokens_i64 != 0) + W_ih [4Hd, n_embed]: the one-hot embedding of
 *       token k is column k - [k > blank] of W_ih (nothing for k == blank), gathered -- no [B,U,V-1] tensor, no GEMM;
 *       OR xw [B,U,4Hd] = x W_ih^T + b_ih precomputed (tsasr_linear_fwd) for a dense input.
 *     lengths: rel_lengths fp32 (SpeechBrain relative; converted as SB/nnet/RNN.py:35 + pack_padded_sequence do: fp32
 *       product, truncation) or abs_lengths int32; positions u >= length read as zeros in `out` and freeze the state.
 *     out [B,U,Hd]; optional (NULL to skip): hprev [B,U,Hd] = h_{u-1}, gates [B,U,4,Hd] (activated), cells [B,U,Hd]
 *       (what tsasr_lstm_bwd needs), h_n / c_n [B,Hd], lengths_out [B] int32.
 *     Hd in {128, 256, 512}, B <= 64; else TSASR_E_UNSUPPORTED.
 *   tsasr_lstm_bwd: back-propagation through time in one cooperative launch; d_out [B,U,Hd] (+ optional d_hn, d_cn [B,Hd])
 *     -> dG [B,U,4Hd] = gradient w.r.t. the gate pre-activations (zeros at padded positions).  From it:
 *       dW_hh = dG^T hprev and db_ih = db_hh = column sums of dG: tsasr_linear_bwd(dG, hprev, NULL, B*U, Hd, 4Hd, NULL, dW_hh, db);
 *       dW_ih: tsasr_onehot_dw (one-hot input; deterministic gather-sum) or tsasr_linear_bwd against the dense input.
 *   Both launches hand a step's results from CTA to CTA through `out` / `dG` themselves (pre-filled with a sentinel by the
 *   call, polled by the consumers): no workspace, no flags. */
int tsasr_lstm_fwd(const void* tokens, int tokens_i64, int blank, int n_embed, const float* xw, const float* W_ih, const float* W_hh,
                   const float* b_ih, const float* b_hh, const float* rel_lengths, const int32_t* abs_lengths, int B, int U, int Hd,
                   float* out, float* hprev, float* gates, float* cells, float* h_n, float* c_n, int32_t* lengths_out,
                   tsasr_stream_t stream);
int tsasr_lstm_bwd(const float* d_out, const float* d_hn, const float* d_cn, const float* W_hh, const float* gates, const float* cells,
                   const int32_t* lengths, int B, int U, int Hd, float* dG, tsasr_stream_t stream);
int tsasr_onehot_dw(const void* tokens, int tokens_i64, int blank, int n_embed, const float* dG, int n_pos, int G, float* dW_ih,
                    tsasr_stream_t stream);

/* ---- the prediction network ---------------------------------------------------------
 * Replaces speechbrain.nnet.embedding.Embedding(consider_as_one_hot=True) (SB/nnet/embedding.py:65-114) followed by
 * speechbrain.nnet.RNN.LSTM (one layer, unidirectional; SB/nnet/RNN.py:170-278, torch.nn.LSTM on a PackedSequence) as
 * chained at train_librispeechmix_scratch.py:125-126 (hparams/LibriSpeechMix/conformer-t_scratch.yaml:176-185).
 * fp32 arithmetic throughout (gate order i, f, g, o; h_0 = c_0 = 0).
 *   tsasr_lstm_fwd: the whole teacher-forced recurrence in one cooperative launch.
 *     input: EITHER tokens [B,U] (int32, or int64 when tokens_i64 != 0) + W_ih [4Hd, n_embed]: the one-hot embedding of
 *       token k is column k - [k > blank] of W_ih (nothing for k == blank), gathered -- no [B,U,V-1] tensor, no GEMM;
 *       OR xw [B,U,4Hd] = x W_ih^T + b_ih precomputed (tsasr_linear_fwd) for a dense input.
 *     lengths: rel_lengths fp32 (SpeechBrain relative; converted as SB/nnet/RNN.py:35 + pack_padded_sequence do: fp32
 *       product, truncation) or abs_lengths int32; positions u >= length read as zeros in `out` and freeze the state.
 *     out [B,U,Hd]; optional (NULL to skip): hprev [B,U,Hd] = h_{u-1}, gates [B,U,4,Hd] (activated), cells [B,U,Hd]
 *       (what tsasr_lstm_bwd needs), h_n / c_n [B,Hd], lengths_out [B] int32.
 *     Hd in {128, 256, 512}, B <= 64; else TSASR_E_UNSUPPORTED.
 *   tsasr_lstm_bwd: back-propagation through time in one cooperative launch; d_out [B,U,Hd] (+ optional d_hn, d_cn [B,Hd])
 *     -> dG [B,U,4Hd] = gradient w.r.t. the gate pre-activations (zeros at padded positions).  From it:
 *       dW_hh = dG^T hprev and db_ih = db_hh = column sums of dG: tsasr_linear_bwd(dG, hprev, NULL, B*U, Hd, 4Hd, NULL, dW_hh, db);
 *       dW_ih: tsasr_onehot_dw (one-hot input; deterministic gather-sum) or tsasr_linear_bwd against the dense input.
 *   Both launches hand a step's results from CTA to CTA through `out` / `dG` themselves (pre-filled with a sentinel by the
 *   call, polled by the consumers): no workspace, no flags. */
int tsasr_lstm_fwd(const void* tokens, int tokens_i64, int blank, int n_embed, const float* xw, const float* W_ih, const float* W_hh,
                   const float* b_ih, const float* b_hh, const float* rel_lengths, const int32_t* abs_lengths, int B, int U, int Hd,
                   float* out, float* hprev, float* gates, float* cells, float* h_n, float* c_n, int32_t* lengths_out,
                   tsasr_stream_t stream);
int tsasr_lstm_bwd(const float* d_out, const float* d_hn, const float* d_cn, const float* W_hh, const float* gates, const float* cells,
                   const int32_t* lengths, int B, int U, int Hd, float* dG, tsasr_stream_t stream);
int tsasr_onehot_dw(const void* tokens, int tokens_i64, int blank, int n_embed, const float* dG, int n_pos, int G, float* dW_ih,
                    tsasr_stream_t stream);

/* ---- decode-time joint step (greedy / beam search) ------------------------------------------------
 * Replaces TransducerBeamSearcher._joint_forward_step (SB/decoders/transducer.py:375-384): Transducer_joint
 * on [B,1,1,H] inputs + the classifier Linear + LogSoftmax, two small launches instead of five.
 *   enc_t fp32 [B,H] (frame t of every hypothesis), dec fp32 [B,H]: rows of H contiguous floats, *_row_stride
 *   elements apart (a strided view of the encoder output needs no copy; 0 broadcasts one row over B);
 *   W fp32 [V,H], bias fp32 [V] or NULL;  out: log_probs fp32 [B,V].  fp32 arithmetic on the same operands as
 *   the eager path.
 * The workspace (tsasr_joint_decode_workspace_bytes(V) bytes, 16-byte aligned) is scratch for per-CTA softmax
 * partials and may be reused by later calls on the same stream. */
size_t tsasr_joint_decode_workspace_bytes(int V);
int tsasr_joint_decode_step(const float* enc_t, const float* dec, long long enc_row_stride, long long dec_row_stride,
                            const float* W, const float* bias, int B, int H, int V, int act_kind, float act_param,
                            float* log_probs, void* workspace, size_t workspace_bytes, tsasr_stream_t stream);

/* Test-only: dump the logits tile-wise recomputed by the tcgen05 mainloop into a dense fp32
 * [B,T,U,V] buffer (small shapes), so the GEMM can be checked in isolation. */
int tsasr_joint_debug_logits(const void* enc, const void* dec, const void* W, const float* bias, int B, int T, int U,
                             int H, int V, int act_kind, float act_param, float* logits_out, tsasr_stream_t stream);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
long long tsasr_launch_count(void);

/* Measurement aid (bench.py's per-kernel roofline lines): while enabled, every kernel launch of this
 * library is bracketed by two CUDA events on the launching stream (no host synchronisation).
 * tsasr_kernel_timings() waits for the recorded events, sums them per kernel name and clears the record:
 * names is max_n x 32 chars, ms / counts are max_n entries; returns the number of distinct kernels.
 * The reference has no counterpart (it times nothing on the device). */
int tsasr_kernel_timing_enable(int on);
int tsasr_kernel_timings(char* names, float* ms, int* counts, int max_n);

#ifdef __cplusplus
}
#endif
#endif /* TSASR_B200_H */
