// predictor.cuh -- parameter blocks and launchers of the prediction-network kernels (predictor.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace tsasr {

static constexpr int kLstmThreads = 256;
static constexpr int kLstmUnits = 4;        // hidden units per CTA (their 16 gate rows of W_hh live in registers)
static constexpr int kLstmBatchTile = 16;   // utterances per pass (two batch groups of 8 per warp row)
static constexpr int kLstmMaxPasses = 4;    // B <= 64

struct LstmFwdParams {
    const long long* tok64;     // one-hot mode (xw == nullptr): token ids [B,U], int64 or ...
    const int* tok32;           // ... int32
    int blank, n_embed;         // blank id, V - 1 = columns of W_ih
    const float* xw;            // dense mode: x W_ih^T + b_ih precomputed [B,U,4Hd]
    const float* W_ih;          // [4Hd, n_embed] (one-hot mode)
    const float* W_hh;          // [4Hd, Hd]
    const float* b_ih;          // [4Hd] or nullptr
    const float* b_hh;          // [4Hd] or nullptr
    const float* rel_lengths;   // [B] relative lengths (SpeechBrain), or nullptr ...
    const int* abs_lengths;     // ... absolute ones
    int B, U, Hd;
    int dbg;                    // development ablations (TSASR_DEBUG_LSTM): 1 = skip the dot products, 2 = skip the hand-off fetch, 4 = no probe before the bulk fetch
    float* out;                 // [B,U,Hd] h_t, zeros at padded positions; also the grid-wide hand-off buffer (sentinel-filled)
    float* hprev;               // [B,U,Hd] h_{t-1} (the X operand of dW_hh), or nullptr
    float* gates;               // [B,U,4,Hd] activated gates i,f,g,o, or nullptr
    float* cells;               // [B,U,Hd] c_t, or nullptr
    float* h_n;                 // [B,Hd] or nullptr
    float* c_n;                 // [B,Hd] or nullptr
    int* lengths_out;           // [B] absolute lengths as used, or nullptr
};

struct LstmBwdParams {
    const float* d_out;         // [B,U,Hd]
    const float* d_hn;          // [B,Hd] or nullptr
    const float* d_cn;          // [B,Hd] or nullptr
    const float* W_hh;
    const float* gates;
    const float* cells;
    const int* lengths;         // [B] absolute
    int B, U, Hd;
    int dbg;
    float* dG;                  // [B,U,4Hd] d loss / d gate pre-activations (zeros at padded positions); sentinel-filled on entry
};

size_t lstm_fwd_smem_bytes(int B, int U, int Hd, bool onehot);
cudaError_t launch_lstm_fwd(const LstmFwdParams& p, cudaStream_t st);
cudaError_t launch_lstm_bwd(const LstmBwdParams& p, cudaStream_t st);
cudaError_t launch_onehot_dw(const void* tokens, int tokens_i64, int blank, int n_embed, const float* dG, int n_pos, int G, float* dW,
                             cudaStream_t st);

}  // namespace tsasr
