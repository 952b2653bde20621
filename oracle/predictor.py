"""CPU restatement of the prediction network (TEST ORACLE ONLY).

* ``onehot_embedding`` -- speechbrain.nnet.embedding.Embedding with consider_as_one_hot=True
  (vendor/speechbrain/speechbrain/nnet/embedding.py:76-103,105-114): a frozen eye matrix whose blank row is zero, i.e.
  token k -> e_{k - [k > blank]} in R^{V-1}, blank -> 0.
* ``packed_lengths`` -- speechbrain.nnet.RNN.pack_padded_sequence (vendor/speechbrain/speechbrain/nnet/RNN.py:25-38):
  ``(lengths * inputs.size(1)).cpu()`` stays a FLOAT tensor; torch.nn.utils.rnn.pack_padded_sequence then casts it to
  int64, i.e. truncates (not the round() the loss applies to the same relative lengths at SB/nnet/losses.py:58-59).
* ``lstm_forward`` -- speechbrain.nnet.RNN.LSTM.forward (:254-278) = torch.nn.LSTM (one layer, batch_first) over the
  packed sequence followed by pad_packed_sequence: gates i,f,g,o = W_ih x + b_ih + W_hh h + b_hh; c' = f c + i g;
  h' = o tanh(c'); positions past an utterance's length are zeros in the output and leave its state frozen.

Written with explicit per-step loops in float64 torch so that autograd gives the exact gradients (no cuDNN anywhere);
pinned against the reference's own modules by tests/golden/predictor_*.npz (oracle/make_golden_predictor.py).
Nothing in tsasr_b200/ imports this.
"""
import torch


def onehot_embedding(tokens, vocab, blank=0, dtype=torch.float64):
    """[B,U] ints -> [B,U,V-1]."""
    tokens = tokens.long()
    col = torch.where(tokens > blank, tokens - 1, tokens)
    out = torch.zeros(*tokens.shape, vocab - 1, dtype=dtype)
    keep = tokens != blank
    out[keep, col[keep]] = 1.0
    return out


def packed_lengths(rel_lengths, U):
    """Absolute lengths as the reference's packing sees them: fp32 product, then truncation (int64 cast)."""
    return (rel_lengths.float() * U).to(torch.int64)


def lstm_forward(x, w_ih, w_hh, b_ih, b_hh, lengths):
    """x [B,U,In] float64, lengths [B] absolute ints -> (out [B,U,Hd] zeros past each length, h_n [B,Hd], c_n [B,Hd])."""
    B, U, _ = x.shape
    Hd = w_hh.shape[1]
    h = x.new_zeros(B, Hd)
    c = x.new_zeros(B, Hd)
    outs = []
    for u in range(U):
        gates = x[:, u] @ w_ih.T + h @ w_hh.T
        if b_ih is not None:
            gates = gates + b_ih + b_hh
        i, f, g, o = gates.split(Hd, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        valid = (u < lengths).to(x.dtype).unsqueeze(1)
        c = valid * c_new + (1 - valid) * c
        h = valid * h_new + (1 - valid) * h
        outs.append(valid * h_new)
    return torch.stack(outs, dim=1), h, c


def predictor_fwd_bwd(tokens, vocab, blank, params, rel_lengths, d_out, x_dense=None):
    """One-hot (tokens) or dense (x_dense) input through the LSTM, loss = sum(out * d_out); float64.
    params: dict weight_ih, weight_hh, bias_ih, bias_hh (torch.nn.LSTM's names without the _l0 suffix).
    -> dict out, h_n, c_n, lengths, d_weight_ih, d_weight_hh, d_bias_ih, d_bias_hh (+ d_x for a dense input)."""
    p = {k: (v.detach().double().clone().requires_grad_() if v is not None else None) for k, v in params.items()}
    if x_dense is None:
        x = onehot_embedding(tokens, vocab, blank)
        U = tokens.shape[1]
    else:
        x = x_dense.detach().double().clone().requires_grad_()
        U = x.shape[1]
    L = packed_lengths(rel_lengths, U) if rel_lengths is not None else torch.full((x.shape[0],), U, dtype=torch.int64)
    out, h_n, c_n = lstm_forward(x, p["weight_ih"], p["weight_hh"], p["bias_ih"], p["bias_hh"], L)
    (out * d_out.double()).sum().backward()
    res = {"out": out.detach(), "h_n": h_n.detach(), "c_n": c_n.detach(), "lengths": L}
    for k, v in p.items():
        if v is not None:
            res["d_" + k] = v.grad
    if x_dense is not None:
        res["d_x"] = x.grad
    return res
