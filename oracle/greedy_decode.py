"""Transducer greedy decoding restated for the decode-step tests (TEST ORACLE ONLY).

Restates ``TransducerBeamSearcher.transducer_greedy_decode`` and its helpers
(vendor/speechbrain/speechbrain/decoders/transducer.py:138-218, 375-384, 411-502): per frame one joint step over the
batch, arg-max, and a prediction-network step for the hypotheses that emitted a non-blank.  The joint step is a
parameter, so the same loop runs with the reference's eager chain (``eager_joint_step``) or with the fused
``tsasr_b200.decode`` kernel.  ``ToyPredictor`` is the prediction network used by the golden vectors
(oracle/make_golden_decode.py runs the REAL searcher class on the same modules).
"""
import torch


class LSTM(torch.nn.Module):
    """Recurrent layer with the call signature the searcher expects; the class NAME matters: ``_forward_PN``
    dispatches on ``layer.__class__.__name__`` (transducer.py:491-499)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.rnn = torch.nn.LSTM(input_size, hidden_size, batch_first=True)

    def forward(self, x, hx=None):
        return self.rnn(x, hx)


class ToyPredictor(torch.nn.Module):
    """Embedding -> LSTM -> Linear, i.e. ``decode_network_lst = [emb, dec, dec_lin]`` of the recipe
    (train_librispeechmix_scratch.py:125-127) with plain torch modules."""

    def __init__(self, vocab, emb_dim, hidden, joint_dim):
        super().__init__()
        self.emb = torch.nn.Embedding(vocab, emb_dim)
        self.dec = LSTM(emb_dim, hidden)
        self.dec_lin = torch.nn.Linear(hidden, joint_dim)

    def layers(self):
        return [self.emb, self.dec, self.dec_lin]


def forward_pn(tokens, layers, hidden=None):
    """transducer.py:468-502"""
    out = tokens
    for layer in layers:
        if layer.__class__.__name__ in ["RNN", "LSTM", "GRU", "LiGRU", "LiGRU_Layer"]:
            out, hidden = layer(out, hidden)
        else:
            out = layer(out)
    return out, hidden


def eager_joint_step(tjoint, classifier_network, softmax):
    """transducer.py:375-384"""
    def step(h_i, out_pn):
        with torch.no_grad():
            out = tjoint(h_i, out_pn)
            for layer in classifier_network:
                out = layer(out)
            return softmax(out)
    return step


def greedy_decode(tn_output, pn_layers, joint_step, blank_id=0):
    """transducer.py:138-218 -> (predictions per utterance, summed log-prob scores per utterance)."""
    B = tn_output.size(0)
    predictions = [[] for _ in range(B)]
    scores = [0.0 for _ in range(B)]
    hidden = None
    input_pn = torch.ones((B, 1), device=tn_output.device, dtype=torch.int32) * blank_id
    with torch.no_grad():
        out_pn, hidden = forward_pn(input_pn, pn_layers)
        for t in range(tn_output.size(1)):
            log_probs = joint_step(tn_output[:, t, :].unsqueeze(1).unsqueeze(1), out_pn.unsqueeze(1))
            logp, pos = torch.max(log_probs.squeeze(1).squeeze(1), dim=1)
            updated = []
            for i in range(B):
                if pos[i].item() != blank_id:
                    predictions[i].append(pos[i].item())
                    scores[i] += logp[i].item()
                    input_pn[i][0] = pos[i]
                    updated.append(i)
            if updated:
                sel_in = input_pn[updated, :]
                sel_hidden = (hidden[0][:, updated, :], hidden[1][:, updated, :]) if isinstance(hidden, tuple) \
                    else hidden[:, updated, :]                                       # transducer.py:411-441
                sel_out, sel_hidden = forward_pn(sel_in, pn_layers, sel_hidden)
                out_pn[updated] = sel_out
                if isinstance(hidden, tuple):                                        # transducer.py:443-466
                    hidden[0][:, updated, :] = sel_hidden[0]
                    hidden[1][:, updated, :] = sel_hidden[1]
                else:
                    hidden[:, updated, :] = sel_hidden
    return predictions, scores
