"""Authoring-container script: golden vectors of the prediction network from the REFERENCE's own modules
(speechbrain.nnet.embedding.Embedding + speechbrain.nnet.RNN.LSTM of /root/reference, CPU) -> tests/golden/predictor_*.npz.

    python oracle/make_golden_predictor.py

Cases: the recipe's shape class (one-hot input, relative lengths incl. a product that truncates below its rounded value: 0.6 * 13 = 7.8 -> 7),
blank != 0, a dense input, no lengths.  Each file holds the inputs, the initial weights and the reference's outputs and
gradients for loss = sum(out * d_out)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/vendor/speechbrain"


def import_reference():
    sys.path.insert(0, REF)
    for name in ("hyperpyyaml",):
        m = types.ModuleType(name)
        m.resolve_references = lambda *a, **k: None
        m.load_hyperpyyaml = lambda *a, **k: {}
        sys.modules[name] = m
    ru, ruy = types.ModuleType("ruamel"), types.ModuleType("ruamel.yaml")
    ru.yaml = ruy
    sys.modules["ruamel"], sys.modules["ruamel.yaml"] = ru, ruy
    from speechbrain.nnet.embedding import Embedding
    from speechbrain.nnet.RNN import LSTM

    return Embedding, LSTM


def main():
    Embedding, LSTM = import_reference()
    cases = {
        # name: (B, U, V, Hd, blank, relative lengths or None, dense input size or None)
        "predictor_onehot_ragged": (5, 13, 40, 128, 0, [1.0, 0.6, 0.3, 0.55, 0.09], None),
        "predictor_onehot_blank3": (3, 9, 17, 128, 3, [1.0, 0.6, 0.35], None),
        "predictor_onehot_full": (2, 6, 12, 128, 0, None, None),
        "predictor_dense": (4, 11, 0, 128, 0, [0.8, 1.0, 0.46, 0.2], 24),
    }
    for seed, (name, (B, U, V, Hd, blank, rel, dense_in)) in enumerate(cases.items()):
        g = torch.Generator().manual_seed(100 + seed)
        torch.manual_seed(11)
        if dense_in is None:
            emb = Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=blank)
            lstm = LSTM(input_shape=[None, None, V - 1], hidden_size=Hd, num_layers=1)
            tokens = torch.randint(0, V, (B, U), generator=g)
            tokens[:, 0] = blank  # <bos> = blank, as tokens_bos in the recipe
            x = emb(tokens)
        else:
            lstm = LSTM(input_size=dense_in, hidden_size=Hd, num_layers=1)
            tokens = torch.zeros((B, U), dtype=torch.long)
            x = torch.randn(B, U, dense_in, generator=g, requires_grad=True)
        rel_t = torch.tensor(rel, dtype=torch.float32) if rel is not None else None
        out, (h_n, c_n) = lstm(x, lengths=rel_t)
        d_out = torch.randn(out.shape, generator=g)
        (out * d_out).sum().backward()
        r = lstm.rnn
        arrays = {
            "tokens": tokens.numpy(), "vocab": np.int64(V), "blank": np.int64(blank), "hidden": np.int64(Hd),
            "rel_lengths": rel_t.numpy() if rel_t is not None else np.zeros((0,), np.float32),
            "weight_ih": r.weight_ih_l0.detach().numpy(), "weight_hh": r.weight_hh_l0.detach().numpy(),
            "bias_ih": r.bias_ih_l0.detach().numpy(), "bias_hh": r.bias_hh_l0.detach().numpy(),
            "d_out": d_out.numpy(), "out": out.detach().numpy(), "h_n": h_n.detach().numpy()[0], "c_n": c_n.detach().numpy()[0],
            "d_weight_ih": r.weight_ih_l0.grad.numpy(), "d_weight_hh": r.weight_hh_l0.grad.numpy(),
            "d_bias_ih": r.bias_ih_l0.grad.numpy(), "d_bias_hh": r.bias_hh_l0.grad.numpy(),
        }
        if dense_in is not None:
            arrays["x"] = x.detach().numpy()
            arrays["d_x"] = x.grad.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **arrays)
        print(name, "out", tuple(out.shape), "lengths(ref packing)", (rel_t * U).to(torch.int64).tolist() if rel_t is not None else None)


if __name__ == "__main__":
    main()
