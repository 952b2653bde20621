"""Golden vector of the BEAM search, produced by the REFERENCE searcher class (authoring container only).

    PYTHONPATH=/root/reference/vendor/speechbrain python -m oracle.make_golden_beam

Rebuilds the toy model of tests/golden/greedy_decode.npz (weights, encoder output) and runs
``speechbrain.decoders.transducer.TransducerBeamSearcher(beam_size=4, nbest=3, state_beam=2.3, expand_beam=2.3)``
(``transducer_beam_search_decode``, SB/decoders/transducer.py:220-373) with the reference's own ``Transducer_joint`` and
``Linear`` on CPU; stores the n-best label sequences and scores of every utterance in tests/golden/beam_decode.npz."""
import os

import numpy as np
import torch

from oracle.greedy_decode import ToyPredictor
from oracle.make_golden import _import_reference, OUT


def main():
    _, Transducer_joint, Linear = _import_reference()
    from speechbrain.decoders.transducer import TransducerBeamSearcher

    g = dict(np.load(os.path.join(OUT, "greedy_decode.npz")))
    B, T, V, E, HID, H = (int(x) for x in g["dims"])
    pred = ToyPredictor(V, E, HID, H)
    pred.load_state_dict({k[5:]: torch.tensor(g[k]) for k in g if k.startswith("pred.")})
    head = Linear(input_shape=(1, 1, 1, H), n_neurons=V)
    with torch.no_grad():
        # half the head gain of the greedy vector: with the peaky head the reference's expansion loop (it has no limit on
        # symbols per frame) does not terminate on this random model
        head.w.weight.copy_(0.5 * torch.tensor(g["W"]))
        head.w.bias.copy_(torch.tensor(g["b"]))
    tjoint = Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    tn = torch.tensor(g["tn"])
    cfg = dict(beam_size=4, nbest=3, state_beam=2.3, expand_beam=2.3)
    searcher = TransducerBeamSearcher(decode_network_lst=pred.layers(), tjoint=tjoint, classifier_network=[head], blank_id=0, **cfg)
    best, score, nbest, nbest_scores = searcher(tn)
    flat, lens, scores, counts = [], [], [], []
    for utt, sc in zip(nbest, nbest_scores):
        counts.append(len(utt))  # an utterance may end with fewer than nbest hypotheses in its beam
        for hyp, s in zip(utt, sc):
            lens.append(len(hyp))
            flat.extend(hyp)
            scores.append(float(s))
    np.savez_compressed(os.path.join(OUT, "beam_decode.npz"), cfg=np.array([cfg["beam_size"], cfg["nbest"]]),
                        beams=np.array([cfg["state_beam"], cfg["expand_beam"]]), hyp_counts=np.array(counts), hyp_lens=np.array(lens), hyp_flat=np.array(flat, dtype=np.int64),
                        scores=np.array(scores, dtype=np.float64), mean_exp_score=float(score), W=head.w.weight.detach().numpy(),
                        b=head.w.bias.detach().numpy())
    print("best:", best, "score:", float(score))


if __name__ == "__main__":
    main()
