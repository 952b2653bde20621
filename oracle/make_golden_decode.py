"""Golden vectors for the decode-step row, produced by the REFERENCE searcher class (authoring container only).

    PYTHONPATH=/root/reference/vendor/speechbrain python -m oracle.make_golden_decode

Runs ``speechbrain.decoders.transducer.TransducerBeamSearcher(beam_size=1).transducer_greedy_decode`` and the
reference ``Transducer_joint`` + ``Linear`` from /root/reference on CPU over a small random model, and stores the
model weights, the encoder output, the decoded hypotheses / scores and the first frames' log-probs in
tests/golden/greedy_decode.npz.  The GPU tests rebuild the same modules from the file.
"""
import os

import numpy as np
import torch

from oracle.greedy_decode import ToyPredictor
from oracle.make_golden import _import_reference, OUT


def main():
    _, Transducer_joint, Linear = _import_reference()
    from speechbrain.decoders.transducer import TransducerBeamSearcher

    torch.manual_seed(1234)
    B, T, V, E, HID, H = 3, 40, 29, 16, 24, 64
    pred = ToyPredictor(V, E, HID, H)
    head = Linear(input_shape=(1, 1, 1, H), n_neurons=V)      # the reference's own Linear wrapper (.w = nn.Linear)
    with torch.no_grad():
        head.w.weight.mul_(4.0)                               # peaky enough that non-blank labels are emitted
        head.w.bias[0] += 6.0  # blank wins on roughly half of the frames
    tjoint = Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    tn = 1.5 * torch.randn(B, T, H)
    searcher = TransducerBeamSearcher(decode_network_lst=pred.layers(), tjoint=tjoint, classifier_network=[head],
                                      blank_id=0, beam_size=1, nbest=1)
    hyps, score, _, _ = searcher(tn)
    # per-utterance summed log-probs are not returned by the reference (only exp().mean()); recompute them with the
    # reference's own step function while replaying the hypotheses
    with torch.no_grad():
        frames = [searcher._joint_forward_step(tn[:, t, :].unsqueeze(1).unsqueeze(1),
                                               pred.dec_lin(pred.dec(pred.emb(torch.zeros(B, 1, dtype=torch.int32)))[0]).unsqueeze(1))
                  for t in range(3)]
    flat = {f"pred.{k}": v.numpy() for k, v in pred.state_dict().items()}
    np.savez_compressed(
        os.path.join(OUT, "greedy_decode.npz"), tn=tn.numpy(), W=head.w.weight.detach().numpy(),
        b=head.w.bias.detach().numpy(), dims=np.array([B, T, V, E, HID, H]),
        hyp_lens=np.array([len(h) for h in hyps]), hyp_flat=np.array([x for h in hyps for x in h], dtype=np.int64),
        mean_exp_score=float(score), first_frames_logp=torch.stack(frames).squeeze(2).squeeze(2).numpy(), **flat)
    print("hyps:", hyps, "score:", float(score))


if __name__ == "__main__":
    main()
