"""Regenerate tests/golden/*.npz from the REFERENCE ITSELF (authoring container only).

    PYTHONPATH=/root/reference/vendor/speechbrain python -m oracle.make_golden

Imports the vendored SpeechBrain from /root/reference (with hyperpyyaml / ruamel.yaml stubbed --
they are missing in the image and unused by the path) and runs, on CPU:

  * speechbrain.nnet.losses.transducer_loss(use_torchaudio=True)   -- what the recipe executes
    (hparams/LibriSpeechMix/conformer-t_scratch.yaml:262-264), relative lengths included, so the
    integer length conversion of losses.py:58-59 is part of the vector;
  * speechbrain.nnet.losses.transducer_loss(use_torchaudio=False)  -- the reference's own Numba
    kernels under NUMBA_ENABLE_CUDASIM=1 (tiny shapes; the simulator is very slow);
  * Transducer_joint("sum", act) -> speechbrain.nnet.linear.Linear -> transducer_loss -> backward
    (train_librispeechmix_scratch.py:132,135,158) for the fused-path vectors.

The GPU box has no /root/reference; tests only read the committed .npz files.
"""
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.path.insert(0, "/root/reference/vendor/speechbrain")
    hp = types.ModuleType("hyperpyyaml")
    hp.resolve_references = lambda *a, **k: None
    hp.load_hyperpyyaml = lambda *a, **k: {}
    sys.modules.setdefault("hyperpyyaml", hp)
    ru, ruy = types.ModuleType("ruamel"), types.ModuleType("ruamel.yaml")
    ru.yaml = ruy
    sys.modules.setdefault("ruamel", ru)
    sys.modules.setdefault("ruamel.yaml", ruy)
    from speechbrain.nnet.losses import transducer_loss
    from speechbrain.nnet.transducer.transducer_joint import Transducer_joint
    from speechbrain.nnet.linear import Linear

    return transducer_loss, Transducer_joint, Linear


def _rel(lengths, dim):
    """Relative lengths as the SpeechBrain dataloader produces them (PaddedBatch: len / max_len)."""
    return (torch.tensor(lengths, dtype=torch.float32) / float(dim)).clamp(max=1.0)


def loss_case(transducer_loss, name, logits, targets, in_lens, tg_lens, blank, use_torchaudio, reduction="mean",
              rel=None):
    """One golden vector through the reference's transducer_loss; stores loss + d loss/d logits."""
    B, T, U, V = logits.shape
    logits = logits.clone().float().requires_grad_()
    in_rel = _rel(in_lens, T) if rel is None else torch.tensor(rel[0], dtype=torch.float32)
    tg_rel = _rel(tg_lens, max(U - 1, 1)) if rel is None else torch.tensor(rel[1], dtype=torch.float32)
    tg = targets.int() if not use_torchaudio else targets.long()  # Numba branch does not cast (losses.py:85)
    loss = transducer_loss(logits, tg, in_rel, tg_rel, blank_index=blank, reduction=reduction,
                           use_torchaudio=use_torchaudio)
    (loss.sum() if loss.dim() else loss).backward()
    in_abs = (in_rel * T).round().int()
    tg_abs = (tg_rel * targets.shape[1]).round().int()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        logits=logits.detach().numpy(), targets=targets.numpy().astype(np.int32),
        input_rel=in_rel.numpy(), target_rel=tg_rel.numpy(),
        input_abs=in_abs.numpy(), target_abs=tg_abs.numpy(),
        blank=np.int32(blank), use_torchaudio=np.bool_(use_torchaudio), reduction=np.array(reduction),
        loss=loss.detach().numpy(), dlogits=logits.grad.numpy(),
    )
    print(f"{name}: loss={loss.detach().numpy()}")


def joint_case(fns, name, B, T, U, H, V, in_lens, tg_lens, act_name, act_cls, seed, reduction="mean"):
    transducer_loss, Transducer_joint, Linear = fns
    g = torch.Generator().manual_seed(seed)
    # operands are bf16-representable so the fused kernels see identical inputs
    enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().float().requires_grad_()
    dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().float().requires_grad_()
    head = Linear(n_neurons=V, input_size=H)
    with torch.no_grad():
        head.w.weight.copy_(head.w.weight.bfloat16().float())
    joiner = Transducer_joint(joint="sum", nonlinearity=act_cls)
    targets = torch.randint(1, V, (B, U - 1), generator=g)
    joint = joiner(enc[..., None, :], dec[:, None, ...])  # train_librispeechmix_scratch.py:132
    logits = head(joint)  # :135
    logits.retain_grad()
    in_rel, tg_rel = _rel(in_lens, T), _rel(tg_lens, U - 1)
    loss = transducer_loss(logits, targets, in_rel, tg_rel, blank_index=0, reduction=reduction, use_torchaudio=True)
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        enc=enc.detach().numpy(), dec=dec.detach().numpy(), W=head.w.weight.detach().numpy(),
        b=head.w.bias.detach().numpy(), targets=targets.numpy().astype(np.int32),
        input_rel=in_rel.numpy(), target_rel=tg_rel.numpy(),
        input_abs=(in_rel * T).round().int().numpy(), target_abs=(tg_rel * (U - 1)).round().int().numpy(),
        act=np.array(act_name), reduction=np.array(reduction), loss=loss.detach().numpy(),
        d_enc=enc.grad.numpy(), d_dec=dec.grad.numpy(), dW=head.w.weight.grad.numpy(), db=head.w.bias.grad.numpy(),
    )
    print(f"{name}: loss={loss.item():.6f}")


def main():
    os.makedirs(OUT, exist_ok=True)
    fns = _import_reference()
    tl = fns[0]

    # (1) the reference's only known-answer tensor, vendor/speechbrain/tests/unittests/test_losses.py:119-142
    ka = torch.tensor([[[[.1, .6, .1, .1, .1], [.1, .1, .6, .1, .1], [.1, .1, .2, .8, .1]],
                        [[.1, .6, .1, .1, .1], [.1, .1, .2, .1, .1], [.7, .1, .2, .1, .1]]]]).log_softmax(-1)
    tg = torch.tensor([[1, 2]])
    loss_case(tl, "known_answer_numba", ka, tg, [2], [2], 0, use_torchaudio=False)  # -> 2.2478
    loss_case(tl, "known_answer_torchaudio", ka, tg, [2], [2], 0, use_torchaudio=True)  # -> 4.4957

    g = torch.Generator().manual_seed(1234)
    # (2) ragged random, both implementations, all reductions of the Numba path (B=2 shows the
    #     un-normalised "mean" gradient of transducer_loss.py:280-293)
    lg = 2.0 * torch.randn(2, 12, 5, 9, generator=g)
    tg = torch.randint(1, 9, (2, 4), generator=g)
    for red in ("mean", "sum", "none"):
        loss_case(tl, f"ragged_numba_{red}", lg, tg, [12, 9], [4, 2], 0, use_torchaudio=False, reduction=red)
        loss_case(tl, f"ragged_torchaudio_{red}", lg, tg, [12, 9], [4, 2], 0, use_torchaudio=True, reduction=red)
    # (3) blank != 0 and repeated labels
    lg = 1.5 * torch.randn(2, 7, 4, 6, generator=g)
    tg = torch.tensor([[2, 2, 2], [0, 1, 0]])
    loss_case(tl, "blank5_repeat_numba", lg, tg, [7, 5], [3, 3], 5, use_torchaudio=False)
    loss_case(tl, "blank5_repeat_torchaudio", lg, tg, [7, 5], [3, 3], 5, use_torchaudio=True)
    # (4) edge rectangles: empty target (U_b = 1), single frame (T_b = 1), torchaudio route
    lg = torch.randn(3, 6, 5, 8, generator=g)
    tg = torch.randint(1, 8, (3, 4), generator=g)
    loss_case(tl, "edges_torchaudio", lg, tg, [6, 1, 4], [4, 4, 0], 0, use_torchaudio=True, reduction="none")
    lg = torch.randn(2, 5, 1, 8, generator=g)  # U = 1: no labels at all
    loss_case(tl, "no_labels_torchaudio", lg, torch.zeros(2, 0, dtype=torch.long), [5, 3], [0, 0], 0,
              use_torchaudio=True, reduction="none", rel=([1.0, 0.6], [1.0, 1.0]))
    # (5) large-magnitude logits (softmax saturation) and a wider vocabulary
    lg = 30.0 * torch.randn(2, 9, 6, 40, generator=g)
    tg = torch.randint(1, 40, (2, 5), generator=g)
    loss_case(tl, "large_magnitude_torchaudio", lg, tg, [9, 8], [5, 3], 0, use_torchaudio=True, reduction="none")
    # (6) relative lengths that hit .5 exactly: (rel*dim).round() is round-half-to-even (losses.py:58-59)
    lg = torch.randn(4, 8, 5, 7, generator=g)
    tg = torch.randint(1, 7, (4, 4), generator=g)
    loss_case(tl, "half_rounding_torchaudio", lg, tg, None, None, 0, use_torchaudio=True, reduction="none",
              rel=([1.0, 0.3125, 0.5625, 0.8125], [1.0, 0.625, 0.375, 0.125]))  # 2.5->2, 4.5->4, 6.5->6 / 2.5->2, 1.5->2, 0.5->0
    # (7) mid-size, V=1000-like row width kept small enough to commit (torchaudio route)
    lg = torch.randn(2, 20, 8, 120, generator=g)
    tg = torch.randint(1, 120, (2, 7), generator=g)
    loss_case(tl, "mid_torchaudio", lg, tg, [20, 13], [7, 4], 0, use_torchaudio=True, reduction="mean")

    # (8) fused-path vectors through the reference's Transducer_joint + Linear + transducer_loss
    joint_case(fns, "joint_leaky", 2, 24, 9, 64, 40, [24, 17], [8, 5], "leaky_relu", torch.nn.LeakyReLU, 7)
    joint_case(fns, "joint_tanh", 2, 16, 6, 64, 33, [16, 11], [5, 5], "tanh", torch.nn.Tanh, 8, reduction="sum")
    joint_case(fns, "joint_relu", 1, 10, 4, 128, 29, [10], [3], "relu", torch.nn.ReLU, 9)


if __name__ == "__main__":
    main()
