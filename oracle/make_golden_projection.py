"""Authoring-container script: golden vectors of the projection module from the REFERENCE's own class
(speechbrain.nnet.linear.Linear of /root/reference, CPU) -> tests/golden/linear_proj*.npz.

    python oracle/make_golden_projection.py

Two cases: a 3-D input as encoder_proj / decoder_proj see it (train_librispeechmix_scratch.py:122,127) and a 4-D input with
combine_dims=True (SB/nnet/linear.py:71-72).  Each file holds the input, the weights and the reference's output and
gradients for loss = sum(out * d_out)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden_predictor import import_reference  # noqa: E402  (stubs hyperpyyaml / ruamel, puts the reference on sys.path)


def main():
    import_reference()
    from speechbrain.nnet.linear import Linear

    g = torch.Generator().manual_seed(5)
    torch.manual_seed(5)
    cases = {
        "linear_proj": (Linear(input_size=24, n_neurons=40), torch.randn(3, 17, 24, generator=g)),
        "linear_proj_combine": (Linear(input_shape=[None, None, 6, 4], n_neurons=9, combine_dims=True), torch.randn(2, 5, 6, 4, generator=g)),
        "linear_proj_nobias": (Linear(input_size=10, n_neurons=7, bias=False), torch.randn(4, 3, 10, generator=g)),
    }
    for name, (lin, x) in cases.items():
        x = x.requires_grad_()
        y = lin(x)
        d_out = torch.randn(y.shape, generator=g)
        (y * d_out).sum().backward()
        arrays = {"x": x.detach().numpy(), "weight": lin.w.weight.detach().numpy(), "d_out": d_out.numpy(), "out": y.detach().numpy(),
                  "d_x": x.grad.numpy(), "d_weight": lin.w.weight.grad.numpy(), "combine_dims": np.int64(lin.combine_dims)}
        if lin.w.bias is not None:
            arrays["bias"] = lin.w.bias.detach().numpy()
            arrays["d_bias"] = lin.w.bias.grad.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **arrays)
        print(name, tuple(y.shape))


if __name__ == "__main__":
    main()
