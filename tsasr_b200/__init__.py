"""tsasr_b200 -- B200-native (sm_100a) joint network + RNN-T loss hot path of lucadellalib/ts-asr.

Drop-in surface (same names/signatures as the vendored SpeechBrain pieces the recipe uses):

    tsasr_b200.Transducer_joint          <- speechbrain.nnet.transducer.transducer_joint.Transducer_joint
    tsasr_b200.transducer_loss           <- speechbrain.nnet.losses.transducer_loss
    tsasr_b200.TransducerLoss/Transducer <- speechbrain.nnet.loss.transducer_loss.{TransducerLoss,Transducer}
    tsasr_b200.Linear                    <- speechbrain.nnet.linear.Linear (encoder_proj / decoder_proj: tcgen05 GEMMs whose
                                            bf16 output copy feeds the fused joint directly)
    tsasr_b200.Embedding / tsasr_b200.LSTM <- speechbrain.nnet.embedding.Embedding / speechbrain.nnet.RNN.LSTM (the prediction
                                            network: one-hot gather + the whole recurrence in one cooperative launch)

plus the functional forms ``rnnt_loss`` (torchaudio signature) and ``fused_joint_rnnt_loss``.
All arithmetic runs in libtsasr_b200.so (hand-written CUDA, C ABI in include/tsasr_b200.h); there
is no CPU path and no PyTorch fallback for the loss.
"""
import torch as _torch

from . import _lib, linear, monitor, ops, predictor  # noqa: F401
from .functional import fused_joint_rnnt_loss, rnnt_loss  # noqa: F401
from .linear import Linear  # noqa: F401
from .predictor import Embedding, LSTM, OneHotHandle  # noqa: F401
from .losses import Transducer, TransducerLoss, transducer_loss  # noqa: F401
from .transducer_joint import JointHandle, Transducer_joint  # noqa: F401

__version__ = "0.1.0"


def install_into_speechbrain():
    """Monkey-patch an importable ``speechbrain`` so unmodified recipes pick up the drop-ins
    (equivalent to editing the three yaml tags)."""
    import speechbrain.nnet.losses as sb_losses
    import speechbrain.nnet.transducer.transducer_joint as sb_joint

    sb_losses.transducer_loss = transducer_loss
    sb_joint.Transducer_joint = Transducer_joint
    try:
        import speechbrain.nnet.loss.transducer_loss as sb_tl

        sb_tl.TransducerLoss = TransducerLoss
        sb_tl.Transducer = Transducer
    except ImportError:  # numba missing: the reference module cannot even be imported
        pass


def adopt_modules(modules, names=("encoder_proj", "decoder_proj", "embedding", "decoder")):
    """Swap already-constructed SpeechBrain modules for the drop-ins IN PLACE of a ``modules`` mapping (the dict handed to
    ``Brain(modules=...)``, or a ``torch.nn.ModuleDict``), sharing their parameter tensors: the drop-in ``Linear`` takes the
    stock module's ``w`` (nn.Linear), ``Embedding`` its ``Embedding`` (nn.Embedding), ``LSTM`` its ``rnn`` (nn.LSTM) -- no
    copy, so optimizer state, checkpoints and DDP buckets built afterwards see the same parameters under the same names.
    The equivalent of editing the yaml tags of INTEGRATION.md for code that builds its modules itself; call it BEFORE the
    Brain wraps the modules in DistributedDataParallel.  Modules of other types (or missing names) are left alone.
    Returns the list of names that were swapped."""
    swapped = []
    for name in names:
        try:
            old = modules[name]
        except (KeyError, IndexError, TypeError):
            continue
        cls = type(old).__name__
        new = None
        if cls == "Linear" and isinstance(getattr(old, "w", None), _torch.nn.Linear) and not isinstance(old, Linear):
            new = Linear.__new__(Linear)
            _torch.nn.Module.__init__(new)
            new.combine_dims = getattr(old, "combine_dims", False)
            new.w = old.w
        elif cls == "Embedding" and hasattr(old, "Embedding") and hasattr(old, "consider_as_one_hot") and not isinstance(old, Embedding):
            new = Embedding.__new__(Embedding)
            _torch.nn.Module.__init__(new)
            new.num_embeddings, new.consider_as_one_hot = old.num_embeddings, old.consider_as_one_hot
            new.embedding_dim, new.blank_id = old.embedding_dim, old.blank_id
            new.Embedding = old.Embedding
        elif cls == "LSTM" and isinstance(getattr(old, "rnn", None), _torch.nn.LSTM) and not isinstance(old, LSTM):
            new = LSTM.__new__(LSTM)
            _torch.nn.Module.__init__(new)
            new.reshape = getattr(old, "reshape", False)
            new.rnn = old.rnn
        if new is not None:
            new.train(old.training)
            modules[name] = new
            swapped.append(name)
    return swapped
