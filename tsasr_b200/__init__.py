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
