"""In-situ harness for BASELINE.json configs[4]: one full training step of train_librispeechmix_scratch.py (causal
Conformer encoder + target-speaker embedding conditioning, injection_mode=cat) with the fused joint + RNN-T loss swapped
in, beside the same step with the stock SpeechBrain modules.

The UNMODIFIED reference is used as it is: the real ``TSASR`` Brain (train_librispeechmix_scratch.py:33-190), its real
``fit_batch`` (SB/core.py:1032-1096), the reference's own per-module DistributedDataParallel wrapping
(SB/core.py:1464-1484) and gradient accumulation (``no_sync``, SB/core.py:1585-1615).  It is found under
``$TSASR_REFERENCE_ROOT``, ``baseline/_ref`` (tools/install_reference.sh; travels to the GPU box) or /root/reference.
HyperPyYAML / ruamel are not installed in this image, so the module tree of
hparams/LibriSpeechMix/conformer-t_scratch.yaml:122-259 is mirrored in code below (values cited), with the overrides of
the ``*_SpkEmbCat_Causal`` task: causal_encoder=True, frontend_padding=causal, injection_mode=cat; V is overridden to
1000.  Data: synthetic 16 s waveforms (-> T = 401 encoder frames) and random token ids.

    python tools/insitu_step.py [--arm both|stock|dropin] [--batch 16] [--seconds 16] [--labels 99] [--steps 8] [--warmup 4]
    python -m torch.distributed.run --nproc-per-node N ... tools/insitu_step.py ...      (reference DDP, one rank per GPU)

This file is measurement / test infrastructure: nothing in tsasr_b200/ imports it.
"""
import argparse
import json
import os
import sys
import time
import types
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def find_reference():
    """(speechbrain parent dir, recipe dir) of the first usable reference install, else None."""
    cands = []
    env = os.environ.get("TSASR_REFERENCE_ROOT")
    if env:
        cands += [(os.path.join(env, "vendor", "speechbrain"), env), (env, os.path.join(env, "recipe"))]
    ref = os.path.join(ROOT, "baseline", "_ref")
    cands += [(ref, os.path.join(ref, "recipe")), ("/root/reference/vendor/speechbrain", "/root/reference")]
    for sb_parent, recipe in cands:
        if os.path.isdir(os.path.join(sb_parent, "speechbrain")) and os.path.isfile(os.path.join(recipe, "train_librispeechmix_scratch.py")):
            return sb_parent, recipe
    return None


def import_reference():
    """Imports the vendored SpeechBrain and the recipe module without HyperPyYAML / ruamel (two stub modules: the
    recipe only uses them in its __main__ block).  Returns (speechbrain, recipe module, ConformerEncoder)."""
    found = find_reference()
    if found is None:
        raise RuntimeError("reference not found: run tools/install_reference.sh (baseline/_ref) or set TSASR_REFERENCE_ROOT")
    sb_parent, recipe = found
    for p in (recipe, sb_parent):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "hyperpyyaml" not in sys.modules:
        hp = types.ModuleType("hyperpyyaml")
        hp.resolve_references = lambda *a, **k: None
        hp.load_hyperpyyaml = lambda *a, **k: {}
        sys.modules["hyperpyyaml"] = hp
    if "ruamel" not in sys.modules:
        ru, ruy = types.ModuleType("ruamel"), types.ModuleType("ruamel.yaml")
        ru.yaml = ruy
        sys.modules["ruamel"], sys.modules["ruamel.yaml"] = ru, ruy
    import speechbrain as sb
    import train_librispeechmix_scratch as rec
    from models.conformer import ConformerEncoder

    return sb, rec, ConformerEncoder


def build_modules(sb, ConformerEncoder, V, dropin, dropout=0.1, which="all"):
    """Module tree of conformer-t_scratch.yaml:122-232 (+ the causal / cat overrides of the SpkEmbCat_Causal task)."""
    from speechbrain.lobes.features import Fbank
    from speechbrain.lobes.models.convolution import ConvolutionFrontEnd
    from speechbrain.nnet.embedding import Embedding
    from speechbrain.nnet.linear import Linear
    from speechbrain.nnet.RNN import LSTM
    from speechbrain.processing.features import InputNormalization

    if dropin:
        import tsasr_b200

        joint_cls = tsasr_b200.Transducer_joint                     # yaml:191-193, tag swapped
        proj_cls = Linear
        if which in ("all", "joint+proj"):
            proj_cls = tsasr_b200.Linear                            # yaml:172-174,187-189 (N1): bf16 operand producers
        if which == "all":
            Embedding, LSTM = tsasr_b200.Embedding, tsasr_b200.LSTM  # yaml:176-185 (N3): the prediction network
    else:
        from speechbrain.nnet.transducer.transducer_joint import Transducer_joint as joint_cls
        proj_cls = Linear
    d_model, joint_dim = 256, 640

    def frontend(padding):
        return ConvolutionFrontEnd(input_shape=[None, None, 80], num_blocks=2, num_layers_per_block=1, out_channels=(128, 128),
                                   kernel_sizes=(3, 3), strides=(2, 2), residuals=(True, True), dropout=dropout, padding=padding)

    mods = {
        "feature_extractor": Fbank(sample_rate=16000, n_fft=512, n_mels=80, win_length=32),
        "normalizer": InputNormalization(norm_type="sentence", update_until_epoch=4),
        "frontend": frontend("causal"),
        "encoder": ConformerEncoder(input_size=2560, d_model=d_model, nhead=4, num_layers=12, d_ffn=2048, dropout=dropout,
                                    activation=torch.nn.LeakyReLU, kernel_size=31, causal=True, injection_mode="cat",
                                    injection_after=0),
        "encoder_proj": proj_cls(input_size=d_model, n_neurons=joint_dim),
        "embedding": Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=0),
        "decoder": LSTM(input_shape=[None, None, V - 1], hidden_size=512, num_layers=1),
        "decoder_proj": proj_cls(input_size=512, n_neurons=joint_dim),
        "joiner": joint_cls(joint="sum", nonlinearity=torch.nn.LeakyReLU),
        "transducer_head": Linear(input_size=joint_dim, n_neurons=V),   # stays the stock class in both arms
        "speaker_feature_extractor": Fbank(sample_rate=16000, n_fft=512, n_mels=80, win_length=32),
        "speaker_normalizer": InputNormalization(norm_type="sentence", update_until_epoch=4),
        "speaker_frontend": frontend("same"),
        "speaker_encoder": ConformerEncoder(input_size=2560, d_model=d_model, nhead=4, num_layers=6, d_ffn=2048, dropout=dropout,
                                            activation=torch.nn.LeakyReLU, kernel_size=31),
        "speaker_proj": Linear(input_size=d_model, n_neurons=d_model),
    }
    return mods


def build_hparams(sb, dropin, grad_accumulation_factor, max_grad_norm=5.0):
    from speechbrain.nnet.schedulers import NoamScheduler
    from speechbrain.utils.epoch_loop import EpochCounter

    if dropin:
        import tsasr_b200

        loss_fn = tsasr_b200.transducer_loss                           # yaml:262-264, tag swapped
    else:
        from speechbrain.nnet.losses import transducer_loss as loss_fn
    return {
        "epoch_counter": EpochCounter(limit=100), "injection_mode": "cat", "plot_embeddings": False, "plot_attentions": False,
        "augment": False, "valid_search_freq": 1, "transducer_loss": partial(loss_fn, use_torchaudio=True, blank_index=0),
        "enable_scheduler": True, "noam_scheduler": NoamScheduler(lr_initial=0.001, n_warmup_steps=10000),
        "grad_accumulation_factor": grad_accumulation_factor, "max_grad_norm": max_grad_norm, "nonfinite_patience": 10,
        "auto_mix_prec": False,
    }


def make_batch(sb, B, seconds, n_labels, V, seed, ragged=False):
    """Synthetic PaddedBatch with the keys the recipe's dataio pipeline produces (train_librispeechmix_scratch.py:271+)."""
    from speechbrain.dataio.batch import PaddedBatch

    g = torch.Generator().manual_seed(seed)
    n = int(16000 * seconds)
    items = []
    for i in range(B):
        ni, li = n, n_labels
        if ragged and i > 0:
            ni = int(n * (0.6 + 0.4 * torch.rand(1, generator=g).item()))
            li = max(1, int(n_labels * (0.4 + 0.6 * torch.rand(1, generator=g).item())))
        tokens = torch.randint(1, V, (li,), generator=g)
        items.append({
            "id": f"utt{seed}_{i}", "mixed_sig": 0.1 * torch.randn(ni, generator=g), "enroll_sig": 0.1 * torch.randn(int(16000 * 5), generator=g),
            "tokens_bos": torch.cat([torch.zeros(1, dtype=torch.long), tokens]), "tokens": tokens, "target_words": ["x"],
        })
    return PaddedBatch(items)


def make_brain(sb, rec, ConformerEncoder, V, dropin, device, distributed, grad_accumulation_factor, seed, dropout, deferred_check=False,
               which="all"):
    torch.manual_seed(seed)  # identical initial weights in both arms (and on every rank, as DDP would broadcast them)
    mods = build_modules(sb, ConformerEncoder, V, dropin, dropout, which)
    hparams = build_hparams(sb, dropin, grad_accumulation_factor)
    rec.hparams = hparams  # TSASR reads the module-global `hparams` for its plot_* switches (:65,98)
    run_opts = {"device": str(device)}
    if distributed:
        run_opts.update({"distributed_launch": True, "distributed_backend": "nccl", "find_unused_parameters": False})
    brain = rec.TSASR(modules=mods, opt_class=partial(torch.optim.AdamW, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01),
                      hparams=hparams, run_opts=run_opts, checkpointer=None)
    brain.on_fit_start()             # _compile, _wrap_distributed (per-module DDP), init_optimizers
    brain.modules.train()
    brain.grad_norm_epoch, brain.nonfinite_count = [], 0   # what _fit_train sets up (SB/core.py:1181-1186)
    if not hasattr(brain, "valid_step"):
        brain.valid_step = 0
    if deferred_check:  # N4: the loss.isfinite() host sync between forward and backward (SB/core.py:1072,1130) deferred by one step
        import tsasr_b200

        tsasr_b200.monitor.install(brain)
    return brain


def head_linear(brain):
    m = brain.modules.transducer_head
    m = getattr(m, "module", m)  # DDP wrapper
    return m.w


def parity_cycle(brain, batches, grad_accumulation_factor, seed):
    """One full accumulation cycle of fit_batch; the gradients are captured right before the optimizer step (after the
    head's DDP reducer has averaged them when distributed).  -> (losses, {name: grad})"""
    captured = {}
    orig_step = brain.optimizer.step

    def capturing_step(*a, **k):
        if not captured:
            w = head_linear(brain)
            captured["head_grad"] = w.weight.grad.detach().float().cpu().clone()
            captured["head_bias_grad"] = w.bias.grad.detach().float().cpu().clone()
            enc_proj = getattr(brain.modules.encoder_proj, "module", brain.modules.encoder_proj)
            captured["enc_proj_grad"] = enc_proj.w.weight.grad.detach().float().cpu().clone()
            dec = getattr(brain.modules.decoder, "module", brain.modules.decoder)
            captured["dec_rnn_hh_grad"] = dec.rnn.weight_hh_l0.grad.detach().float().cpu().clone()
            captured["dec_rnn_ih_grad"] = dec.rnn.weight_ih_l0.grad.detach().float().cpu().clone()
        return orig_step(*a, **k)

    brain.optimizer.step = capturing_step
    losses = []
    for i in range(grad_accumulation_factor):
        torch.manual_seed(seed + 17 * i)  # same dropout masks in both arms
        losses.append(float(brain.fit_batch(batches[i % len(batches)])))
    brain.optimizer.step = orig_step
    return losses, captured


def run_arm(sb, rec, ConformerEncoder, args, dropin, device, world, deferred_check=False):
    """-> dict(ms_per_step, peak_mem_gib, losses, gradients of the first accumulation cycle)."""
    brain = make_brain(sb, rec, ConformerEncoder, args.vocab, dropin, device, world > 1, args.grad_accumulation_factor, args.seed, args.dropout,
                       deferred_check, args.dropins)
    rank = int(os.environ.get("RANK", "0"))
    batches = [make_batch(sb, args.batch, args.seconds, args.labels, args.vocab, seed=1000 * rank + i, ragged=args.ragged)
               for i in range(4)]
    losses, captured = parity_cycle(brain, batches, args.grad_accumulation_factor, args.seed)
    # --- timing ---
    for i in range(args.warmup):
        brain.fit_batch(batches[i % len(batches)])
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.reset_peak_memory_stats(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        brain.fit_batch(batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize(device)
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
    peak = torch.cuda.max_memory_allocated(device) / 2 ** 30
    out = {"ms_per_step": ms, "wall_ms_per_step": wall, "peak_mem_gib": peak, "losses": losses, **captured}
    del brain
    torch.cuda.empty_cache()
    return out


def compare(stock, dropin):
    """Parity line between the two arms (same seeds, same initial weights, same batches)."""
    rel = max(abs(a - b) / max(abs(a), 1e-12) for a, b in zip(stock["losses"], dropin["losses"]))
    out = {"loss_rel_err_max": rel, "losses_stock": stock["losses"], "losses_dropin": dropin["losses"]}
    for k in ("head_grad", "head_bias_grad", "enc_proj_grad", "dec_rnn_hh_grad", "dec_rnn_ih_grad"):
        a, b = stock[k].double(), dropin[k].double()
        out[k + "_max_err_over_max"] = ((a - b).abs().max() / a.abs().max()).item()
        out[k + "_rel_l2"] = ((a - b).norm() / a.norm()).item()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["both", "stock", "dropin"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=16.0)
    ap.add_argument("--labels", type=int, default=99)
    ap.add_argument("--vocab", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--grad-accumulation-factor", dest="grad_accumulation_factor", type=int, default=4)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--dropins", default="all", choices=["joint", "joint+proj", "all"],
                    help="which drop-ins the dropin arm uses: the fused joint + loss only (round-2 sessions 1-12), plus the projection "
                         "GEMMs (N1), plus the prediction network (N3)")
    ap.add_argument("--deferred-check", dest="deferred_check", action="store_true",
                    help="third arm: drop-ins + tsasr_b200.monitor.install(brain) (no loss.isfinite() host sync between forward and backward)")
    args = ap.parse_args()
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=device,
                                             timeout=datetime.timedelta(seconds=300))
    sb, rec, ConformerEncoder = import_reference()
    res = {}
    if args.arm in ("both", "stock"):
        res["stock"] = run_arm(sb, rec, ConformerEncoder, args, False, device, world)
    if args.arm in ("both", "dropin"):
        res["dropin"] = run_arm(sb, rec, ConformerEncoder, args, True, device, world)
        if args.deferred_check:
            res["dropin_deferred_check"] = run_arm(sb, rec, ConformerEncoder, args, True, device, world, deferred_check=True)
    if rank == 0:
        cells = args.batch * world  # utterances per step over all ranks
        out = {"what": "full fit_batch of train_librispeechmix_scratch.py TSASR (causal Conformer, injection_mode=cat, V=%d), "
                       "synthetic %.0f s audio, B=%d per GPU, %d labels, grad_accumulation_factor=%d, reference per-module DDP"
                       % (args.vocab, args.seconds, args.batch, args.labels, args.grad_accumulation_factor),
               "n_gpus": world, "utterances_per_step": cells, "dropins": args.dropins}
        for k, v in res.items():
            out[k] = {kk: vv for kk, vv in v.items() if not isinstance(vv, torch.Tensor)}
        if "stock" in res and "dropin" in res:
            out["parity"] = compare(res["stock"], res["dropin"])
            out["speedup_fit_batch"] = res["stock"]["ms_per_step"] / res["dropin"]["ms_per_step"]
        if "dropin_deferred_check" in res:
            out["parity_deferred_check_vs_dropin"] = compare(res["dropin"], res["dropin_deferred_check"])
            if "stock" in res:
                out["speedup_fit_batch_deferred_check"] = res["stock"]["ms_per_step"] / res["dropin_deferred_check"]["ms_per_step"]
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
