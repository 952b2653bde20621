#!/usr/bin/env python
"""bench.py -- lattice cells/s of the fused joint + RNN-T loss forward+backward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    extras: --shape recipe|config4, --ragged, --global-batch 128 (strong scaling), --sustain-s S, --config insitu

A "step" is one forward+backward pass of the hot path over one batch of synthetic input:
joint("sum") + activation + head Linear + log-softmax + RNN-T loss, gradients w.r.t. enc_out,
dec_out, W, b (BASELINE configs[1]: B=16, T=400, U=100, V=1000, H=640, bf16 joint / fp32 lattice).
At N > 1 every rank runs the same per-GPU batch (utterance sharding, weak scaling, global batch
16*N = 128 at N=8) and the step ends with one NCCL all-reduce of {dW, db, loss}; the e2e leg goes
through DistributedDataParallel(head) exactly as SpeechBrain wraps it.  --global-batch G splits a fixed,
ragged global batch over the ranks instead (strong scaling, slowest rank reported).

Prints ONE JSON line (rank 0).  ``value`` is device-timed with inputs resident in HBM; ``e2e`` goes
through the public drop-in modules with pinned HOST buffers (H2D + D2H inside the timed region);
``sustained`` is the same step back to back for seconds (power-cap regime); ``reference_gpu`` is the
reference's own GPU path (eager + torchaudio CUDA) on the same box; ``cpu_baseline`` its CPU path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=16, T=400, U=100, V=1000, H=640, act="leaky_relu", act_param=0.01, blank=0)
# other shapes (--shape): never the headline, each names itself in config.workload
SHAPES = {
    "config2": CFG,
    # the recipe as shipped (hparams/LibriSpeechMix/conformer-t_scratch.yaml:76-77): 29 characters, U ~ 0.6 T
    "recipe": dict(B=16, T=400, U=240, V=29, H=640, act="leaky_relu", act_param=0.01, blank=0),
    # BASELINE configs[3], long-mixture stress
    "config4": dict(B=8, T=750, U=200, V=5000, H=640, act="leaky_relu", act_param=0.01, blank=0),
}
CPU_SAMPLE = dict(B=4, T=200, U=40, V=1000, H=640)  # BASELINE configs[0] shape: ~2 s per CPU step
METRIC = "lattice cells/sec (B*T*U) joint+RNN-T fwd+bwd"
UNIT = "cells/s"


def synth(cfg, device, seed=0, ragged=False):
    """Synthetic inputs of SURVEY.md section 8d: enc/dec ~ 0.5*randn, W/b ~ nn.Linear init; full lengths, or
    (ragged) T_b in [0.6 T, T] and label counts in [0.4 U, U-1] with one utterance at the maximum of each."""
    g = torch.Generator().manual_seed(seed)
    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    enc = 0.5 * torch.randn(B, T, H, generator=g)
    dec = 0.5 * torch.randn(B, U, H, generator=g)
    bound = 1.0 / H ** 0.5
    W = (torch.rand(V, H, generator=g) * 2 - 1) * bound
    b = (torch.rand(V, generator=g) * 2 - 1) * bound
    targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
    ll = torch.full((B,), T, dtype=torch.int32)
    tl = torch.full((B,), U - 1, dtype=torch.int32)
    if ragged:
        ll = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
        tl = torch.randint(int(0.4 * U), U, (B,), generator=g, dtype=torch.int32)
        ll[0], tl[0] = T, U - 1
    return [x.to(device) for x in (enc, dec, W, b, targets, ll, tl)]


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks line").

    NVML is polled in-process every few milliseconds (the timed region of a default run is ~100 ms, too
    short for `nvidia-smi -lms`); every sample carries a host timestamp and only samples taken between
    mark_begin() and mark_end() are reported.  Falls back to one `nvidia-smi` query if NVML is missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.004):
        self.rows, self.t0, self.t1 = [], None, None
        self.period, self.stop_flag, self.h, self.max_mhz = period_s, False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001  (no NVML: nvidia-smi fallback in stop())
            self.h = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001  (older bindings)
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), float(mhz), int(mask)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.h is None:
            return self._smi_once()
        self.thread.join(timeout=1.0)
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        window = "timed region"
        if len(rows) < 3:  # very short timed regions: take the nearest samples around it as well
            rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
            window = "timed region +-50 ms"
        sm = [r[1] for r in rows]
        reasons = [n for n, bit in self.REASONS if any(r[2] & bit for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "sm_mhz_min": min(sm) if sm else None, "window": window, "source": "nvml"}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True,
                                 text=True, timeout=10).stdout.strip().splitlines()[0]
            parts = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(parts[0]), "sm_max_mhz": float(parts[1]),
                    "reasons": [n for i, n in enumerate(names) if parts[2 + i].lower().startswith("active")],
                    "samples": 1, "window": "after the timed region", "source": "nvidia-smi"}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def measured_peaks():
    """-> (bf16 burst TFLOP/s, bf16 sustained TFLOP/s, HBM GB/s, source).  B200_PROFILING.md: the burst figure is the
    denominator for a kernel timed alone / in a short region at full clocks (the default run: ~50 ms of 2.4 ms steps,
    1965 MHz, no power cap), the sustained one for a kernel timed inside a seconds-long loop under the power cap (the
    ``sustained`` leg)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        burst = d.get("bf16_tflops", d.get("bf16_tflops_sustained"))
        return burst, d.get("bf16_tflops_sustained", burst), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_step(sample, threads):
    """The reference's CPU path for this hot path: torch eager joint + Linear + torchaudio rnnt_loss
    (what SB/nnet/losses.py:72-79 calls) + backward, restated in oracle/reference_chain.py."""
    from oracle.reference_chain import reference_joint_loss_fwd_bwd

    enc, dec, W, b, targets, ll, tl = synth(sample, "cpu", seed=1)
    torch.set_num_threads(threads)
    best = None
    for i in range(4):  # 1 warm-up + best of 3
        t0 = time.perf_counter()
        reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=False, reduction="mean")
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    return sample["B"] * sample["T"] * sample["U"] / best, best


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_single_thread(sample):
    """Same chain on ONE host thread (SURVEY 8d asks for both): one untimed warm-up on all threads has already run,
    so a single timed step bounds the extra cost to a few seconds."""
    from oracle.reference_chain import reference_joint_loss_fwd_bwd

    enc, dec, W, b, targets, ll, tl = synth(sample, "cpu", seed=1)
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, round_bf16=False, reduction="mean")
    dt = time.perf_counter() - t0
    return sample["B"] * sample["T"] * sample["U"] / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    from oracle.reference_chain import reference_joint_loss_fwd_bwd

    sample = CPU_SAMPLE
    enc, dec, W, b, targets, ll, tl = synth(sample, "cpu", seed=1)
    torch.set_num_threads(threads)
    for _ in range(max(1, min(args.warmup, 2))):
        reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, reduction="mean")
    times = []
    for _ in range(min(steps, 10)):
        t0 = time.perf_counter()
        reference_joint_loss_fwd_bwd(enc, dec, W, b, targets, ll, tl, 0, "leaky_relu", 0.01, reduction="mean")
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = sample["B"] * sample["T"] * sample["U"] / (ms / 1e3)
    sample_txt = ("torch eager Transducer_joint(sum,LeakyReLU)+Linear + torchaudio.functional.rnnt_loss CPU + backward "
                  "(the libraries the reference calls, chain restated in oracle/reference_chain.py); bounded sample "
                  f"B={sample['B']},T={sample['T']},U={sample['U']},V={sample['V']},H={sample['H']} fp32 of the config workload")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(CFG, "config2", max(1, args.gpus), False), "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample_txt,
                         "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def kernel_lines(kernel_ms, cfg, peak_burst, peak_sust, peak_hbm, active_tiles=None, live_tiles=None, work_cells=None):
    """Per-kernel device time of one step (CUDA events inside the library, 5 extra steps after the timed region)
    with the bound that applies: GEMM kernels against the measured bf16 peaks (burst: the regime of this short timed
    region; sustained beside it), the lattice DP and the fold kernels against the measured HBM copy bandwidth
    (algorithmic bytes).  The forward GEMM is credited the algorithmic 2*M*H*V; with tile pruning
    the three backward GEMM kernels are credited the FLOPs they EXECUTE (2 * 128 * active tiles * H * V), so their
    fraction stays a statement about the kernel, not about the pruning."""
    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    cells = work_cells if work_cells else B * T * U
    gemm = 2.0 * cells * H * V
    gemm_bwd = 2.0 * 128 * active_tiles * H * V if active_tiles else gemm
    tiles = active_tiles if active_tiles else B * ((T + 15) // 16) * ((U + 7) // 8)
    algo = {
        "joint_gemm_kernel<FWD>": ("tensor", gemm), "joint_gemm_kernel<GRAD>": ("tensor", gemm_bwd),
        "dj_gemm_kernel": ("tensor", gemm_bwd), "dw_gemm_kernel": ("tensor", gemm_bwd),
        "alpha_beta_kernel": ("hbm", 24.0 * cells),
        "reduce_dpre_kernel": ("hbm", 4.0 * (tiles * 24 * H + B * T * H + 2 * B * U * H)),
        "reduce_dw_kernel": ("hbm", 4.0 * V * H * 10),
    }
    lines = []
    for name, ms in kernel_ms.items():
        bound, work = algo.get(name, (None, None))
        line = {"kernel": name, "ms": round(ms, 4)}
        if bound == "tensor":
            a = work / (ms / 1e3) / 1e12
            line.update(bound="tensor", achieved=round(a, 1), unit="TFLOP/s", frac=round(a / peak_burst, 3),
                        frac_sustained=round(a / peak_sust, 3))
        elif bound == "hbm":
            a = work / (ms / 1e3) / 1e9
            line.update(bound="hbm", achieved=round(a, 1), unit="GB/s", frac=round(a / peak_hbm, 3))
        lines.append(line)
    return lines


def workload_name(cfg, shape, world, strong):
    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    tag = {"config2": "BASELINE configs[1]; configs[2] at N>1", "recipe": "the recipe as shipped: 29 characters, U = 0.6 T (not a BASELINE config)",
           "config4": "BASELINE configs[3], long-mixture stress"}[shape]
    per = f"B={B} per GPU" + (f" (global batch {B * world} split over {world} GPUs, strong scaling)" if strong else "")
    return (f"conformer-t_scratch joint+RNN-T loss fwd+bwd, synthetic {per} T={T} U={U} V={V} H={H}, "
            f"bf16 joint / fp32 lattice ({tag})")


def reference_gpu_leg(cfg, dev):
    """The reference's own GPU path on this box (SURVEY.md section 8d "Reference GPU path", the existing-Blackwell bar):
    eager Transducer_joint("sum", LeakyReLU) + nn.Linear head + torchaudio.functional.rnnt_loss CUDA (sm_100 cubins in
    the wheel) + backward, fp32 and under fp16 autocast (torchaudio rejects bf16 logits), inputs resident in HBM."""
    from torchaudio.functional import rnnt_loss

    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    g = torch.Generator().manual_seed(0)
    enc = (0.5 * torch.randn(B, T, H, generator=g)).to(dev)
    dec = (0.5 * torch.randn(B, U, H, generator=g)).to(dev)
    head = torch.nn.Linear(H, V).to(dev)
    tg = torch.randint(1, V, (B, U - 1), generator=g).to(dev)
    il, tl = torch.ones(B, device=dev), torch.ones(B, device=dev)
    act = torch.nn.LeakyReLU(cfg["act_param"])

    def step(autocast_dtype):
        e_, d_ = enc.detach().requires_grad_(), dec.detach().requires_grad_()
        with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
            logits = head(act(e_[..., None, :] + d_[:, None, ...]))      # transducer_joint.py:73-74,95; linear.py:74
        in_l = (il * logits.shape[1]).round().int()                      # losses.py:58-59
        tg_l = (tl * tg.shape[1]).round().int()
        loss = rnnt_loss(logits, tg.int(), in_l, tg_l, blank=0, reduction="mean")
        loss.backward()
        head.zero_grad(set_to_none=True)
        return loss

    out = {"what": "eager joint + Linear + torchaudio rnnt_loss (CUDA) + backward on the same shape, CUDA events, 2 warm-ups, median of 5"}
    for name, dt in (("fp32", None), ("fp16_autocast", torch.float16)):
        try:
            torch.cuda.reset_peak_memory_stats(dev)
            base = torch.cuda.memory_allocated(dev)
            for _ in range(2):
                loss = step(dt)
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(5):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                step(dt)
                e_.record()
                torch.cuda.synchronize(dev)
                ts.append(s_.elapsed_time(e_))
            ms = statistics.median(ts)
            out[name] = {"ms_per_step": ms, "value": B * T * U / (ms / 1e3), "unit": UNIT, "loss": float(loss),
                         "peak_extra_mem_gib": (torch.cuda.max_memory_allocated(dev) - base) / 2 ** 30}
        except Exception as ex:  # noqa: BLE001  (e.g. out of memory on a shared box: report, do not fail the bench)
            out[name] = {"unavailable": f"{type(ex).__name__}: {str(ex)[:200]}"}
        torch.cuda.empty_cache()
    return out


def next_rows_leg(cfg, dev):
    """SURVEY.md section 8f rows N1 and N3 on this box, beside the reference's own ops at the workload's shape: the two
    projections (tsasr_b200.Linear vs torch fp32 nn.Linear = what speechbrain's Linear runs) and the prediction network
    (tsasr_b200.Embedding + LSTM vs one-hot embedding -> pack_padded_sequence with its .cpu() -> cuDNN LSTM), forward +
    backward through autograd, CUDA events, 3 warm-ups, median of 10.  Not part of `value` / `e2e` (BASELINE's metric is the
    joint + loss path); reported so that the rows either side of the path have driver-visible numbers."""
    import tsasr_b200
    from tsasr_b200 import _lib

    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    out = {"what": "forward + backward, us, ours vs the reference's ops on the same box (CUDA events, median of 10)"}

    def med(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(10):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            fn()
            e_.record()
            torch.cuda.synchronize(dev)
            ts.append(s_.elapsed_time(e_) * 1e3)
        return statistics.median(ts)

    try:
        torch.manual_seed(0)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        for name, R, K in (("encoder_proj", B * T, 256), ("decoder_proj", B * U, 512)):   # yaml:172-174,187-189
            ours, ref = tsasr_b200.Linear(H, input_size=K).to(dev), torch.nn.Linear(K, H).to(dev)
            x = torch.randn(R, K, device=dev, requires_grad=True)
            gy = torch.randn(R, H, device=dev)
            out[name] = {"shape": [R, K, H], "ours_us": med(lambda: ours(x).backward(gy)), "reference_ops_us": med(lambda: ref(x).backward(gy))}
            _lib.kernel_timing(True)   # GPU time of our kernels alone (the event-bracketed loop above is host-bound at this size)
            ours(x).backward(gy)
            torch.cuda.synchronize(dev)
            out[name]["ours_kernels_us"] = {k: round(v[0] * 1e3, 1) for k, v in _lib.kernel_timings().items()}
            _lib.kernel_timing(False)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        Hd = 512                                                                         # yaml:176-185
        emb = tsasr_b200.Embedding(num_embeddings=V, consider_as_one_hot=True, blank_id=0).to(dev)
        ours = tsasr_b200.LSTM(input_shape=[None, None, V - 1], hidden_size=Hd).to(dev)
        ref = torch.nn.LSTM(V - 1, Hd, batch_first=True).to(dev)
        ref.load_state_dict({k[4:]: v for k, v in ours.state_dict().items()})
        tokens = torch.randint(1, V, (B, U), device=dev)
        tokens[:, 0] = 0
        rel = torch.rand(B, device=dev) * 0.6 + 0.4
        rel[0] = 1.0
        gy = torch.randn(B, U, Hd, device=dev)

        def ours_step():
            ours(emb(tokens), lengths=rel)[0].backward(gy)

        def ref_step():
            x = torch.nn.functional.embedding(tokens, emb.Embedding.weight, padding_idx=0)          # SB/nnet/embedding.py:114
            packed = torch.nn.utils.rnn.pack_padded_sequence(x, (rel * U).cpu(), batch_first=True, enforce_sorted=False)  # RNN.py:35
            torch.nn.utils.rnn.pad_packed_sequence(ref(packed)[0], batch_first=True)[0].backward(gy)

        with torch.no_grad():
            a = ours(emb(tokens), lengths=rel)[0]
            x = torch.nn.functional.embedding(tokens, emb.Embedding.weight, padding_idx=0)
            packed = torch.nn.utils.rnn.pack_padded_sequence(x, (rel * U).cpu(), batch_first=True, enforce_sorted=False)
            b = torch.nn.utils.rnn.pad_packed_sequence(ref(packed)[0], batch_first=True)[0]
        out["predictor"] = {"shape": {"B": B, "U": U, "V": V, "hidden": Hd}, "ours_us": med(ours_step), "reference_ops_us": med(ref_step),
                            "max_abs_diff_forward": float((a - b).abs().max())}
        _lib.kernel_timing(True)
        ours_step()
        torch.cuda.synchronize(dev)
        out["predictor"]["ours_kernels_us"] = {k: round(v[0] * 1e3, 1) for k, v in _lib.kernel_timings().items()}
        _lib.kernel_timing(False)
    except Exception as ex:  # noqa: BLE001  (report, do not fail the bench)
        out["unavailable"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    torch.cuda.empty_cache()
    return out


def run_insitu(args, rank, world):
    """--config insitu: BASELINE configs[4], the full fit_batch of the real TSASR Brain with the drop-ins beside the stock
    modules (tools/insitu_step.py).  Needs the reference install under baseline/_ref."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import insitu_step

    if insitu_step.find_reference() is None:
        if rank == 0:
            print(json.dumps({"config": {"workload": "insitu"}, "unavailable": "no reference install (run tools/install_reference.sh)"}))
        return
    sys.argv = [sys.argv[0], "--steps", str(max(2, args.steps)), "--warmup", str(max(2, args.warmup))]
    insitu_step.main()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="hotpath", choices=["hotpath", "insitu"],
                    help="hotpath: the joint + loss path alone (the BASELINE metric); insitu: full TSASR fit_batch (configs[4])")
    ap.add_argument("--shape", default="config2", choices=sorted(SHAPES), help="config2 is the headline; the others name themselves")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: this many utterances split over the ranks (BASELINE configs[2]: 128); default: weak, B=16 per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference's own GPU path (eager + torchaudio CUDA) leg")
    ap.add_argument("--sustain-s", type=float, default=3.0,
                    help="seconds of back-to-back steps (no L2 flush, power-cap regime) reported beside the burst number; 0: off")
    ap.add_argument("--ragged", action="store_true",
                    help="ragged utterance lengths (SURVEY 8d) instead of the full-length batch the headline is quoted on")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config == "insitu":
        run_insitu(args, rank, world)
        return

    import tsasr_b200
    from tsasr_b200 import _lib, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (tsasr_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a short collective timeout: a rank-divergence bug must fail in seconds, not hold the box for 10 minutes
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    W_steps = max(3, args.warmup)
    K = max(1, args.steps)

    cfg = dict(SHAPES[args.shape])
    strong = args.global_batch > 0
    if strong:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        cfg["B"] = args.global_batch // world
    ragged = args.ragged or strong  # the strong-scaling arm is the ragged global batch of SURVEY 8e: the slowest rank is reported
    B, T, U, V, H = (cfg[k] for k in "BTUVH")
    if strong:
        # ONE global batch (fixed seed, ragged), dealt to the ranks in contiguous chunks of utterances;
        # every rank pads to the longest utterance of ITS chunk, as the recipe's per-rank collate function does
        gcfg = dict(cfg, B=args.global_batch)
        g_enc, g_dec, Wt, bias, g_tg, g_ll, g_tl = synth(gcfg, "cpu", seed=0, ragged=True)
        sl = slice(rank * B, (rank + 1) * B)
        ll, tl = g_ll[sl].clone(), g_tl[sl].clone()
        T, U = int(ll.max()), int(tl.max()) + 1
        enc, dec, targets = g_enc[sl, :T].contiguous(), g_dec[sl, :U].contiguous(), g_tg[sl, : U - 1].contiguous()
        enc, dec, Wt, bias, targets, ll, tl = (x.to(dev) for x in (enc, dec, Wt, bias, targets, ll, tl))
        cfg.update(T=T, U=U)
    else:
        enc, dec, Wt, bias, targets, ll, tl = synth(cfg, dev, seed=rank, ragged=ragged)
    cells = B * T * U  # the metric counts the padded lattice (BASELINE.json: B*T*U, U = logits.shape[2])
    # cells that carry work: sum_b T_b * (labels_b + 1); the roofline lines credit FLOPs for these only (== cells when
    # every utterance has the full length, the headline case)
    work_cells = int((ll.long() * (tl.long() + 1)).sum().item())
    enc16, dec16, W16 = enc.bfloat16().contiguous(), dec.bfloat16().contiguous(), Wt.bfloat16().contiguous()
    dcost = torch.full((B,), 1.0 / (B * world), dtype=torch.float32, device=dev)
    act = _lib.ACT_CODES[cfg["act"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    comm = torch.empty(V * H + V + 1, dtype=torch.float32, device=dev)

    fwd_ev = []

    def step(record, prune=None):
        """One fwd+bwd pass; inputs already resident in HBM (bf16 operands, fp32 bias)."""
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        lat2, logz = ops.joint_fwd(enc16, dec16, W16, bias, targets, ll, tl, cfg["blank"], act, cfg["act_param"])
        if record:
            e1.record()
            fwd_ev.append((e0, e1))
        alpha, beta, cost, _, _ = ops.alpha_beta(lat2, ll, tl, B, T, U)
        d_enc, d_dec, dW, db = ops.joint_bwd(enc16, dec16, W16, bias, targets, ll, tl, cfg["blank"], act, cfg["act_param"],
                                             lat2, logz, alpha, beta, cost, dcost, prune_log2_eps=prune)
        if dist is not None:
            comm[: V * H].copy_(dW.view(-1))
            comm[V * H: V * H + V].copy_(db)
            comm[-1] = cost.sum() / (B * world)
            dist.all_reduce(comm)
        return cost

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def fence():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None  # polls from the warm-up on; reports the timed window
    for _ in range(W_steps):
        flush.zero_()
        step(False)
    fence()
    launches0 = _lib.launch_count()
    evs = []
    if sampler:
        sampler.mark_begin()
    for _ in range(K):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(True)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if sampler:
        sampler.mark_end()
    fence()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(sum(s.elapsed_time(e) for s, e in evs)) / K
    total_cells = world * cells  # whole-job lattice cells per step
    if strong and dist is not None:  # every rank pads to its own longest utterance: add the per-rank lattices up
        tc = torch.tensor([float(cells)], dtype=torch.float64, device=dev)
        dist.all_reduce(tc)
        total_cells = int(tc.item())
    value = total_cells / (ms_per_step / 1e3)
    fwd_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in fwd_ev)

    # ---- the same steps with backward tile pruning switched off (reported beside the headline, not instead of it) ----
    prune_eps = ops.default_prune_log2_eps()
    active_tiles, live_tiles = (ops.last_backward_tile_stats(dev) if prune_eps < 0 else (None, None))
    dense_ms = None
    if prune_eps < 0:
        for _ in range(3):
            flush.zero_()
            step(False, prune=0.0)
        fence()
        evs_d = []
        for _ in range(min(K, 10)):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step(False, prune=0.0)
            e.record()
            evs_d.append((s, e))
        torch.cuda.synchronize()
        dense_ms = max_over_ranks(sum(a.elapsed_time(b_) for a, b_ in evs_d) / len(evs_d))

    # ---- sustained: seconds of back-to-back steps, no L2 flush in between (the regime of a training loop: power cap,
    #      lower clocks); each step's working set (GBs of operand images) exceeds L2 by itself ----
    sustained = None
    if args.sustain_s > 0:
        n_target = max(20, int(args.sustain_s * 1e3 / ms_per_step * 1.15))  # same count on every rank (collectives inside)
        n_target = int(max_over_ranks(float(n_target)))
        s_sampler = ClockSampler(local_rank, period_s=0.02) if rank == 0 else None
        fence()
        if s_sampler:
            s_sampler.mark_begin()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n_target):
            step(False)
            if i % 64 == 63:
                torch.cuda.current_stream().synchronize()  # bound the launch queue; ~10 us every 64 steps
        e.record()
        torch.cuda.synchronize()
        if s_sampler:
            s_sampler.mark_end()
        s_ms = max_over_ranks(s.elapsed_time(e)) / n_target
        sustained = {"seconds": s_ms * n_target / 1e3, "steps": n_target, "ms_per_step": s_ms,
                     "value": total_cells / (s_ms / 1e3), "unit": UNIT, "clocks": s_sampler.stop() if s_sampler else None,
                     "what": "back-to-back steps for seconds, no L2 flush, CUDA events around the whole loop, max over ranks"}
        fence()

    # ---- e2e: public drop-in modules, pinned host inputs, H2D + D2H inside the timed region ----
    joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
    head = torch.nn.Linear(H, V).to(dev)
    with torch.no_grad():
        head.weight.copy_(Wt)
        head.bias.copy_(bias)
    # N > 1: the head goes through DistributedDataParallel exactly as SpeechBrain wraps it (SB/core.py:1479-1483); its
    # reducer all-reduces (averages) dW / db during backward -- no hand-written collective on this path
    head_call = torch.nn.parallel.DistributedDataParallel(head, device_ids=[dev]) if dist is not None else head
    h_enc, h_dec = enc.cpu().pin_memory(), dec.cpu().pin_memory()
    h_tg = targets.cpu().long().pin_memory()
    # relative lengths (SpeechBrain convention)
    h_il, h_tl = (ll.cpu().float() / T).pin_memory(), (tl.cpu().float() / (U - 1)).pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (h_enc, h_dec, h_tg, h_il, h_tl))

    copy_stream = torch.cuda.Stream(dev)

    def h2d_async():
        """This step's inputs, pinned host -> device on the copy stream (what a pin_memory DataLoader prefetch does)."""
        with torch.cuda.stream(copy_stream):
            bufs = tuple(x.to(dev, non_blocking=True) for x in (h_enc, h_dec, h_tg, h_il, h_tl))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    def e2e_compute(bufs, ev, keep_grads=False):
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        for x in bufs:
            x.record_stream(cur)
        e_, d_, tg_, il_, tl_ = bufs
        e_.requires_grad_()
        d_.requires_grad_()
        logits = head_call(joiner(e_[..., None, :], d_[:, None, ...]))  # train_librispeechmix_scratch.py:132,135
        loss = tsasr_b200.transducer_loss(logits, tg_, il_, tl_, blank_index=0, reduction="mean", use_torchaudio=True)
        loss.backward()
        if not keep_grads:
            head.zero_grad(set_to_none=True)
        return loss  # still on the device: the caller reads it (D2H) after queueing the next step's input copy

    def e2e_loop(n):
        """n steps; step i+1's host->device copy is queued (copy stream) right after step i's kernels have been queued and
        before step i's loss is read, so both the copy and the host work of issuing it overlap step i's kernels -- what a
        pin_memory DataLoader worker does beside the training loop.  Every step still ends with a synchronous D2H read of
        its own loss."""
        val = None
        cur = h2d_async()
        for i in range(n):
            loss = e2e_compute(*cur)
            if i + 1 < n:
                cur = h2d_async()
            val = loss.item()  # D2H read of the step's result
        return val

    e2e_loop(3)
    fence()
    Ke = min(K, 10)
    flush.zero_()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    loss_val = e2e_loop(Ke)  # every step's H2D copy and D2H loss read happen inside this region
    e.record()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(s.elapsed_time(e) / Ke)
    e2e_value = total_cells / (e2e_ms / 1e3)

    # ---- N > 1: the head gradient DDP produced must be the rank average of the local gradients ----
    ddp_check = None
    if dist is not None:
        e2e_compute(*h2d_async(), keep_grads=True).item()
        g_ddp = [head.weight.grad.detach().clone(), head.bias.grad.detach().clone()]
        head.zero_grad(set_to_none=True)
        e_, d_ = enc.detach().clone().requires_grad_(), dec.detach().clone().requires_grad_()
        loss = tsasr_b200.transducer_loss(head(joiner(e_[..., None, :], d_[:, None, ...])), targets.long(), ll.float() / T,
                                          tl.float() / (U - 1), blank_index=0, reduction="mean", use_torchaudio=True)  # no DDP wrapper
        loss.backward()
        g_loc = [head.weight.grad.detach().clone(), head.bias.grad.detach().clone()]
        head.zero_grad(set_to_none=True)
        for g_ in g_loc:
            dist.all_reduce(g_)
            g_ /= world
        err = max(((a - b_).abs().max() / b_.abs().max()).item() for a, b_ in zip(g_ddp, g_loc))
        gathered = [torch.zeros_like(g_ddp[0]) for _ in range(world)]
        dist.all_gather(gathered, g_ddp[0])
        ddp_check = {"head_grad_vs_rank_average_of_local_grads_max_rel": max_over_ranks(err),
                     "identical_on_all_ranks": all(torch.equal(gathered[0], x) for x in gathered),
                     "route": "DistributedDataParallel(head, device_ids=[dev]) called with the deferred handle (SB/core.py:1479-1483)"}

    # ---- per-kernel pass (after the timed regions): the library brackets each of its launches with CUDA events ----
    # (every rank runs the steps -- they contain the all-reduce -- rank 0 reports its own kernels)
    _lib.kernel_timing(True)
    n_prof = 5
    for _ in range(n_prof):
        flush.zero_()
        step(False)
    torch.cuda.synchronize()
    kernel_ms = {k: v[0] / n_prof for k, v in _lib.kernel_timings().items()}
    _lib.kernel_timing(False)

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        del flush
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_leg(cfg, dev)
    next_rows = None
    if rank == 0 and world == 1 and not args.no_reference_gpu and cfg["B"] <= 64:
        next_rows = next_rows_leg(cfg, dev)

    if rank == 0:
        peak_burst, peak_sust, peak_hbm, peak_src = measured_peaks()
        flops = 2.0 * work_cells * H * V  # algorithmic FLOPs of the forward joint GEMM launch (cells inside the T_b x U_b rectangles)
        achieved = flops / (fwd_ms / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and args.shape == "config2" and not strong:
            traffic = json.load(open(tpath)).get("joint_gemm_kernel_fwd_dram_bytes_per_launch")
        # step-level fractions: ALGORITHMIC FLOPs only (forward GEMM + dJ + dW; the softmax-recompute GEMM is overhead).
        # With tile pruning the backward executes 2 * 128 * active tiles * H * V per GEMM, and only that is credited.
        bwd_exec = 2.0 * 128 * active_tiles * H * V if active_tiles else flops
        step_exec_tf = (flops + 2.0 * bwd_exec) / (ms_per_step / 1e3) / 1e12
        dense_tf = 3.0 * flops / (dense_ms / 1e3) / 1e12 if dense_ms else step_exec_tf
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_steps,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.shape, world, strong), "B_per_gpu": B, "T": T, "U": U, "V": V, "H": H,
                       "lengths": "ragged (T_b in [0.6T,T], labels in [0.4U,U-1])" if ragged else "full", "cells_with_work": work_cells, "activation": cfg["act"],
                       "parallelism": f"utterance-sharded dp{world}",
                       "l2": "flushed between timed steps with a 256 MiB write (untimed); step timed with CUDA events"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                    "api": "Transducer_joint -> nn.Linear head" + (" inside DistributedDataParallel" if dist is not None else "") +
                           " -> transducer_loss(handle) -> backward, fp32 pinned host inputs; "
                           "each step copies its own inputs H2D (copy stream, issued one step ahead) and reads its loss D2H; "
                           "steps timed back to back, per-step working set (GBs of operand images) exceeds L2",
                    "loss": loss_val},
            "gpu_launches": launches,
            "roofline": {"kernel": "joint_gemm_kernel<MODE_FWD> (tcgen05 joint GEMM + online log-softmax)", "bound": "tensor",
                         "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                         "frac_burst": achieved / peak_burst, "frac_sustained": achieved / peak_sust,
                         "peak_sustained": peak_sust, "traffic": traffic, "peak_source": peak_src,
                         "peak_choice": "burst bf16 peak: the kernel is timed inside a ~50 ms region of 2-3 ms steps at full clocks "
                                        "(see clocks); the fraction of the sustained (power-capped) peak is given beside it",
                         "kernel_ms": fwd_ms, "algorithmic_flops_per_launch": flops,
                         "step": {"executed_algorithmic_tflops": step_exec_tf, "frac_burst": step_exec_tf / peak_burst,
                                  "frac_sustained": step_exec_tf / peak_sust,
                                  "what": "fwd 2MHV + dJ + dW on the tiles the backward executes (recompute GEMM not counted) / ms_per_step"},
                         "dense_step": {"algorithmic_tflops": dense_tf, "frac_burst": dense_tf / peak_burst,
                                        "frac_sustained": dense_tf / peak_sust,
                                        "what": "6MHV / ms_per_step of the run with tile pruning off"}},
            "clocks": clocks,
            "kernels": kernel_lines(kernel_ms, cfg, peak_burst, peak_sust, peak_hbm, active_tiles, live_tiles, work_cells),
        }
        out["config"]["backward_tile_pruning"] = (
            {"log2_eps": prune_eps, "active_tiles": active_tiles, "live_tiles": live_tiles,
             "what": "128-cell tiles whose largest alignment posterior exp(alpha+beta-L) < 2^log2_eps are left out of the "
                     "backward (every dlogits term carries that factor; below fp32 resolution of the kept terms); "
                     "TSASR_PRUNE_LOG2_EPS=0 switches it off"} if prune_eps < 0 else "off")
        if dense_ms is not None:
            out["dense_backward"] = {"ms_per_step": dense_ms, "value": total_cells / (dense_ms / 1e3), "unit": UNIT,
                                     "what": "same steps with tile pruning off (every live tile recomputed and fed to the GEMMs)"}
        if sustained is not None:
            sust_tf = (flops + 2.0 * bwd_exec) / (sustained["ms_per_step"] / 1e3) / 1e12
            sustained["step_frac_sustained_peak"] = sust_tf / peak_sust
            out["sustained"] = sustained
        if ddp_check is not None:
            out["ddp_check"] = ddp_check
        if ref_gpu is not None:
            for k in ("fp32", "fp16_autocast"):
                if "ms_per_step" in ref_gpu.get(k, {}):
                    ref_gpu[k]["speedup_of_e2e_over_it"] = ref_gpu[k]["ms_per_step"] / e2e_ms
            out["reference_gpu"] = ref_gpu
        if next_rows is not None:
            out["next_rows"] = next_rows
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, secs = cpu_reference_step(CPU_SAMPLE, threads)
            out["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": ("torch eager joint+Linear + torchaudio.functional.rnnt_loss CPU + backward (the reference's CPU path, "
                           f"oracle/reference_chain.py) on B={CPU_SAMPLE['B']},T={CPU_SAMPLE['T']},U={CPU_SAMPLE['U']},V={CPU_SAMPLE['V']},"
                           f"H={CPU_SAMPLE['H']} fp32, 1 warm-up + best of 3 ({secs:.2f} s/step)"),
                "cpu_model": cpu_model()}
            v1, secs1 = cpu_reference_single_thread(CPU_SAMPLE)
            torch.set_num_threads(threads)
            out["cpu_baseline"]["single_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "s_per_step": secs1}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
