"""Two-GPU NCCL test of the drop-in modules inside DistributedDataParallel (skipped below 2 GPUs).

Closes the item SURVEY.md section 8b leaves "still to verify on the GPU box": SpeechBrain wraps the head in
``DDP(module, device_ids=[device])`` (SB/core.py:1479-1483), whose input scatter moves tensor inputs that are not on
the target device -- the storage-less ``JointHandle`` must pass through it untouched, the fused kernels must run, and
the head's gradients must come out all-reduced (average over ranks) while enc/dec gradients stay local."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["TSASR_ROOT"])
import tsasr_b200
from oracle.reference_chain import reference_joint_loss_fwd_bwd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
import datetime
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=60))
B, T, U, H, V = 3, 40, 12, 128, 50
torch.manual_seed(0)
head = torch.nn.Linear(H, V).to(dev)
with torch.no_grad():
    head.weight.copy_(head.weight.bfloat16().float())
ddp_head = torch.nn.parallel.DistributedDataParallel(head, device_ids=[dev])      # SB/core.py:1479-1483
joiner = tsasr_b200.Transducer_joint(joint="sum", nonlinearity=torch.nn.LeakyReLU)
g = torch.Generator().manual_seed(100 + rank)                                      # different utterances per rank
enc = (0.5 * torch.randn(B, T, H, generator=g)).bfloat16().float()
dec = (0.5 * torch.randn(B, U, H, generator=g)).bfloat16().float()
targets = torch.randint(1, V, (B, U - 1), generator=g)
in_rel = torch.tensor([1.0, 0.8, 0.55])
tg_rel = torch.tensor([1.0, 0.5, 0.75])
e, d = enc.to(dev).requires_grad_(), dec.to(dev).requires_grad_()
launches0 = tsasr_b200._lib.launch_count()
logits = ddp_head(joiner(e[..., None, :], d[:, None, ...]))                        # train_librispeechmix_scratch.py:132,135
assert isinstance(logits, tsasr_b200.JointHandle) and logits.has_head and logits.device == dev
loss = tsasr_b200.transducer_loss(logits, targets.to(dev), in_rel.to(dev), tg_rel.to(dev), blank_index=0)
loss.backward()
torch.cuda.synchronize()
assert tsasr_b200._lib.launch_count() > launches0                                  # the fused kernels ran
# CPU oracle of THIS rank's batch; the head's gradient must be the average over ranks
ll, tl = (in_rel * T).round().int(), (tg_rel * (U - 1)).round().int()
ref = reference_joint_loss_fwd_bwd(enc, dec, head.weight.detach().cpu(), head.bias.detach().cpu(), targets.int(), ll, tl, 0,
                                   "leaky_relu", 0.01, round_bf16=True, reduction="mean", dcost=torch.full((B,), 1.0 / B))
local = [ref["dW"].to(dev), ref["db"].to(dev)]
for t in local:
    dist.all_reduce(t)
    t /= world
rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
errs = dict(dW=rel(head.weight.grad, local[0]), db=rel(head.bias.grad, local[1]),
            d_enc=rel(e.grad.cpu(), ref["d_enc"]), d_dec=rel(d.grad.cpu(), ref["d_dec"]),
            loss=abs(loss.item() - ref["loss"].item()) / ref["loss"].item())
assert all(v < 1e-2 for k, v in errs.items() if k != "loss") and errs["loss"] < 1e-4, errs
# every rank holds the same reduced head gradient
gw = [torch.zeros_like(head.weight.grad) for _ in range(world)]
dist.all_gather(gw, head.weight.grad)
assert all(torch.equal(gw[0], x) for x in gw)

# ---- the projection / prediction-network drop-ins inside their OWN DDP wrappers: SpeechBrain wraps every module that has
# trainable parameters (SB/core.py:1469-1484): decoder (tsasr_b200.LSTM) and decoder_proj (tsasr_b200.Linear); the one-hot
# Embedding has none and stays unwrapped.  The deferred OneHotHandle must pass through DDP's input scatter, the cooperative
# recurrence must run, and the wrapped modules' gradients must come out as the rank average of the local gradients.
import copy
DDP = torch.nn.parallel.DistributedDataParallel
Vp, Hd = 30, 128
torch.manual_seed(1)
emb = tsasr_b200.Embedding(num_embeddings=Vp, consider_as_one_hot=True, blank_id=0).to(dev)
lstm = tsasr_b200.LSTM(input_shape=[None, None, Vp - 1], hidden_size=Hd).to(dev)
proj = tsasr_b200.Linear(H, input_size=Hd).to(dev)
lstm_local, proj_local, head_local = copy.deepcopy(lstm), copy.deepcopy(proj), copy.deepcopy(head)
ddp_lstm, ddp_proj = DDP(lstm, device_ids=[dev]), DDP(proj, device_ids=[dev])
tok = torch.randint(0, Vp, (B, U), generator=g)
tok[:, 0] = 0
bos_rel = torch.tensor([1.0, 0.5, 0.75])

def chain(lstm_m, proj_m, head_m):
    for m in (lstm_m, proj_m, head_m):
        m.zero_grad(set_to_none=True)
    e2 = enc.to(dev).requires_grad_()
    x = emb(tok.to(dev))
    assert isinstance(x, tsasr_b200.OneHotHandle)
    dec_out, _ = lstm_m(x, lengths=bos_rel.to(dev))
    logits2 = head_m(joiner(e2[..., None, :], proj_m(dec_out)[:, None, ...]))
    assert isinstance(logits2, tsasr_b200.JointHandle)
    l2 = tsasr_b200.transducer_loss(logits2, targets.to(dev), in_rel.to(dev), tg_rel.to(dev), blank_index=0)
    l2.backward()
    torch.cuda.synchronize()
    return l2

n0 = tsasr_b200._lib.launch_count()
chain(ddp_lstm, ddp_proj, ddp_head)
assert tsasr_b200._lib.launch_count() - n0 >= 10          # recurrence, projection GEMMs and the fused loss all ran
got = [p.grad.detach().clone() for m in (lstm, proj, head) for p in m.parameters()]
chain(lstm_local, proj_local, head_local)
want = [p.grad.detach().clone() for m in (lstm_local, proj_local, head_local) for p in m.parameters()]
for t in want:
    dist.all_reduce(t)
    t /= world
worst = max(rel(a, b) for a, b in zip(got, want))
assert worst < 1e-5, worst                                  # deterministic kernels: DDP's average IS the average of the local runs
print(f"rank {rank} ok {errs} predictor/projection DDP grads vs rank average: {worst:.1e}", flush=True)
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_drop_in_through_ddp_on_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, TSASR_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
