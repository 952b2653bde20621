"""world_size-2 gloo tests (CPU) of the multi-GPU host path.

The path shards by utterance: every rank runs the whole hot path on its own
batch and the only collectives are DDP's gradient all-reduce on the head's parameters (the head is
individually DDP-wrapped by SpeechBrain, SB/core.py:1469-1484) and one scalar loss all-reduce.  What can
go wrong on the host side is the deferred handle by-passing ``DDP.forward`` (gradients would silently
stay local); these tests pin that the handle goes THROUGH the wrapper and the reducer fires.
No kernels are launched here: the handle is consumed with the reference's eager math on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tsasr_b200
from oracle.reference_chain import reference_rnnt_abs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical head on every rank (DDP broadcasts anyway)
        B, T, U, H, V = 2, 6, 4, 64, 9
        head = torch.nn.Linear(H, V)
        ddp_head = torch.nn.parallel.DistributedDataParallel(head)
        g = torch.Generator().manual_seed(100 + rank)  # different utterances per rank
        enc = torch.randn(B, T, 1, H, generator=g, requires_grad=True)
        dec = torch.randn(B, 1, U, H, generator=g, requires_grad=True)
        targets = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32)
        ll = torch.full((B,), T, dtype=torch.int32)
        tl = torch.full((B,), U - 1, dtype=torch.int32)

        act = torch.nn.LeakyReLU()
        handle = tsasr_b200.JointHandle(enc, dec, act, 0, 0.01)
        logits = ddp_head(handle)  # the recipe's call: self.modules.transducer_head(joiner_out)
        assert isinstance(logits, tsasr_b200.JointHandle) and logits.has_head
        assert logits._weight is head.weight  # parameters reached THROUGH DDP.forward
        # consume with the reference's eager math (CPU stand-in for the fused kernels)
        loss = reference_rnnt_abs(logits.materialize(), targets, ll, tl, blank=0, reduction="mean")
        loss.backward()
        # scalar loss all-reduce the build adds for reporting
        rep = loss.detach().clone()
        dist.all_reduce(rep)
        rep /= world

        # what the gradient would be without DDP (local), for the cross-check on rank 0
        head2 = torch.nn.Linear(H, V)
        head2.load_state_dict(head.state_dict())
        l2 = reference_rnnt_abs(head2(act(enc.detach() + dec.detach())), targets, ll, tl, blank=0, reduction="mean")
        l2.backward()
        gathered = [torch.zeros_like(head2.weight.grad) for _ in range(world)]
        dist.all_gather(gathered, head2.weight.grad)
        expect = sum(gathered) / world
        ok = torch.allclose(head.weight.grad, expect, atol=1e-6)
        differs_from_local = not torch.allclose(head.weight.grad, head2.weight.grad, atol=1e-6)
        out.put((rank, bool(ok), bool(differs_from_local), float(rep), enc.grad is not None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_handle_through_ddp_head_all_reduces_gradients():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    res = sorted(out.get(timeout=5) for _ in range(world))
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "head gradient is not the all-reduced average"
    assert all(r[2] for r in res), "gradient equals the local one: the DDP reducer did not fire"
    assert res[0][3] == pytest.approx(res[1][3])  # same reported loss on both ranks
    assert all(r[4] for r in res)  # autograd reached enc_out on every rank


def test_bench_shards_by_utterance_weak_scaling():
    """bench.py gives every rank the same per-GPU batch (weak scaling) with rank-dependent seeds."""
    import bench

    a = bench.synth(dict(bench.CFG, B=2, T=8, U=4, H=64, V=16), "cpu", seed=0)
    b = bench.synth(dict(bench.CFG, B=2, T=8, U=4, H=64, V=16), "cpu", seed=1)
    assert a[0].shape == b[0].shape and not torch.equal(a[0], b[0])
    assert int(a[5].max()) == 8 and int(a[6].max()) == 3
