"""Decode-time joint step (the other caller of the joint + head modules).

The reference searchers call ``self._joint_forward_step(h_i, out_PN)`` once per decoded frame and hypothesis
batch (vendor/speechbrain/speechbrain/decoders/transducer.py:177-180, 303-309; definition :375-384):
``tjoint`` on [B,1,1,H] inputs, the classifier Linear(s) and ``LogSoftmax`` -- five tiny launches.
``joint_decode_step`` does the same arithmetic (fp32, same operands) in two small launches (csrc/decode.cu);
``patch_searcher`` swaps it into an existing ``TransducerBeamSearcher`` without touching its search logic.
"""
import ctypes
import types

import torch

from . import _lib
from .transducer_joint import activation_code

_workspaces = {}


def _workspace(dev, V):
    key = (dev, V)
    ws = _workspaces.get(key)
    if ws is None:
        n = int(_lib.load().tsasr_joint_decode_workspace_bytes(int(V)))
        ws = _workspaces[key] = torch.empty((max(n, 16),), dtype=torch.uint8, device=dev)
    return ws


def _rows(x, H):
    """(tensor keeping the storage alive, row stride in elements, rows) for a [..., H] fp32 view whose last dim is
    contiguous and whose leading dims collapse to one stride; anything else is copied."""
    if x.dtype == torch.float32 and x.stride(-1) == 1:
        lead = [(n, s) for n, s in zip(x.shape[:-1], x.stride()[:-1]) if n != 1]
        if not lead:
            return x, 0, 1
        if len(lead) == 1:
            return x, lead[0][1], lead[0][0]
    x = x.reshape(-1, H).to(torch.float32).contiguous()
    return x, H, x.shape[0]


def joint_decode_step(enc_t, dec, weight, bias=None, activation="leaky_relu", act_param=0.01):
    """log_softmax(act(enc_t + dec) @ weight.T + bias) for B hypotheses.

    enc_t, dec : [B, H] or [B,1,1,H] fp32 (strided views are read in place; a single row broadcasts over B)
    weight     : [V, H] fp32, bias : [V] fp32 or None  ->  log-probs [B, V] fp32
    """
    if not (enc_t.is_cuda and dec.is_cuda and weight.is_cuda):
        raise ValueError("tsasr_b200 needs CUDA tensors; there is no CPU path")
    V, H = weight.shape
    e, es, eb = _rows(enc_t, H)
    d, ds, db = _rows(dec, H)
    B = max(eb, db)
    if (eb != B and eb != 1) or (db != B and db != 1):
        raise ValueError(f"cannot broadcast {tuple(enc_t.shape)} with {tuple(dec.shape)}")
    if eb == 1:
        es = 0
    if db == 1:
        ds = 0
    w = weight if (weight.dtype == torch.float32 and weight.is_contiguous()) else weight.detach().to(torch.float32).contiguous()
    b = bias
    if b is not None and not (b.dtype == torch.float32 and b.is_contiguous()):
        b = b.detach().to(torch.float32).contiguous()
    dev = w.device
    out = torch.empty((B, V), dtype=torch.float32, device=dev)
    ws = _workspace(dev, V)
    code = _lib.ACT_CODES[activation] if isinstance(activation, str) else int(activation)
    rc = _lib.load().tsasr_joint_decode_step(
        e.data_ptr(), d.data_ptr(), es, ds, w.data_ptr(), b.data_ptr() if b is not None else None, B, H, V, code,
        float(act_param), out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        _lib.check(rc)
    return out


def _head_linear(classifier_network):
    """The single Linear of ``classifier_network`` (SpeechBrain ``Linear`` keeps it in ``.w``), else None."""
    if len(classifier_network) != 1:
        return None
    layer = classifier_network[0]
    lin = getattr(layer, "w", layer)
    return lin if isinstance(lin, torch.nn.Linear) else None


def fused_joint_forward_step(tjoint, classifier_network, softmax):
    """Callable with the signature and result of ``_joint_forward_step`` (transducer.py:375-384), or None when the
    searcher's modules are not the (sum joint, one Linear, LogSoftmax over the last dim) chain the kernel implements."""
    act = activation_code(getattr(tjoint, "nonlinearity", None))
    lin = _head_linear(classifier_network)
    ok_softmax = isinstance(softmax, torch.nn.LogSoftmax) and softmax.dim in (-1, 3)
    if getattr(tjoint, "joint", None) != "sum" or act is None or lin is None or not ok_softmax:
        return None

    def step(h_i, out_PN):
        with torch.no_grad():
            H = lin.weight.shape[1]
            fp32 = h_i.dtype == out_PN.dtype == lin.weight.dtype == torch.float32
            if not (h_i.is_cuda and lin.weight.is_cuda) or H % 4 != 0 or H > 1536 or not fp32:
                # shapes / dtypes the kernel does not implement (it is fp32 in, fp32 out, H <= 1536): the reference's own chain
                out = tjoint(h_i, out_PN)
                for layer in classifier_network:
                    out = layer(out)
                return softmax(out)
            lp = joint_decode_step(h_i, out_PN, lin.weight, lin.bias, act[0], act[1])
            return lp.view(lp.shape[0], 1, 1, lp.shape[1]) if h_i.dim() == 4 else lp

    return step


_RECURRENT = ("RNN", "LSTM", "GRU", "LiGRU", "LiGRU_Layer")


def _forward_pn(tokens, layers, hidden=None):
    """``TransducerBeamSearcher._forward_PN`` (SB/decoders/transducer.py:468-502): recurrent layers are recognised by
    their class NAME and take / return the hidden state."""
    out = tokens
    for layer in layers:
        if layer.__class__.__name__ in _RECURRENT:
            out, hidden = layer(out, hidden)
        else:
            out = layer(out)
    return out, hidden


def _select_hidden(mask_b, new, old):
    """Per-utterance choice between the freshly computed and the kept hidden state ([layers, B, hidden] tensors, or the
    (h, c) tuple of an LSTM) -- the device-side form of ``_update_hiddens`` (transducer.py:443-466)."""
    if isinstance(old, tuple):
        return tuple(_select_hidden(mask_b, n, o) for n, o in zip(new, old))
    return torch.where(mask_b.view(1, -1, 1), new, old)


def greedy_decode_on_device(tn_output, decode_network_lst, joint_step, blank_id=0):
    """``TransducerBeamSearcher.transducer_greedy_decode`` (SB/decoders/transducer.py:138-218) with the hypothesis
    bookkeeping on the device: no host synchronisation inside the frame loop.

    The reference reads ``positions[i].item()`` for every utterance of every frame (B x T device->host syncs) and runs
    the prediction network on the utterances that emitted a label.  Here the frame loop only queues work: the arg-max,
    the "emitted a non-blank label" mask, the running scores and the [B, T] table of emitted labels stay on the device;
    the prediction network steps ALL utterances and a masked select keeps the old output / hidden state of those that
    emitted blank (rows of a recurrent step are independent, so the kept rows are exactly what the reference keeps).
    ONE device->host copy at the end turns the table into the reference's lists.

    Returns the reference's 4-tuple: (list of label lists, mean of exp(summed log-probs) as a CPU tensor, None, None)."""
    B, T = tn_output.shape[0], tn_output.shape[1]
    dev = tn_output.device
    with torch.no_grad():
        input_pn = torch.full((B, 1), int(blank_id), device=dev, dtype=torch.int32)
        out_pn, hidden = _forward_pn(input_pn, decode_network_lst)                     # transducer.py:172-173
        emitted = torch.empty((T, B), device=dev, dtype=torch.int64)
        scores = torch.zeros((B,), device=dev, dtype=torch.float32)
        for t in range(T):
            log_probs = joint_step(tn_output[:, t, :].unsqueeze(1).unsqueeze(1), out_pn.unsqueeze(1))
            logp, pos = torch.max(log_probs.reshape(B, -1), dim=1)                    # :181-184 (log-softmax is idempotent)
            emit = pos != blank_id                                                    # :189-194
            emitted[t] = pos
            scores = scores + torch.where(emit, logp.to(torch.float32), torch.zeros_like(scores))
            input_pn = torch.where(emit.view(B, 1), pos.to(input_pn.dtype).view(B, 1), input_pn)
            new_out, new_hidden = _forward_pn(input_pn, decode_network_lst, hidden)   # :195-207, all rows
            out_pn = torch.where(emit.view(B, 1, 1), new_out, out_pn)                 # :208-212, rows that emitted
            hidden = _select_hidden(emit, new_hidden, hidden) if hidden is not None else new_hidden
        table = emitted.t().cpu()                                                     # the only device->host copy
        score = scores.exp().mean().cpu()
    hyps = [[int(x) for x in row[row != blank_id].tolist()] for row in table]
    return hyps, score, None, None


def greedy_decode_cuda_graph(tn_output, decode_network_lst, joint_step, blank_id=0):
    """``greedy_decode_on_device`` with the per-frame body captured ONCE in a CUDA graph and replayed T times: a frame is
    one graph launch instead of ~25 kernel launches (the body is launch-bound: tiny kernels over B rows).  The frame
    index lives on the device and is advanced inside the graph, all state is updated in place in static buffers.
    Same results as ``greedy_decode_on_device``; falls back to it when capture is not possible (e.g. a prediction
    network whose kernels cannot be captured)."""
    B, T = tn_output.shape[0], tn_output.shape[1]
    dev = tn_output.device
    if not tn_output.is_cuda or T < 8:
        return greedy_decode_on_device(tn_output, decode_network_lst, joint_step, blank_id)
    try:
        with torch.no_grad():
            tn_t = tn_output.transpose(0, 1).contiguous()                                  # [T, B, H]: one row block per frame
            input_pn0 = torch.full((B, 1), int(blank_id), device=dev, dtype=torch.int32)
            out_pn0, hidden0 = _forward_pn(input_pn0, decode_network_lst)
            hidden0 = hidden0 if isinstance(hidden0, tuple) else (hidden0,)
            state = {"t": torch.zeros((1,), device=dev, dtype=torch.int64), "input_pn": input_pn0.clone(), "out_pn": out_pn0.clone(),
                     "hidden": tuple(h.clone() for h in hidden0), "scores": torch.zeros((B,), device=dev, dtype=torch.float32),
                     "emitted": torch.full((T, B), int(blank_id), device=dev, dtype=torch.int64)}

            def reset():
                state["t"].zero_()
                state["input_pn"].copy_(input_pn0)
                state["out_pn"].copy_(out_pn0)
                for h, h0 in zip(state["hidden"], hidden0):
                    h.copy_(h0)
                state["scores"].zero_()

            def body():
                frame = torch.index_select(tn_t, 0, state["t"]).squeeze(0)                 # [B, H]
                log_probs = joint_step(frame.unsqueeze(1).unsqueeze(1), state["out_pn"].unsqueeze(1))
                logp, pos = torch.max(log_probs.reshape(B, -1), dim=1)
                emit = pos != blank_id
                state["emitted"].index_copy_(0, state["t"], pos.unsqueeze(0))
                state["scores"].add_(torch.where(emit, logp.to(torch.float32), torch.zeros_like(logp, dtype=torch.float32)))
                state["input_pn"].copy_(torch.where(emit.view(B, 1), pos.to(torch.int32).view(B, 1), state["input_pn"]))
                hid = state["hidden"] if len(state["hidden"]) > 1 else state["hidden"][0]
                new_out, new_hidden = _forward_pn(state["input_pn"], decode_network_lst, hid)
                new_hidden = new_hidden if isinstance(new_hidden, tuple) else (new_hidden,)
                state["out_pn"].copy_(torch.where(emit.view(B, 1, 1), new_out, state["out_pn"]))
                for h, nh in zip(state["hidden"], new_hidden):
                    h.copy_(torch.where(emit.view(1, B, 1), nh, h))
                state["t"].add_(1)

            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                                                 # warm-up outside capture (cuDNN plans, allocator)
                for _ in range(3):
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            reset()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            reset()
            for _ in range(T):
                graph.replay()
            table = state["emitted"].t().cpu()
            score = state["scores"].exp().mean().cpu()
    except Exception as ex:  # noqa: BLE001  (capture failed: same search without the graph)
        from .transducer_joint import _warn_once

        _warn_once(f"CUDA-graph capture of the greedy frame body failed ({type(ex).__name__}: {str(ex)[:120]}); using the plain on-device loop")
        return greedy_decode_on_device(tn_output, decode_network_lst, joint_step, blank_id)
    hyps = [[int(x) for x in row[row != blank_id].tolist()] for row in table]
    return hyps, score, None, None


def _hyp_key(h):
    """``get_transducer_key`` (SB/decoders/transducer.py:527-542): length-normalised log-score."""
    return h["logp_score"] / len(h["prediction"])


def _beam_search_one(n_frames, beam_size, nbest, state_beam, expand_beam, blank_id):
    """The reference's per-utterance beam search (SB/decoders/transducer.py:245-352, no LM) as a GENERATOR: wherever the
    reference evaluates the prediction network + joint for ``a_best_hyp`` (:296-309) the generator yields
    ``(frame, last token, hidden state)`` and is sent back ``(top-k log-probs, top-k positions, new hidden)``.
    Scores are float32 scalars added in float32, as the reference's 0-dim CUDA tensors are, so every comparison
    (arg-max over hypotheses, state_beam / expand_beam pruning, the final sort) sees the reference's numbers.
    Returns (n-best label lists, their length-normalised scores)."""
    beam_hyps = [{"prediction": [blank_id], "logp_score": 0.0, "hidden_dec": None}]
    for t_step in range(n_frames):
        process_hyps, beam_hyps = beam_hyps, []                                       # :274-276
        while True:
            if len(beam_hyps) >= beam_size:                                           # :278-279
                break
            a_best = max(process_hyps, key=_hyp_key)                                  # :281-283
            if len(beam_hyps) > 0:                                                    # :286-293
                b_best = max(beam_hyps, key=_hyp_key)
                if b_best["logp_score"] >= state_beam + a_best["logp_score"]:
                    break
            for i, h in enumerate(process_hyps):                                      # :296 (remove THIS hypothesis)
                if h is a_best:
                    del process_hyps[i]
                    break
            logp_targets, positions, hidden = yield t_step, a_best["prediction"][-1], a_best["hidden_dec"]
            best_logp = logp_targets[0] if positions[0] != blank_id else logp_targets[1]  # :320-324
            for j in range(len(logp_targets)):                                        # :327-351
                topk_hyp = {"prediction": a_best["prediction"][:], "logp_score": a_best["logp_score"] + logp_targets[j],
                            "hidden_dec": a_best["hidden_dec"]}
                if positions[j] == blank_id:
                    beam_hyps.append(topk_hyp)
                    continue
                if logp_targets[j] >= best_logp - expand_beam:
                    topk_hyp["prediction"].append(int(positions[j]))
                    topk_hyp["hidden_dec"] = hidden
                    process_hyps.append(topk_hyp)
    best = sorted(beam_hyps, key=_hyp_key, reverse=True)[:nbest]                      # :353-360
    return [h["prediction"][1:] for h in best], [h["logp_score"] / len(h["prediction"]) for h in best]


def _stack_hidden(hiddens, template):
    """Per-row hidden states (each a tensor [layers,1,H], a tuple of such, or None = initial state) -> one batched state
    with the structure of ``template`` (the state a batched call returned); None when every row is at its initial state
    and no template exists yet."""
    if template is None:
        return None  # first evaluation of every utterance: all rows start from the initial state
    def zeros_like_row(t):
        return torch.zeros_like(t[:, :1])
    if isinstance(template, tuple):
        return tuple(torch.cat([(h[k] if h is not None else zeros_like_row(template[k])) for h in hiddens], dim=1)
                     for k in range(len(template)))
    return torch.cat([(h if h is not None else zeros_like_row(template)) for h in hiddens], dim=1)


def _row_hidden(hidden, r):
    if hidden is None:
        return None
    if isinstance(hidden, tuple):
        return tuple(h[:, r:r + 1] for h in hidden)
    return hidden[:, r:r + 1]


def beam_search_batched(tn_output, decode_network_lst, joint_step, blank_id=0, beam_size=4, nbest=5, state_beam=2.3,
                        expand_beam=2.3):
    """``TransducerBeamSearcher.transducer_beam_search_decode`` (SB/decoders/transducer.py:220-373, ``lm_weight == 0``)
    with the B utterances searched CONCURRENTLY: the reference decodes them one after the other and evaluates the
    prediction network and the joint for ONE hypothesis at a time (batch 1, ~15 launches and five ``.item()`` / tensor
    comparisons that synchronise the stream per expansion); here every utterance's search is a coroutine, one round
    evaluates the pending hypothesis of every utterance in one batched prediction-network step and one batched joint
    step, and ONE device->host copy per round brings the top-k back.  Rounds = the longest utterance's expansions
    instead of the sum over utterances.  The searches themselves are the reference's, decision for decision (rows of a
    batched step are independent).  Returns the reference's 4-tuple."""
    B, T = tn_output.shape[0], tn_output.shape[1]
    dev = tn_output.device
    searches = [_beam_search_one(T, beam_size, nbest, state_beam, expand_beam, blank_id) for _ in range(B)]
    pending, results = {}, [None] * B
    for i, g in enumerate(searches):
        try:
            pending[i] = next(g)
        except StopIteration as stop:  # T == 0
            results[i] = stop.value
    template = None
    with torch.no_grad():
        while pending:
            idx = sorted(pending)
            frames = torch.tensor([pending[i][0] for i in idx], device=dev)
            tokens = torch.tensor([[pending[i][1]] for i in idx], device=dev, dtype=torch.int32)
            hidden_in = _stack_hidden([pending[i][2] for i in idx], template)
            out_pn, hidden = _forward_pn(tokens, decode_network_lst, hidden_in)        # :297-303, all utterances at once
            if template is None:
                template = hidden
            h_t = tn_output[torch.tensor(idx, device=dev), frames]                     # [n, H]
            log_probs = joint_step(h_t.unsqueeze(1).unsqueeze(1), out_pn.unsqueeze(1))  # :304-309
            logp, pos = torch.topk(log_probs.reshape(len(idx), -1), k=beam_size, dim=-1)  # :317-319
            logp_h, pos_h = logp.to(torch.float32).cpu().numpy(), pos.cpu().numpy()    # the round's only device->host copy
            for r, i in enumerate(idx):
                try:
                    pending[i] = searches[i].send((logp_h[r], pos_h[r], _row_hidden(hidden, r)))
                except StopIteration as stop:
                    results[i] = stop.value
                    del pending[i]
    nbest_batch = [r[0] for r in results]
    nbest_batch_score = [r[1] for r in results]
    score = torch.Tensor([float(s[0]) for s in nbest_batch_score]).exp().mean()
    return [n[0] for n in nbest_batch], score, nbest_batch, nbest_batch_score


def patch_searcher(searcher, on_device_greedy=False):
    """Replace ``searcher._joint_forward_step`` by the fused step (returns True when patched).
    ``on_device_greedy``: additionally replace the greedy searcher (``beam_size <= 1``) by ``greedy_decode_on_device``
    -- same hypotheses and score, no per-frame host synchronisation; ``"graph"`` selects the CUDA-graph variant (one
    graph launch per frame) -- and the beam searcher (``beam_size > 1``, no LM) by ``beam_search_batched`` (all utterances
    searched concurrently, one batched network evaluation and one device->host copy per expansion round)."""
    step = fused_joint_forward_step(searcher.tjoint, searcher.classifier_network, searcher.softmax)
    if step is None:
        return False
    searcher._joint_forward_step = types.MethodType(lambda self, h_i, out_PN: step(h_i, out_PN), searcher)
    if on_device_greedy and getattr(searcher, "beam_size", 2) <= 1:
        search = greedy_decode_cuda_graph if on_device_greedy == "graph" else greedy_decode_on_device

        def greedy(self, tn_output):
            return search(tn_output, self.decode_network_lst, self._joint_forward_step, self.blank_id)

        searcher.transducer_greedy_decode = types.MethodType(greedy, searcher)
        searcher.searcher = searcher.transducer_greedy_decode
    elif on_device_greedy and getattr(searcher, "lm_weight", 0.0) <= 0:

        def beam(self, tn_output):
            return beam_search_batched(tn_output, self.decode_network_lst, self._joint_forward_step, self.blank_id, self.beam_size,
                                       getattr(self, "nbest", 1), getattr(self, "state_beam", 2.3), getattr(self, "expand_beam", 2.3))

        searcher.transducer_beam_search_decode = types.MethodType(beam, searcher)
        searcher.searcher = searcher.transducer_beam_search_decode
    return True
