"""Data-parallel training step for the MM-DFN hot path.

One process per GPU.  Dialogues are independent units, so each rank runs the whole
forward/backward on its shard of the batch (padded to the GLOBAL max length T, which is the
only cross-dialogue coupling) and the only collective is ONE NCCL all-reduce (sum)
per step over a single flat fp32 gradient bucket, followed by a fused flat-buffer Adam(+L2)
kernel (optim.Adam(lr, weight_decay=l2), code/run_train_erc.py:512).

FocalLoss is a mean over the batch's utterances (code/loss.py:34): each rank scales its local
loss by N_rank / N_global so that the summed gradients equal the single-process gradient even
when shards hold different numbers of utterances.  Parameters the reference never touches on
this path (grad is None there; Adam skips them) stay outside the bucket."""
import torch
import torch.distributed as dist

from ._lib import call, ptr, stream

def used_prefixes(model):
    """Name prefixes of the parameters that receive a gradient for THIS model configuration -- the set the reference's
    Adam actually updates (parameters whose grad stays None are skipped by torch.optim.Adam, weight decay included;
    SURVEY 8b "Used on GDF path").  Depends on graph_type / att_type / use_crn_speaker / reason_flag / layer count."""
    if not getattr(model, "multi_modal", True):
        # single-stream relation model: rnn_parties is constructed but unused (use_crn_speaker is off); with
        # nodal_attention=False the attention's transform gets no gradient either
        pre = ["linear_.", "lstm.", "att_model.scalar.", "graph_net.conv1.", "graph_net.conv2.", "graph_net.linear.",
               "graph_net.smax_fc."]
        if getattr(model, "nodal_attention", True):
            pre.append("graph_net.matchatt.")
        return tuple(pre)
    pre = ["linear_a.", "linear_v.", "linear_l.", "lstm_l.", "smax_fc."]
    if getattr(model, "use_crn_speaker", False):
        pre.append("rnn_parties.")
    if getattr(model, "att_type", "") in ("mfn", "mfn_only"):
        # out_fc1 / out_fc2 of the memory fusion network are constructed but never used: no gradient, Adam skips them
        pre += ["mfn.lstm_", "mfn.att1_", "mfn.att2_", "mfn.gamma1_", "mfn.gamma2_"]
    if getattr(model, "graph_type", "GDF") == "DeepGCN":
        for n in "avl":
            net = getattr(model, "graph_net_" + n)
            pre += [f"graph_net_{n}.fcs.0.", f"graph_net_{n}.convs."]
            if net.reason_flag and len(net.convs) > 0:
                pre.append(f"graph_net_{n}.rnn.")
        if model.att_type == "gated":
            pre.append("gatedatt.")
        return tuple(pre)
    if getattr(model, "graph_type", "GDF") == "None":
        pre += ["graph_net_a.", "graph_net_v.", "graph_net_l."]
        if model.att_type == "gated":
            pre.append("gatedatt.")
        if model.att_type == "lmf_only":
            pre.append("lmf.")
        if model.att_type == "tfn_only":
            pre.append("tfn.")
        return tuple(pre)
    if getattr(model, "graph_type", "GDF") == "relation":
        pre += ["graph_net_a.", "graph_net_v.", "graph_net_l.", "att_model.scalar."]
        if getattr(model, "att_type", "") == "gated":
            pre.append("gatedatt.")
    else:
        net = model.graph_model.graph_net
        pre += ["graph_model.graph_net.fcs.0.", "graph_model.graph_net.convs."]
        if net.reason_flag and len(net.convs) > 0:
            pre.append("graph_model.graph_net.rnn.")
    return tuple(pre)


# the scripted configuration (GDF, crn-speaker encoders, --reason_flag, K > 0): what tests/test_dp_gloo.py shards
USED_PREFIXES = ("linear_a.", "linear_v.", "linear_l.", "lstm_l.", "smax_fc.", "rnn_parties.",
                 "graph_model.graph_net.fcs.0.", "graph_model.graph_net.convs.", "graph_model.graph_net.rnn.")


def used_parameters(model):
    pre = used_prefixes(model)
    return [(n, p) for n, p in model.named_parameters() if n.startswith(pre)]


def shard_dialogues(n_dialogues, rank, world):
    """contiguous shard [lo, hi) of the dialogue axis for this rank"""
    per, rem = divmod(n_dialogues, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def gradient_groups(model):
    """[(sink key, [parameters in the order of the backward function's internal gradient buffer])] for the GDF model,
    early-final groups first: the head and graph-stack gradients are complete ~1 ms before the encoder gradients, so
    they form the first all-reduce bucket.  None for configurations whose functions do not write into a sink."""
    if getattr(model, "graph_type", None) not in ("GDF", "GF") or getattr(model, "att_type", "") == "mfn":
        return None
    from . import ops
    net = model.graph_model.graph_net
    gcn = [net.fcs[0].weight, net.fcs[0].bias]
    if net.reason_flag and len(net.convs) > 0:
        gcn += [net.rnn.weight_ih_l0, net.rnn.weight_hh_l0, net.rnn.bias_ih_l0, net.rnn.bias_hh_l0]
    gcn += [c.weight for c in net.convs]
    groups = [("head", [model.smax_fc.weight, model.smax_fc.bias]), ("gcn", gcn)]
    if model.use_crn_speaker:
        groups.append(("gru_p", [getattr(model.rnn_parties, ops.GRU_KEYS[i]) for i in ops._GRU_GRAD_ORDER]))
    groups.append(("gru_l", [getattr(model.lstm_l, ops.GRU_KEYS[i]) for i in ops._GRU_GRAD_ORDER]))
    groups.append(("proj", [model.linear_a.weight, model.linear_a.bias, model.linear_v.weight, model.linear_v.bias,
                            model.linear_l.weight, model.linear_l.bias]))
    return groups


EARLY_KEYS = ("head", "gcn")          # first all-reduce bucket (final when the graph stack's backward returns)


class FlatAdamTrainer:
    """Flat parameter / gradient buckets + all-reduce + one fused Adam launch per step.

    GDF models run in *direct* mode: the backward functions write every weight gradient straight into this trainer's
    bucket (ops.GradSink; the bucket is zeroed by one memset per step, nothing is gathered afterwards), and the
    all-reduce is split in two: the head + graph-stack segment is reduced on NCCL's stream as soon as the stack's
    backward has been enqueued, i.e. while adjacency / party-pack / GRU / projection backward kernels still run; the
    encoder segment follows at the end.  Other configurations (relation graph type) use the gather path."""

    def __init__(self, model, loss_fn, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 direct_grads=None):
        self.model, self.loss_fn = model, loss_fn
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        named = used_parameters(model)
        groups = gradient_groups(model) if direct_grads is not False else None
        if groups is not None:
            # bucket order = group order; must cover exactly the used parameters
            order = [p for _, ps in groups for p in ps]
            if sorted(id(p) for p in order) != sorted(id(p) for _, p in named):
                groups = None
        if direct_grads and groups is None:
            raise ValueError("direct_grads needs a GDF model")
        name_of = {id(p): n for n, p in named}
        params = [p for _, ps in groups for p in ps] if groups is not None else [p for _, p in named]
        self.names = [name_of[id(p)] for p in params]
        dev = params[0].device
        # bucket layout: parameters back to back; in direct mode every gradient group starts on a 256-byte boundary (the
        # split-K weight-gradient kernels reduce into the bucket with 128-bit atomics, which need 16-byte aligned rows --
        # the head's C*900 + C floats would otherwise shift every later group off alignment).  Pad floats stay zero in
        # the parameter, gradient and moment buckets, so Adam leaves them at zero.
        ALIGN = 64
        offs, seg_range, off = [], {}, 0
        if groups is not None:
            for key, ps in groups:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
                start = off
                for q in ps:
                    offs.append(off)
                    off += q.numel()
                seg_range[key] = (start, off)
        else:
            for q in params:
                offs.append(off)
                off += q.numel()
        total = off
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.params, self.grad_views = params, []
        for q, off in zip(params, offs):
            n = q.numel()
            self.flat_p[off:off + n].copy_(q.data.reshape(-1))
            q.data = self.flat_p[off:off + n].view_as(q)          # parameters become views of the bucket
            self.grad_views.append(self.flat_g[off:off + n].view_as(q))
            q.grad = None
        self.total = total
        self.sink = None
        if groups is not None:
            from . import ops
            self.sink = ops.GradSink()
            for key, (a, b) in seg_range.items():
                self.sink.seg[key] = self.flat_g[a:b]
            self.early = max(seg_range[k][1] for k in EARLY_KEYS)                 # bucket A = flat_g[:early]
            self.sink.on_ready = self._group_ready
        self._early_work = None
        self.step_count = 0
        self._state = None            # device-resident step state (graph mode, see capture())
        self._graph = None            # most recently captured step
        self._graphs = {}             # lengths tuple -> (graph, static inputs, static loss, n_global)
        in_bucket = {id(p) for p in params}
        self._outside = [(n, p) for n, p in model.named_parameters() if p.requires_grad and id(p) not in in_bucket]
        self._checked = False
        if self.world > 1:                                        # replicas start identical
            dist.broadcast(self.flat_p, src=0, group=self.pg)

    def _group_ready(self, key):
        """called from the backward when a gradient group is final: after the graph stack (the head came earlier), start
        reducing bucket A on NCCL's stream -- it overlaps the rest of the backward"""
        if key == "gcn" and self.world > 1 and self._early_work is None:
            self._early_work = dist.all_reduce(self.flat_g[:self.early], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def step(self, textf, qmask, umask, lengths, acouf, visuf, label, n_global=None):
        """One fwd + bwd + (all-reduce) + Adam step on this rank's shard.  Returns the local loss tensor
        (already scaled by N_rank/N_global; the sum over ranks is the global mean loss)."""
        for p in self.params:
            p.grad = None                                         # autograd then stores each gradient without an add kernel
        from . import ops
        if self.sink is not None:
            call("mmdfn_memset_zero", self.flat_g.data_ptr(), self.total * 4, stream())     # the step's ONE gradient zero-fill
            self._early_work = None
        with ops.use_sink(self.sink):
            log_prob = self.model(textf, qmask, umask, lengths, acouf, visuf)[0]
            self.last_log_prob = log_prob.detach()                # for the trainer loop's device-side metrics
            loss = self.loss_fn(log_prob, label)
            n_local = int(sum(lengths))
            if n_global is not None and n_global != n_local:
                loss = loss * (float(n_local) / float(n_global))
            loss.backward()
        if not self._checked:
            # once per trainer: the bucket must hold exactly the parameters that receive a gradient -- a parameter
            # outside it would silently stay at its initial value, one inside it without a gradient would be decayed
            self._checked = True
            stray = [n for n, p in self._outside if p.grad is not None]
            missing = [] if self.sink is not None else [n for n, p in zip(self.names, self.params) if p.grad is None]
            leaked = [n for n, p in zip(self.names, self.params) if p.grad is not None] if self.sink is not None else []
            if stray or missing or leaked:
                raise RuntimeError(f"FlatAdamTrainer bucket mismatch: gradients outside the bucket {stray}, bucket parameters "
                                   f"without a gradient {missing}, direct-mode parameters that still got an autograd gradient {leaked}")
        if self.sink is None:
            torch._foreach_copy_(self.grad_views, [p.grad for p in self.params])   # one multi-tensor gather into the bucket
            if self.world > 1:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
        elif self.world > 1:
            if self._early_work is not None:
                dist.all_reduce(self.flat_g[self.early:], op=dist.ReduceOp.SUM, group=self.pg)     # bucket B: encoder gradients
                self._early_work.wait()                           # current stream waits for bucket A
                self._early_work = None
            else:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
        if self._state is not None:
            # graph mode: the step index and Adam's bias corrections live on the device (captured launches cannot
            # take per-step arguments)
            call("mmdfn_step_advance", ptr(self._state, torch.int64), float(self.betas[0]), float(self.betas[1]), stream())
            call("mmdfn_adam_step_dev", self.total, ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg),
                 ptr(self.exp_avg_sq), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                 float(self.wd), ptr(self._state, torch.int64), 1.0, stream())
            if not torch.cuda.is_current_stream_capturing():
                self.step_count += 1
            return loss.detach()
        self.step_count += 1
        call("mmdfn_adam_step", self.total, ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd),
             self.step_count, 1.0, stream())
        return loss.detach()

    # ---- whole-step CUDA graph -------------------------------------------------------------------------------------
    def capture(self, textf, qmask, umask, lengths, acouf, visuf, label, n_global=None, warmup=1):
        """Capture fwd + bwd + (all-reduce) + Adam for batches of exactly these shapes and lengths into ONE CUDA graph.
        Afterwards `replay(textf, qmask, umask, acouf, visuf, label)` runs a step with a single launch: the ~130 kernel
        launches and their inter-kernel gaps disappear from the host and shrink on the device.  Per-step quantities
        (dropout counter, Adam step index / bias corrections) are read from a device-resident state, so replays draw fresh
        masks and apply the right corrections.  The eager `step` keeps working for other shapes (it shares the state)."""
        from . import ops
        from .modules import _geom_of
        dev = self.flat_p.device
        _geom_of(lengths, dev)                  # host -> device copies of the geometry happen here, not inside the capture
        self._static = tuple(torch.empty_like(x) for x in (textf, qmask, umask, acouf, visuf, label))
        for dst, src in zip(self._static, (textf, qmask, umask, acouf, visuf, label)):
            dst.copy_(src)
        self._lengths, self._n_global = [int(x) for x in lengths], n_global
        if self._state is None:                 # one device-resident step state shared by every captured geometry
            self._state = torch.zeros(2, dtype=torch.int64, device=dev)
            self._state[0] = self.step_count
        ops.set_step_state(self._state)
        t, q, u, a, v, lab = self._static
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the default stream, as graph capture requires
            for _ in range(max(0, warmup)):      # real training steps on this batch (0 if eager steps ran before)
                self.step(t, q, u, self._lengths, a, v, lab, n_global)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self.step(t, q, u, self._lengths, a, v, lab, n_global)
        torch.cuda.synchronize(dev)
        self._graphs[tuple(self._lengths)] = (self._graph, self._static, self._static_loss, n_global)
        return self

    def release_graphs(self):
        """drop every captured step (needed before tearing down an NCCL communicator the graphs reference)"""
        for g, _, _, _ in self._graphs.values():
            g.reset()
        self._graphs.clear()
        self._graph = None

    def has_graph(self, lengths):
        return tuple(int(x) for x in lengths) in self._graphs

    def replay(self, textf, qmask, umask, acouf, visuf, label, lengths=None):
        """One captured step on a new batch of a captured geometry; returns the (static) loss tensor.  The captured
        graph bakes in the dialogue lengths (row offsets, adjacency block offsets, label count, loss scale), not just the
        padded shapes: pass `lengths` to select the graph captured for exactly these lengths (one graph per batch
        geometry); without it the most recent capture is used and the caller vouches for identical lengths."""
        if self._graph is None:
            raise RuntimeError("replay() needs a captured step: call capture() first")
        if lengths is not None:
            key = tuple(int(x) for x in lengths)
            if key not in self._graphs:
                raise ValueError("replay(): no step was captured for these dialogue lengths (capture one graph per batch geometry)")
            self._graph, self._static, self._static_loss, self._n_global = self._graphs[key]
            self._lengths = list(key)
        for dst, src in zip(self._static, (textf, qmask, umask, acouf, visuf, label)):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"replay() got a tensor of shape {tuple(src.shape)} / {src.dtype}; the captured step takes "
                                 f"{tuple(dst.shape)} / {dst.dtype} (capture one graph per batch geometry)")
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.step_count += 1
        return self._static_loss
