"""torch.autograd bindings of the C ABI (include/mmdfn_b200.h).

PyTorch is used for device memory, streams and the autograd tape only; every arithmetic
step of the hot path is a kernel of libmmdfn_b200.so.  Nothing here falls back to torch
ops or to the CPU."""
import ctypes
import itertools
import os

import torch

from ._lib import MMDFNError, call, ptr, ptr_table, query, stream

F32 = torch.float32
U8 = torch.uint8


def _empty(shape, device, dtype=F32):
    return torch.empty(shape, device=device, dtype=dtype)


def _f32c(t):
    """contiguous fp32 CUDA view of t (no copy when already so)."""
    if not t.is_cuda:
        raise MMDFNError("mmdfn_b200 needs CUDA tensors: the hot path has no CPU fallback")
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


def _zeros(n, device, dtype=F32):
    """zero-filled buffer: torch.empty + one cudaMemsetAsync (a memset node, not a fill kernel)"""
    t = torch.empty((int(n),), device=device, dtype=dtype)
    if n:
        call("mmdfn_memset_zero", t.data_ptr(), int(n) * t.element_size(), stream())
    return t


# ---------------------------------------------------------------------------------------------
# gradient sink: the data-parallel trainer (dp.FlatAdamTrainer) hands every weight-gradient group of the model a
# segment of its flat bucket.  Inside `with use_sink(sink)` the backward functions below write their weight gradients
# straight into those segments (the bucket is zeroed once per step) and return None for them: no per-function zero
# fills, no gather copy before the all-reduce.  `sink.ready(key)` tells the trainer that a group's gradients are final,
# so that it can start reducing them while the rest of the backward still runs.
# ---------------------------------------------------------------------------------------------
class GradSink:
    def __init__(self):
        self.seg = {}            # key -> flat fp32 view of the bucket, laid out as the function's internal buffer
        self.on_ready = None

    def ready(self, key):
        if self.on_ready is not None:
            self.on_ready(key)


_SINK = [None]
_SINK_KEY = [None]


class use_sink:
    def __init__(self, sink):
        self.sink, self.prev = sink, None

    def __enter__(self):
        self.prev, _SINK[0] = _SINK[0], self.sink

    def __exit__(self, *exc):
        _SINK[0] = self.prev


class sink_key:
    """names the gradient group of the Function.apply calls issued inside the block (read in their forward)"""

    def __init__(self, key):
        self.key, self.prev = key, None

    def __enter__(self):
        self.prev, _SINK_KEY[0] = _SINK_KEY[0], self.key

    def __exit__(self, *exc):
        _SINK_KEY[0] = self.prev


def _grad_buffer(key, n, device):
    """(flat buffer of n floats for a function's weight gradients, direct?) -- the sink's segment when one is active"""
    sink = _SINK[0]
    if sink is not None and key is not None and key in sink.seg:
        seg = sink.seg[key]
        if seg.numel() != n:
            raise MMDFNError(f"gradient sink segment {key!r} holds {seg.numel()} floats, the function needs {n}")
        return seg, True
    return _zeros(n, device), False


class DialogGeom:
    """Ragged geometry of one batch: dialogue i owns rows dia_off[i]..dia_off[i+1]-1 of each
    modality third of every (3N, .) array and 3*L_i^2 floats of the block-compact adjacency."""

    def __init__(self, lengths, device):
        self.lengths = [int(x) for x in lengths]
        self.B = len(self.lengths)
        self.N = int(sum(self.lengths))
        self.Lmax = int(max(self.lengths)) if self.lengths else 0
        off = [0] + list(itertools.accumulate(self.lengths))
        blk = [0] + list(itertools.accumulate(3 * L * L for L in self.lengths))
        self.nblk = blk[-1]
        self.device = device
        self.dia_off = torch.tensor(off, dtype=torch.int32).to(device)
        self.blk_off = torch.tensor(blk, dtype=torch.int64).to(device)

    def args(self):
        return (self.B, self.N, self.Lmax, ptr(self.dia_off, torch.int32), ptr(self.blk_off, torch.int64))


# ---------------------------------------------------------------------------------------------
# dropout keep-masks (counter based; one launch per mask)
# ---------------------------------------------------------------------------------------------
_mask_counter = [0]


def make_mask(shape, p, device):
    return make_masks([tuple(shape)], p, device)[0]


# CUDA-graph mode: device-resident step state (int64[2], see mmdfn_step_advance).  When set, make_masks draws with a
# counter base that advances with the device-side step index, so every replay of a captured step gets fresh masks.
_STEP_STATE = [None]


def set_step_state(state):
    """state: int64[2] CUDA tensor (or None to leave graph mode)."""
    _STEP_STATE[0] = state


def make_masks(shapes, p, device):
    """Several keep-masks with ONE launch: one uint8 buffer, views in request order.  Element i of the concatenation
    uses counter base + i, so the bits equal those of consecutive make_mask calls in the same order."""
    sizes = []
    for shape in shapes:
        n = 1
        for s in shape:
            n *= int(s)
        sizes.append(n)
    total = sum(sizes)
    buf = _empty((total,), device, U8)
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    st = _STEP_STATE[0]
    if total and st is not None:
        # counter base = this call's offset + (steps taken so far) * 2^40: distinct per call and per replayed step
        call("mmdfn_dropout_mask_dev", total, float(p), seed, ptr(st, torch.int64), 1 << 40, _mask_counter[0], ptr(buf, U8),
             stream())
    elif total:
        call("mmdfn_dropout_mask", total, float(p), seed, _mask_counter[0], ptr(buf, U8), stream())
    _mask_counter[0] = (_mask_counter[0] + total) & 0xFFFFFFFFFFFFFFFF
    out, off = [], 0
    for shape, n in zip(shapes, sizes):
        out.append(buf[off:off + n].view(*shape))
        off += n
    return out


# ---------------------------------------------------------------------------------------------
# k1: projections
# ---------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b over the last dim (nn.Linear)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = _f32c(x).reshape(-1, x.shape[-1])
        w = _f32c(w)
        rows, K = x2.shape
        n_out = w.shape[0]
        y = _empty((rows, n_out), x.device)
        call("mmdfn_gemm", 0, 1, rows, n_out, K, 1.0, ptr(x2), K, ptr(w), K, 0.0, ptr(y), n_out,
             ptr(_f32c(b)) if b is not None else None, 0, stream())
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], n_out)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        rows, K = x2.shape
        n_out = w.shape[0]
        dy2 = _f32c(dy).reshape(rows, n_out)
        dx = dw = db = None
        st = stream()
        if ctx.needs_input_grad[0]:
            dx = _empty((rows, K), dy.device)
            call("mmdfn_gemm", 0, 0, rows, K, n_out, 1.0, ptr(dy2), n_out, ptr(w), K, 0.0, ptr(dx), K, None, 0, st)
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = _empty((n_out, K), dy.device)
            call("mmdfn_gemm", 1, 0, n_out, K, rows, 1.0, ptr(dy2), n_out, ptr(x2), K, 0.0, ptr(dw), K, None, 0, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _empty((n_out,), dy.device)
            call("mmdfn_colsum", rows, n_out, ptr(dy2), n_out, 0.0, ptr(db), st)
        return dx, dw, db


class MaskScaleFn(torch.autograd.Function):
    """y = mask ? x * scale : 0  (nn.Dropout with an explicit uint8 keep mask)."""

    @staticmethod
    def forward(ctx, x, mask, scale):
        x = _f32c(x)
        y = _empty(x.shape, x.device)
        call("mmdfn_mask_scale", x.numel(), ptr(x), ptr(mask, U8), float(scale), ptr(y), stream())
        ctx.mask, ctx.scale = mask, float(scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _f32c(dy)
        dx = _empty(dy.shape, dy.device)
        call("mmdfn_mask_scale", dy.numel(), ptr(dy), ptr(ctx.mask, U8), ctx.scale, ptr(dx), stream())
        return dx, None, None


class LogSoftmaxFn(torch.autograd.Function):
    """log_softmax over dim 1 of (N, C) logits."""

    @staticmethod
    def forward(ctx, logits):
        logits = _f32c(logits)
        N, C = logits.shape
        lp = _empty((N, C), logits.device)
        call("mmdfn_log_softmax_fwd", N, C, ptr(logits), ptr(lp), stream())
        ctx.save_for_backward(lp)
        return lp

    @staticmethod
    def backward(ctx, dlp):
        (lp,) = ctx.saved_tensors
        N, C = lp.shape
        dlp = _f32c(dlp)
        dlogits = _empty((N, C), lp.device)
        call("mmdfn_log_softmax_bwd", N, C, ptr(lp), ptr(dlp), ptr(dlogits), stream())
        return dlogits


class GatedFuseFn(torch.autograd.Function):
    """MMGatedAttention 'general' after the three projections (code/model.py:761-781): gates from the inputs, tanh of
    the projections, the three pairwise mixes.  w (3, 3D) / b (3): transform_av/al/vl stacked."""

    @staticmethod
    def forward(ctx, xa, xv, xl, Pa, Pv, Pl, w, b):
        xa, xv, xl, Pa, Pv, Pl, w, b = (_f32c(t) for t in (xa, xv, xl, Pa, Pv, Pl, w, b))
        N, D = xa.shape
        C = Pa.shape[1]
        out = _empty((N, 3 * C), xa.device)
        z = _empty((N, 3), xa.device)
        call("mmdfn_gated_fuse_fwd", N, D, C, ptr(xa), ptr(xv), ptr(xl), ptr(Pa), ptr(Pv), ptr(Pl), ptr(w), ptr(b),
             ptr(out), ptr(z), stream())
        ctx.save_for_backward(xa, xv, xl, Pa, Pv, Pl, w, b, z)
        return out

    @staticmethod
    def backward(ctx, dout):
        xa, xv, xl, Pa, Pv, Pl, w, b, z = ctx.saved_tensors
        N, D = xa.shape
        C = Pa.shape[1]
        dout = _f32c(dout)
        dev = xa.device
        dP = [_empty((N, C), dev) for _ in range(3)]
        dx = [_empty((N, D), dev) for _ in range(3)]
        dw, db, ws = _empty((3, 3 * D), dev), _empty((3,), dev), _empty((N, 3), dev)
        call("mmdfn_gated_fuse_bwd", N, D, C, ptr(dout), ptr(xa), ptr(xv), ptr(xl), ptr(Pa), ptr(Pv), ptr(Pl), ptr(w), ptr(b),
             ptr(z), ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), ptr(dx[0]), ptr(dx[1]), ptr(dx[2]), ptr(dw), ptr(db), ptr(ws),
             stream())
        return dx[0], dx[1], dx[2], dP[0], dP[1], dP[2], dw, db


class Proj3Fn(torch.autograd.Function):
    """U (3,T,B,200) = stack over (a, v, l) of x_m W_m^T + b_m   (code/model.py:1065,1094,1129)."""
    text_ready = None          # event recorded after the text slice U[2] of the most recent forward was enqueued

    @staticmethod
    def forward(ctx, xa, xv, xl, wa, ba, wv, bv, wl, bl):
        xs = [_f32c(x) for x in (xa, xv, xl)]
        ws = [_f32c(w) for w in (wa, wv, wl)]
        bs = [_f32c(b) for b in (ba, bv, bl)]
        T, B = xs[0].shape[0], xs[0].shape[1]
        rows = T * B
        U = _empty((3, T, B, 200), xa.device)
        st = stream()
        # the text projection first: the text encoder (the longer of the two concurrent encoder chains) depends on it alone
        # and may start on its side stream while the audio / visual projections still run (event: Proj3Fn.text_ready)
        for m in (2, 0, 1):
            K = xs[m].shape[2]
            xa_, wa_, Kp = xs[m], ws[m], K
            if K % 4 and rows * K >= (1 << 20):
                # a feature width that is not a multiple of 4 floats (IEMOCAP audio: 1582) leaves rows off 16-byte
                # alignment, which the TMA-fed tensor-core GEMM cannot take: zero-pad both operands' K (two copies, ~10 us)
                Kp = (K + 3) // 4 * 4
                xa_ = torch.nn.functional.pad(xs[m].reshape(rows, K), (0, Kp - K))
                wa_ = torch.nn.functional.pad(ws[m], (0, Kp - K))
            call("mmdfn_gemm", 0, 1, rows, 200, Kp, 1.0, ptr(xa_), Kp, ptr(wa_), Kp, 0.0,
                 U.data_ptr() + m * rows * 200 * 4, 200, ptr(bs[m]), 0, st)
            if m == 2:
                Proj3Fn.text_ready = torch.cuda.Event()
                Proj3Fn.text_ready.record(torch.cuda.current_stream(xa.device))
        ctx.save_for_backward(*xs, *ws)
        ctx.sink_key = _SINK_KEY[0]
        return U

    @staticmethod
    def backward(ctx, dU):
        xs, ws = ctx.saved_tensors[:3], ctx.saved_tensors[3:]
        dU = _f32c(dU)
        T, B = xs[0].shape[0], xs[0].shape[1]
        rows = T * B
        st = stream()
        out = [None] * 9
        Ks = [x.shape[2] for x in xs]
        flat, direct = _grad_buffer(ctx.sink_key, 200 * sum(Ks) + 600, dU.device)      # all six gradients, one memset
        off = 0
        for m in range(3):
            K = Ks[m]
            g = dU.data_ptr() + m * rows * 200 * 4
            if ctx.needs_input_grad[m]:
                dx = _empty(xs[m].shape, dU.device)
                call("mmdfn_gemm", 0, 0, rows, K, 200, 1.0, g, 200, ptr(ws[m]), K, 0.0, ptr(dx), K, None, 0, st)
                out[m] = dx
            dw = flat[off:off + 200 * K].view(200, K)
            db = flat[off + 200 * K:off + 200 * K + 200]
            off += 200 * K + 200
            call("mmdfn_gemm", 1, 0, 200, K, rows, 1.0, g, 200, ptr(xs[m]), K, 1.0, ptr(dw), K, None, 0, st)
            call("mmdfn_colsum", rows, 200, g, 200, 1.0, ptr(db), st)
            if not direct:
                out[3 + 2 * m], out[4 + 2 * m] = dw, db
        if direct:
            _SINK[0].ready(ctx.sink_key)
        return tuple(out)


# ---------------------------------------------------------------------------------------------
# k2: 2-layer bidirectional GRU
# ---------------------------------------------------------------------------------------------
GRU_KEYS = [f"{k}_l{l}{sfx}" for l in (0, 1) for sfx in ("", "_reverse")
            for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]


# placement order of the 16 GRU gradients inside their shared buffer: [ih_l0, ih_l0_rev, hh_l0, hh_l0_rev, biases, l1 ...]
_GRU_GRAD_ORDER = [0, 4, 1, 5, 2, 3, 6, 7, 8, 12, 9, 13, 10, 11, 14, 15]


_GRU_TILE = [0]      # sequences per CTA for the BiGRU2Fn calls issued inside a `gru_tile(nb)` block (0 = library default)


class gru_tile:
    """Context manager: recurrence tile (sequences per CTA: 2, 3, 4, 8; 0 = automatic) for the BiGRU2Fn.apply calls in
    the block, remembered by each call for its backward.  Used by the model to fit two concurrently running encoders
    into one wave of CTAs (mmdfn_gru_set_tile)."""

    def __init__(self, nb):
        self.nb, self.prev = int(nb), 0

    def __enter__(self):
        self.prev, _GRU_TILE[0] = _GRU_TILE[0], self.nb

    def __exit__(self, *exc):
        _GRU_TILE[0] = self.prev


def plan_gru_tiles(T, n_a, n_b, sms=148):
    """Tiles (nb_a, nb_b) for two 2-layer BiGRU encoders with n_a and n_b sequences that run concurrently: minimise the
    longer of the two, subject to both grids (2 * ceil(n / nb) CTAs each) being co-resident.  Per-step cost model from
    clock64 stamps on B200 (profiles/r01_gru_phase_stamps_s2.log): 260 + 750 nb cycles of mat-vec + 600 per pointwise
    pass of 320 items; the hoisted input GEMMs are charged at 40 TFLOP/s.  Returns (0, 0) if nothing fits."""
    def enc_us(n, nb, rows_l0):
        step = 260 + 750 * nb + 600 * -(-(nb * 100) // 320)
        gemm = 2.0 * 600 * 200 * (rows_l0 + T * n) / 40e12 * 1e6
        return gemm + 2 * T * step / 1965.0
    best, pick = None, (0, 0)
    for na in (2, 3, 4, 8):
        for nb in (2, 3, 4, 8):
            if 2 * -(-n_a // na) + 2 * -(-n_b // nb) > sms:
                continue
            ta, tb = enc_us(n_a, na, T * n_a), enc_us(n_b, nb, T * n_b / 2)
            key = (max(ta, tb), ta + tb)
            if best is None or key < best:
                best, pick = key, (na, nb)
    return pick


def _wgrad_stream(device, idx=0, cache={}):
    key = (str(device), idx)
    if key not in cache:
        cache[key] = torch.cuda.Stream(device=device)
    return cache[key]


# Where the trainer-mode weight-gradient contractions of the encoders run (measured at the bench shard, whole-step graph):
# 0: one stream for all of them -- its chain of 12 GEMMs was the tail of the step (2.105 ms);
# 1: one stream per encoder call, the two chains side by side (2.067 ms) -- the default;
# 2: as 1, and the layer-1 contractions are enqueued right behind the layer-1 recurrence (2.071 ms: the two concurrent
#    recurrences hold 144 of the 148 SMs, so nothing runs in their shadow and the early GEMMs only delay the layer-0 launch)
_WGRAD_MODE = [int(os.environ.get("MMDFN_WGRAD_MODE", "1"))]
_WGRAD_JOIN = [False]
_WGRAD_KEEP = []
_WGRAD_NEXT = [0]
_WGRAD_NSTREAMS = 2


def _join_wgrad_stream(device):
    """once per backward pass: when autograd has run its last node, the stream that called backward() waits for the
    weight-gradient streams (the optimizer / all-reduce / a captured graph's end come after that)"""
    if _WGRAD_JOIN[0]:
        return
    _WGRAD_JOIN[0] = True

    def join():
        _WGRAD_JOIN[0] = False
        cur = torch.cuda.current_stream(device)
        for i in range(min(_WGRAD_NEXT[0], _WGRAD_NSTREAMS)):      # only the streams this pass used (a captured step may
            cur.wait_stream(_wgrad_stream(device, i))               # not wait for a stream outside its capture)
        _WGRAD_NEXT[0] = 0
        _WGRAD_KEEP.clear()               # freed now: reused only by work ordered after this wait

    torch.autograd.Variable._execution_engine.queue_callback(join)


class BiGRU2Fn(torch.autograd.Function):
    """x (rows, in_dim) [+ rowmap (T,nseq) gather] -> y (T,nseq,200).  16 weights in GRU_KEYS order; in_dim = 200 on the
    MM-DFN path, the width of `linear_` in the text-only configuration."""

    @staticmethod
    def forward(ctx, x, rowmap, T, nseq, mask, mask_scale, *w):
        x = _f32c(x)
        w = [_f32c(t) for t in w]
        rows = x.shape[0]
        y = _empty((T, nseq, 200), x.device)
        ws = _empty((query("mmdfn_bigru2_ws_floats", T, nseq, rows),), x.device)
        tab = ptr_table(w)
        ctx.tile = _GRU_TILE[0]
        ctx.sink_key = _SINK_KEY[0]
        call("mmdfn_gru_set_tile", ctx.tile)
        try:
            call("mmdfn_bigru2_fwd_in", x.shape[1], T, nseq, rows, ptr(x), ptr(rowmap, torch.int32), tab, ptr(mask, U8),
                 float(mask_scale), ptr(y), ptr(ws), stream())
        finally:
            call("mmdfn_gru_set_tile", 0)
        ctx.save_for_backward(x, y, ws, *w)
        ctx.rowmap, ctx.mask, ctx.mask_scale, ctx.T, ctx.nseq = rowmap, mask, float(mask_scale), T, nseq
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, ws = ctx.saved_tensors[:3]
        w = list(ctx.saved_tensors[3:])
        T, nseq, rows = ctx.T, ctx.nseq, x.shape[0]
        dy = _f32c(dy)
        dx = _empty(x.shape, x.device) if ctx.needs_input_grad[0] else None
        # all 16 gradients live in ONE zero-filled buffer (a single memset instead of per-GEMM zero-init launches);
        # weight_ih of the two directions of a layer are adjacent so that one GEMM produces both
        flat, direct = _grad_buffer(ctx.sink_key, sum(t.numel() for t in w), x.device)
        dw, off = [None] * 16, 0
        for i in _GRU_GRAD_ORDER:
            n = w[i].numel()
            dw[i] = flat[off:off + n].view(w[i].shape)
            off += n
        wsb = _empty((query("mmdfn_bigru2_bwd_ws_floats", T, nseq, rows),), x.device)
        tab, dtab = ptr_table(w), ptr_table(dw)
        xd = x.shape[1]
        rm, mk = ptr(ctx.rowmap, torch.int32), ptr(ctx.mask, U8)
        dev = x.device

        def data(parts):
            # the dependency chain of the step: recurrences, the input gradients between and after them
            call("mmdfn_gru_set_tile", ctx.tile)
            try:
                call("mmdfn_bigru2_bwd_data_part", parts, xd, T, nseq, rows, ptr(x), rm, tab, mk, ctx.mask_scale, ptr(y), ptr(dy),
                     ptr(ws), ptr(dx), 0, dtab, 1, ptr(wsb), stream())
            finally:
                call("mmdfn_gru_set_tile", 0)

        def wgrad(parts):
            call("mmdfn_bigru2_bwd_wgrad_part", parts, xd, T, nseq, rows, ptr(x), rm, mk, ptr(y), ptr(ws), dtab, 1, ptr(wsb), stream())

        if not direct:
            # gradients returned to autograd must be complete on the current stream
            data(3)
            wgrad(3)
            return (dx, None, None, None, None, None, *dw)
        # trainer mode (gradients go straight into the flat bucket): the weight-gradient contractions only feed the optimizer,
        # so they run on their own stream behind an event, off the critical stream, and are joined when the backward pass
        # ends (_join_wgrad_stream) -- before the all-reduce and Adam
        mode = _WGRAD_MODE[0]
        cur = torch.cuda.current_stream(dev)
        wst = _wgrad_stream(dev, _WGRAD_NEXT[0] % _WGRAD_NSTREAMS if mode else 0)
        _WGRAD_NEXT[0] += 1 if mode or not _WGRAD_NEXT[0] else 0

        def wgrad_behind(parts):
            ev = torch.cuda.Event()
            ev.record(cur)
            wst.wait_event(ev)
            with torch.cuda.stream(wst):
                wgrad(parts)

        if mode >= 2:
            data(2)
            wgrad_behind(2)
            data(1)
            wgrad_behind(1)
        else:
            data(3)
            wgrad_behind(3)
        # the buffers the other stream still reads stay referenced until the join (autograd drops this node's saved tensors as
        # soon as it returns; record_stream would do, but its deferred frees made the caching allocator grow in eager loops)
        _WGRAD_KEEP.append((x, y, ws, wsb, flat, dy))
        _join_wgrad_stream(dev)
        _SINK[0].ready(ctx.sink_key)
        return (dx, None, None, None, None, None, *([None] * 16))


# ---------------------------------------------------------------------------------------------
# k3/k4: speaker partition (integer) and fused scatter + combine + ragged pack
# ---------------------------------------------------------------------------------------------
def spk_partition(qmask, want_rowmap=True):
    """qmask (T,B,S) -> pos (T,B,S), cnt (B,S), sel (T,B), rowmap (T, 3*B*S) int32."""
    qmask = _f32c(qmask)
    T, B, S = qmask.shape
    dev = qmask.device
    pos = _empty((T, B, S), dev, torch.int32)
    cnt = _empty((B, S), dev, torch.int32)
    sel = _empty((T, B), dev, torch.int32)
    rowmap = _empty((T, 3 * B * S), dev, torch.int32) if want_rowmap else None
    call("mmdfn_spk_partition", T, B, S, ptr(qmask), ptr(pos, torch.int32), ptr(cnt, torch.int32),
         ptr(sel, torch.int32), ptr(rowmap, torch.int32), stream())
    return pos, cnt, sel, rowmap


class PartyPackFn(torch.autograd.Function):
    """X (3N,200): rows (m, off_b + t) = base_m[t,b] + w_m * Q[pos[t,b,sel], (m*B+b)*S+sel]; bases are
    U[0], U[1] (raw projections) and E_l (text BiGRU output)."""

    @staticmethod
    def forward(ctx, U, E_l, Q, geom, sel, pos, S, weights):
        U, E_l = _f32c(U), _f32c(E_l)
        Q = _f32c(Q) if Q is not None else None
        _, T, B, _ = U.shape
        X = _empty((3 * geom.N, 200), U.device)
        rows = T * B * 200 * 4
        call("mmdfn_party_pack_fwd", T, B, S, geom.N, ptr(geom.dia_off, torch.int32), ptr(sel, torch.int32),
             ptr(pos, torch.int32), U.data_ptr(), U.data_ptr() + rows, ptr(E_l), ptr(Q), float(weights[0]),
             float(weights[1]), float(weights[2]), ptr(X), stream())
        ctx.geom, ctx.sel, ctx.pos, ctx.S, ctx.weights = geom, sel, pos, S, weights
        ctx.shape = (T, B)
        ctx.q_shape = None if Q is None else tuple(Q.shape)
        return X

    @staticmethod
    def backward(ctx, dX):
        dX = _f32c(dX)
        T, B = ctx.shape
        geom = ctx.geom
        dU = _empty((3, T, B, 200), dX.device)
        call("mmdfn_memset_zero", dU.data_ptr() + 2 * T * B * 200 * 4, T * B * 200 * 4, stream())      # dU[2] = 0
        dE = _empty((T, B, 200), dX.device)
        dQ = _empty(ctx.q_shape, dX.device) if ctx.q_shape is not None else None
        rows = T * B * 200 * 4
        call("mmdfn_party_pack_bwd", T, B, ctx.S, geom.N, ptr(geom.dia_off, torch.int32), ptr(ctx.sel, torch.int32),
             ptr(ctx.pos, torch.int32), ptr(dX), float(ctx.weights[0]), float(ctx.weights[1]), float(ctx.weights[2]),
             dU.data_ptr(), dU.data_ptr() + rows, ptr(dE), ptr(dQ), stream())
        return dU, dE, dQ, None, None, None, None, None


class UnpackPadFn(torch.autograd.Function):
    """packed (N,D) rows of one modality + padded tail of `pad_src` (T,B,D) -> padded (T,B,D)."""

    @staticmethod
    def forward(ctx, packed, pad_src, geom):
        packed, pad_src = _f32c(packed), _f32c(pad_src)
        T, B, D = pad_src.shape
        out = _empty((T, B, D), packed.device)
        call("mmdfn_unpack_pad_fwd", T, B, D, ptr(geom.dia_off, torch.int32), ptr(packed), ptr(pad_src), ptr(out), stream())
        ctx.geom, ctx.shape, ctx.n = geom, (T, B, D), packed.shape[0]
        return out

    @staticmethod
    def backward(ctx, dM):
        T, B, D = ctx.shape
        dM = _f32c(dM)
        d_packed, d_pad = _empty((ctx.n, D), dM.device), _empty((T, B, D), dM.device)
        call("mmdfn_unpack_pad_bwd", T, B, D, ptr(ctx.geom.dia_off, torch.int32), ptr(dM), ptr(d_packed), ptr(d_pad), stream())
        return d_packed, d_pad, None


# ---------------------------------------------------------------------------------------------
# a13: Memory Fusion Network block and the glue of the 'mfn' head
# ---------------------------------------------------------------------------------------------
MFN_KEYS = tuple(f"lstm_{m}.{k}" for m in "lav" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")) + tuple(
    f"{name}.{k}" for name in ("att1_fc1", "att1_fc2", "att2_fc1", "att2_fc2", "gamma1_fc1", "gamma1_fc2", "gamma2_fc1", "gamma2_fc2")
    for k in ("weight", "bias"))


def _u8_table(masks):
    import ctypes
    return (ctypes.c_void_p * len(masks))(*[ptr(m, U8) for m in masks])


class MFNFn(torch.autograd.Function):
    """MFN.forward (code/model_fusion.py:62-120): x (T, n, 900) -> (T, n, 400); masks = None or 4 uint8 keep masks (T n, 100)."""

    @staticmethod
    def forward(ctx, x, masks, mask_scale, *w):
        x = _f32c(x)
        w = [_f32c(t) for t in w]
        T, n = x.shape[0], x.shape[1]
        out = _empty((T, n, 400), x.device)
        ws = _empty((query("mmdfn_mfn_ws_floats", T, n),), x.device)
        tab = ptr_table(w)
        mt = _u8_table(masks) if masks is not None else None
        call("mmdfn_mfn_fwd", T, n, ptr(x), tab, mt, float(mask_scale), ptr(out), ptr(ws), stream())
        ctx.save_for_backward(x, out, ws, *w)
        ctx.masks, ctx.mask_scale = masks, float(mask_scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out, ws, *w = ctx.saved_tensors
        T, n = x.shape[0], x.shape[1]
        dout = _f32c(dout)
        dx = _empty(x.shape, x.device)
        dw = [_empty(t.shape, x.device) for t in w]
        wsb = _empty((query("mmdfn_mfn_bwd_ws_floats", T, n),), x.device)
        tab, dtab = ptr_table(w), ptr_table(dw)
        mt = _u8_table(ctx.masks) if ctx.masks is not None else None
        call("mmdfn_mfn_bwd", T, n, ptr(x), tab, mt, ctx.mask_scale, ptr(out), ptr(ws), ptr(dout), ptr(dx), dtab, ptr(wsb), stream())
        return (dx, None, None, *dw)


class MFNPackFn(torch.autograd.Function):
    """stacked node features F (3N, 300) -> padded time-major window (T, B, 900), block j from modality perm[j]."""

    @staticmethod
    def forward(ctx, F_, geom, T, perm):
        F_ = _f32c(F_)
        B, N = geom.B, geom.N
        x = _empty((T, B, 900), F_.device)
        call("mmdfn_mfn_pack_fwd", T, B, N, ptr(geom.dia_off, torch.int32), *perm, ptr(F_), ptr(x), stream())
        ctx.geom, ctx.T, ctx.perm = geom, T, tuple(perm)
        return x

    @staticmethod
    def backward(ctx, dx):
        geom = ctx.geom
        dx = _f32c(dx)
        dF = _empty((3 * geom.N, 300), dx.device)
        call("mmdfn_mfn_pack_bwd", ctx.T, geom.B, geom.N, ptr(geom.dia_off, torch.int32), *ctx.perm, ptr(dx), ptr(dF), stream())
        return dF, None, None, None


class MFNUnpadFn(torch.autograd.Function):
    """(T, B, 400) -> the valid rows in node order (N, 400)   (code/model.py:1280-1285)."""

    @staticmethod
    def forward(ctx, out, geom):
        out = _f32c(out)
        T, B = out.shape[0], out.shape[1]
        feat = _empty((geom.N, 400), out.device)
        call("mmdfn_mfn_unpad_fwd", T, B, ptr(geom.dia_off, torch.int32), ptr(out), ptr(feat), stream())
        ctx.geom, ctx.shape = geom, (T, B)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        T, B = ctx.shape
        dfeat = _f32c(dfeat)
        dout = _empty((T, B, 400), dfeat.device)
        call("mmdfn_mfn_unpad_bwd", T, B, ptr(ctx.geom.dia_off, torch.int32), ptr(dfeat), ptr(dout), stream())
        return dout, None


class ReluMaskFn(torch.autograd.Function):
    """y = relu(x) * keep * scale: nn.Dropout followed by nn.ReLU (code/model.py:1290-1291)."""

    @staticmethod
    def forward(ctx, x, mask, scale):
        x = _f32c(x)
        y = _empty(x.shape, x.device)
        call("mmdfn_relu_mask_fwd", x.numel(), ptr(x), ptr(mask, U8), float(scale), ptr(y), stream())
        ctx.save_for_backward(y)
        ctx.ind = float(scale) if mask is not None else 1.0
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = _empty(y.shape, y.device)
        call("mmdfn_relu_mask_bwd", y.numel(), ptr(dy), ptr(y), ctx.ind, ptr(dx), stream())
        return dx, None, None


# ---------------------------------------------------------------------------------------------
# f4: low-rank multimodal fusion (LMF)
# ---------------------------------------------------------------------------------------------
class LMFFuseFn(torch.autograd.Function):
    """out = sum_r w_r prod_m ([1, h_m] . factor_m[r]) + bias (code/model_fusion.py:292-305); h_m (N, H) from the three
    sub-network Linears, factor_m (R, H + 1, O), w (1, R), bias (1, O)."""

    @staticmethod
    def forward(ctx, ha, hv, ht, fa, fv, ft, w, bias):
        hs = [_f32c(x) for x in (ha, hv, ht)]
        fs = [_f32c(x) for x in (fa, fv, ft)]
        w, bias = _f32c(w), _f32c(bias)
        N, H = hs[0].shape
        R, _, O = fs[0].shape
        fz = _empty((query("mmdfn_lmf_ws_floats", N, R, O),), ha.device)
        out = _empty((N, O), ha.device)
        ht_, ft_ = ptr_table(hs), ptr_table(fs)
        call("mmdfn_lmf_fuse_fwd", N, H, O, R, ht_, ft_, ptr(w), ptr(bias), ptr(fz), ptr(out), stream())
        ctx.save_for_backward(*hs, *fs, w, fz)
        ctx.dims = (N, H, O, R)
        return out

    @staticmethod
    def backward(ctx, dout):
        hs, fs = list(ctx.saved_tensors[:3]), list(ctx.saved_tensors[3:6])
        w, fz = ctx.saved_tensors[6], ctx.saved_tensors[7]
        N, H, O, R = ctx.dims
        dev = dout.device
        dout = _f32c(dout)
        dfz = _empty((max(3 * R * N * O, 1),), dev)
        dh = [_empty((N, H), dev) for _ in range(3)]
        df = [_empty((R, H + 1, O), dev) for _ in range(3)]
        dw = torch.zeros((1, R), device=dev, dtype=torch.float32)
        dbias = _empty((1, O), dev)
        call("mmdfn_lmf_fuse_bwd", N, H, O, R, ptr_table(hs), ptr_table(fs), ptr(w), ptr(fz), ptr(dout), ptr(dfz), ptr_table(dh),
             ptr_table(df), ptr(dw), ptr(dbias), stream())
        return (*dh, *df, dw, dbias)


# ---------------------------------------------------------------------------------------------
# f4: tensor fusion network (TFN)
# ---------------------------------------------------------------------------------------------
class TFNFuseFn(torch.autograd.Function):
    """y1 = dropout([1,h_a] (x) [1,h_v] (x) [1,h_t]) W1^T + b1 (code/model_fusion.py:186-205, before the ReLU); the fusion
    tensor lives one row chunk at a time inside the call.  p > 0: keep bits from the counter-based generator (same counter
    in forward and backward)."""

    @staticmethod
    def forward(ctx, ha, hv, ht, W1, b1, p):
        hs = [_f32c(x) for x in (ha, hv, ht)]
        W1, b1 = _f32c(W1), _f32c(b1)
        N = hs[0].shape[0]
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        offset = _mask_counter[0]
        if p > 0:
            if _STEP_STATE[0] is not None:
                raise MMDFNError("TFN dropout inside a captured step is not supported (its counter base is a launch argument)")
            _mask_counter[0] = (_mask_counter[0] + N * 1030301) & 0xFFFFFFFFFFFFFFFF
        ws = _empty((query("mmdfn_tfn_ws_floats", N),), ha.device)
        y1 = _empty((N, 300), ha.device)
        call("mmdfn_tfn_fuse_fwd", N, ptr(hs[0]), ptr(hs[1]), ptr(hs[2]), ptr(W1), ptr(b1), float(p), seed, offset, ptr(y1), ptr(ws),
             stream())
        ctx.save_for_backward(*hs, W1)
        ctx.cfg = (N, float(p), seed, offset)
        return y1

    @staticmethod
    def backward(ctx, dy1):
        ha, hv, ht, W1 = ctx.saved_tensors
        N, p, seed, offset = ctx.cfg
        dev = dy1.device
        dy1 = _f32c(dy1)
        dh = [_empty((N, 100), dev) for _ in range(3)]
        dW1, db1 = _empty(tuple(W1.shape), dev), _empty((300,), dev)
        d1 = _empty((max(3 * N * 101, 1),), dev)
        ws = _empty((query("mmdfn_tfn_ws_floats", N),), dev)
        call("mmdfn_tfn_fuse_bwd", N, ptr(ha), ptr(hv), ptr(ht), ptr(W1), p, seed, offset, ptr(dy1), ptr(dh[0]), ptr(dh[1]), ptr(dh[2]),
             ptr(dW1), ptr(db1), 0, ptr(d1), ptr(ws), stream())
        return dh[0], dh[1], dh[2], dW1, db1, None


# ---------------------------------------------------------------------------------------------
# k13 (f3): nodal attention of the relation path's classifier head
# ---------------------------------------------------------------------------------------------
class NodalGeom:
    """per-dialogue L x L block offsets and the dialogue index of every node row (device), built once per batch geometry"""

    def __init__(self, lengths, device):
        self.lengths = [int(x) for x in lengths]
        self.B, self.N = len(self.lengths), int(sum(self.lengths))
        self.Lmax = int(max(self.lengths)) if self.lengths else 0
        sq = [0] + list(itertools.accumulate(L * L for L in self.lengths))
        self.nsq = sq[-1]
        self.dia_off = torch.tensor([0] + list(itertools.accumulate(self.lengths)), dtype=torch.int32).to(device)
        self.sq_off = torch.tensor(sq, dtype=torch.int64).to(device)
        self.row_dia = torch.tensor([b for b, L in enumerate(self.lengths) for _ in range(L)], dtype=torch.int32).to(device)


class NodalAttnFn(torch.autograd.Function):
    """O_b = softmax_rows(tanh(Q_b E_b^T)) E_b per dialogue (code/model.py:614-645 with MatchingAttention 'general2', :66-76);
    Q = transform(E) comes from LinearFn, so autograd adds the projection's path (dE += dQ W, dW, db) itself."""

    @staticmethod
    def forward(ctx, E, Q, g):
        E, Q = _f32c(E), _f32c(Q)
        N, D = E.shape
        P, S = _empty((max(g.nsq, 1),), E.device), _empty((max(g.nsq, 1),), E.device)
        O = _empty((N, D), E.device)
        call("mmdfn_nodal_attn_fwd", g.B, N, D, g.Lmax, ptr(g.dia_off, torch.int32), ptr(g.sq_off, torch.int64),
             ptr(g.row_dia, torch.int32), ptr(E), ptr(Q), ptr(P), ptr(S), ptr(O), stream())
        ctx.save_for_backward(E, Q, P, S)
        ctx.g = g
        return O

    @staticmethod
    def backward(ctx, dO):
        E, Q, P, S = ctx.saved_tensors
        g, (N, D) = ctx.g, E.shape
        dO = _f32c(dO)
        dA = _empty((max(g.nsq, 1),), E.device)
        dQ, dE = _empty((N, D), E.device), _empty((N, D), E.device)
        call("mmdfn_nodal_attn_bwd", g.B, N, D, g.Lmax, ptr(g.dia_off, torch.int32), ptr(g.sq_off, torch.int64),
             ptr(g.row_dia, torch.int32), ptr(E), ptr(Q), ptr(P), ptr(S), ptr(dO), ptr(dA), ptr(dQ), ptr(dE), stream())
        return dE, dQ, None


# ---------------------------------------------------------------------------------------------
# k5: block-compact adjacency
# ---------------------------------------------------------------------------------------------
class AdjFn(torch.autograd.Function):
    """X (3N,200) -> (adj_blk (sum 3L^2), adj_diag (3,N)), differentiable w.r.t. X."""

    @staticmethod
    def forward(ctx, X, geom, modal_weight):
        X = _f32c(X)
        dev = X.device
        N = geom.N
        adj_blk = _empty((geom.nblk,), dev)
        adj_diag = _empty((3, N), dev)
        dinv, rinv, deg = _empty((3 * N,), dev), _empty((3 * N,), dev), _empty((3 * N,), dev)
        cos_blk, cos_diag = _empty((geom.nblk,), dev), _empty((3, N), dev)
        call("mmdfn_adj_fwd", *geom.args(), ptr(X), float(modal_weight), ptr(adj_blk), ptr(adj_diag), ptr(dinv),
             ptr(rinv), ptr(cos_blk), ptr(cos_diag), ptr(deg), stream())
        ctx.save_for_backward(X, adj_blk, adj_diag, dinv, rinv, cos_blk, cos_diag)
        ctx.geom, ctx.modal_weight = geom, float(modal_weight)
        return adj_blk, adj_diag

    @staticmethod
    def backward(ctx, d_blk, d_diag):
        X, adj_blk, adj_diag, dinv, rinv, cos_blk, cos_diag = ctx.saved_tensors
        geom = ctx.geom
        dev = X.device
        d_blk = _f32c(d_blk).clone() if d_blk is not None else torch.zeros_like(adj_blk)   # clobbered below
        d_diag = _f32c(d_diag) if d_diag is not None else torch.zeros_like(adj_diag)
        dX = _empty(X.shape, dev)
        dd = _empty((3 * geom.N,), dev)
        call("mmdfn_adj_bwd", *geom.args(), ptr(X), ctx.modal_weight, ptr(adj_blk), ptr(adj_diag), ptr(dinv), ptr(rinv),
             ptr(cos_blk), ptr(cos_diag), ptr(d_blk), ptr(d_diag), None, ptr(dX), ptr(dd), stream())
        return dX, None, None


def adj_densify(adj_blk, adj_diag, geom):
    dense = _empty((3 * geom.N, 3 * geom.N), adj_blk.device)
    call("mmdfn_adj_densify", *geom.args(), ptr(adj_blk), ptr(adj_diag), ptr(dense), stream())
    return dense


class SpmmFn(torch.autograd.Function):
    """y = A_hat x on the block-compact adjacency (torch.spmm(adj, input), code/model_GCN.py:178)."""

    @staticmethod
    def forward(ctx, adj_blk, adj_diag, x, geom):
        x = _f32c(x)
        G = x.shape[1]
        y = _empty(x.shape, x.device)
        call("mmdfn_adj_spmm", *geom.args(), ptr(adj_blk), ptr(adj_diag), ptr(x), G, ptr(y), stream())
        ctx.save_for_backward(adj_blk, adj_diag, x)
        ctx.geom = geom
        return y

    @staticmethod
    def backward(ctx, dy):
        adj_blk, adj_diag, x = ctx.saved_tensors
        geom = ctx.geom
        dy = _f32c(dy)
        G = x.shape[1]
        d_blk = d_diag = dx = None
        if ctx.needs_input_grad[2]:
            dx = _empty(x.shape, x.device)    # A_hat is symmetric: dx = A_hat dy
            call("mmdfn_adj_spmm", *geom.args(), ptr(adj_blk), ptr(adj_diag), ptr(dy), G, ptr(dx), stream())
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            d_blk, d_diag = _empty(adj_blk.shape, x.device), _empty(adj_diag.shape, x.device)
            call("mmdfn_adj_grad", *geom.args(), ptr(dy), ptr(x), G, ptr(d_blk), ptr(d_diag), 0, stream())
        return d_blk, d_diag, dx, None


# ---------------------------------------------------------------------------------------------
# k6/k7/k8: GCNII_lyc stack
# ---------------------------------------------------------------------------------------------
class GCNStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, adj_blk, adj_diag, geom, K, reason_flag, lamda, alpha, mask_x, mask_h0, mask_layers,
                mask_scale, W0, b0, w_ih, w_hh, b_ih, b_hh, *convW):
        X = _f32c(X)
        dev = X.device
        n3 = 3 * geom.N
        convW = [_f32c(w) for w in convW]
        W0, b0, w_ih, w_hh, b_ih, b_hh = (_f32c(t) for t in (W0, b0, w_ih, w_hh, b_ih, b_hh))
        F_ = _empty((n3, 300), dev)
        ws = _empty((query("mmdfn_gcn_stack_ws_floats", n3, K),), dev)
        tab = ptr_table(convW) if K > 0 else None
        call("mmdfn_gcn_stack_fwd", *geom.args(), ptr(adj_blk), ptr(adj_diag), ptr(X), K, int(reason_flag),
             float(lamda), float(alpha), ptr(W0), ptr(b0), tab, ptr(w_ih), ptr(w_hh), ptr(b_ih), ptr(b_hh),
             ptr(mask_x, U8), ptr(mask_h0, U8), ptr(mask_layers, U8), float(mask_scale), ptr(F_), ptr(ws), stream())
        ctx.save_for_backward(adj_blk, adj_diag, F_, ws, W0, w_ih, w_hh, *convW)
        ctx.sink_key = _SINK_KEY[0]
        ctx.cfg = (geom, K, int(reason_flag), float(lamda), float(alpha), mask_x, mask_h0, mask_layers, float(mask_scale))
        return F_

    @staticmethod
    def backward(ctx, dF):
        adj_blk, adj_diag, F_, ws, W0, w_ih, w_hh = ctx.saved_tensors[:7]
        convW = list(ctx.saved_tensors[7:])
        geom, K, reason_flag, lamda, alpha, mask_x, mask_h0, mask_layers, mask_scale = ctx.cfg
        dev = F_.device
        n3 = 3 * geom.N
        dF = _f32c(dF)
        dX = _empty((n3, 200), dev)
        want_adj = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        d_blk = _empty(adj_blk.shape, dev) if want_adj else None
        d_diag = _empty(adj_diag.shape, dev) if want_adj else None
        if want_adj and K == 0:
            call("mmdfn_memset_zero", d_blk.data_ptr(), d_blk.numel() * 4, stream())
            call("mmdfn_memset_zero", d_diag.data_ptr(), d_diag.numel() * 4, stream())
        # every parameter gradient of the stack lives in ONE zero-filled buffer (a single memset, or the trainer's
        # bucket segment); the LSTM's four tensors are part of it only when the gate is in use
        use_rnn = bool(reason_flag) and K > 0
        sizes = [20000, 100] + ([40000, 40000, 400, 400] if use_rnn else []) + [20000] * K
        flat, direct = _grad_buffer(ctx.sink_key, sum(sizes), dev)
        views, off = [], 0
        for n in sizes:
            views.append(flat[off:off + n])
            off += n
        dW0, db0 = views[0].view(100, 200), views[1]
        if use_rnn:
            dw_ih, dw_hh, db_ih, db_hh = views[2].view(400, 100), views[3].view(400, 100), views[4], views[5]
        else:
            dw_ih = dw_hh = db_ih = db_hh = None
        dconv = [v.view(200, 100) for v in views[(6 if use_rnn else 2):]]
        wsb = _empty((query("mmdfn_gcn_stack_bwd_ws_floats", n3, K),), dev)
        tab = ptr_table(convW) if K > 0 else None
        dtab = ptr_table(dconv) if K > 0 else None
        call("mmdfn_gcn_stack_bwd", *geom.args(), ptr(adj_blk), ptr(adj_diag), K, reason_flag, lamda, alpha, ptr(W0),
             tab, ptr(w_ih), ptr(w_hh), ptr(mask_x, U8), ptr(mask_h0, U8), ptr(mask_layers, U8), mask_scale,
             ptr(F_), ptr(ws), ptr(dF), ptr(dX), ptr(d_blk), ptr(d_diag), ptr(dW0), ptr(db0), dtab, ptr(dw_ih),
             ptr(dw_hh), ptr(db_ih), ptr(db_hh), 1, ptr(wsb), stream())
        # (reason_flag off or K == 0: the reference never touches the LSTM -- its grads stay None and Adam skips the weights)
        if direct:
            _SINK[0].ready(ctx.sink_key)
            return (dX, d_blk, d_diag, None, None, None, None, None, None, None, None, None,
                    None, None, None, None, None, None, *([None] * K))
        return (dX, d_blk, d_diag, None, None, None, None, None, None, None, None, None,
                dW0, db0, dw_ih, dw_hh, db_ih, db_hh, *dconv)


# ---------------------------------------------------------------------------------------------
# k9: head and focal loss
# ---------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    """log_softmax(relu(dropout([F_a | F_v | F_l])) Wc^T + bc)  (code/model.py:1328-1337)."""

    @staticmethod
    def forward(ctx, F_, N, mask, mask_scale, Wc, bc, relu=True):
        F_, Wc, bc = _f32c(F_), _f32c(Wc), _f32c(bc)
        C = Wc.shape[0]
        R = _empty(F_.shape, F_.device)
        lp = _empty((N, C), F_.device)
        call("mmdfn_head_fwd", N, C, ptr(F_), ptr(mask, U8), float(mask_scale), int(relu), ptr(Wc), ptr(bc), ptr(R), ptr(lp),
             stream())
        ctx.save_for_backward(R, lp, Wc)
        ctx.sink_key = _SINK_KEY[0]
        ctx.mask, ctx.mask_scale, ctx.N, ctx.relu = mask, float(mask_scale), N, int(relu)
        return lp

    @staticmethod
    def backward(ctx, dlp):
        R, lp, Wc = ctx.saved_tensors
        N, C = ctx.N, Wc.shape[0]
        dev = R.device
        dlp = _f32c(dlp)
        dF = _empty(R.shape, dev)
        flat, direct = _grad_buffer(ctx.sink_key, C * 900 + C, dev)
        dWc, dbc = flat[:C * 900].view(C, 900), flat[C * 900:]
        scratch = _empty((max(N, 1), C), dev)
        call("mmdfn_head_bwd", N, C, ptr(ctx.mask, U8), ctx.mask_scale, ctx.relu, ptr(Wc), ptr(R), ptr(lp), ptr(dlp), ptr(dF),
             ptr(dWc), ptr(dbc), 1, ptr(scratch), stream())
        if direct:
            _SINK[0].ready(ctx.sink_key)
            return dF, None, None, None, None, None, None
        return dF, None, None, None, dWc, dbc, None


class FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_prob, target, alpha, gamma, size_average):
        log_prob = _f32c(log_prob)
        target = target.contiguous().view(-1)
        if target.dtype != torch.int64:
            target = target.long()
        N, C = log_prob.shape
        loss = _empty((1,), log_prob.device)
        call("mmdfn_focal_loss_fwd", N, C, ptr(log_prob), ptr(target, torch.int64), ptr(alpha), float(gamma),
             int(size_average), ptr(loss), stream())
        ctx.save_for_backward(log_prob, target, alpha)
        ctx.gamma, ctx.size_average = float(gamma), int(size_average)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        log_prob, target, alpha = ctx.saved_tensors
        N, C = log_prob.shape
        dlp = _empty(log_prob.shape, log_prob.device)
        dloss = _f32c(dloss).reshape(1)
        call("mmdfn_focal_loss_bwd", N, C, ptr(log_prob), ptr(target, torch.int64), ptr(alpha), ctx.gamma,
             ctx.size_average, ptr(dloss), ptr(dlp), stream())
        return dlp, None, None, None, None
