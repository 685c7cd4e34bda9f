"""Network plans: the reference's modules compiled to a DAG of kernel launches.

A *plan* is built by the modules of Networks.py emitting nodes into a PlanBuilder (conv layers with
their input transform, the VAE bottleneck, the discriminator head).  `run_forward` executes it on the
current CUDA stream through the C ABI only (ops.py); `run_backward` executes the exact adjoint in
reverse order: for every activation the gradient is *gathered* from the padded-input gradients of its
consumers (halo folded, pixel shuffle inverted), pushed through the activation / InstanceNorm
derivative into a zero-haloed dY, then the weight-gradient and data-gradient GEMMs run.

Activations ("Act") are what a reference module returns (Networks.py:57-149): the raw conv output
plus a *pending* normalisation / activation / residual that is applied by the consumer's input
transform, so no normalised tensor is ever written to HBM on its own."""
from __future__ import annotations

import torch

from . import lib as L
from . import ops

_STATE = {"dtype": torch.bfloat16, "input_grads": True, "wgrad_side": True}


def set_wgrad_side(flag):
    """True (default): the weight-gradient GEMMs of a backward pass run on a low-priority side stream of their chain
    (they are off the critical path dY -> data gradient -> next layer's gather) and are joined when the pass ends.
    Measured (CUDA-graph replay, one B200): batch 8 13.21 -> 12.76 ms, batch 16 22.49 -> 22.43, batch 64 80.93 -> 80.72."""
    _STATE["wgrad_side"] = bool(flag)


def set_precision(mode):
    """'bf16' (tcgen05 tensor-core kernels, fp32 accumulate) or 'fp32' (FFMA parity kernels)."""
    _STATE["dtype"] = {"bf16": torch.bfloat16, "fp32": torch.float32}[mode]


def get_precision():
    return "bf16" if _STATE["dtype"] == torch.bfloat16 else "fp32"


class no_wgrad:
    """Context manager: skip weight gradients of the given modules (the reference computes the
    discriminators' weight gradients during the generator backward and then discards them,
    Networks.py:2020-2025)."""

    def __init__(self, *modules):
        self.params = [p for m in modules for p in m.parameters()]

    def __enter__(self):
        self.prev = [getattr(p, "_vcg_skip_wgrad", False) for p in self.params]
        for p in self.params:
            p._vcg_skip_wgrad = True

    def __exit__(self, *exc):
        for p, v in zip(self.params, self.prev):
            p._vcg_skip_wgrad = v


def _wants_grad(p):
    return p.requires_grad and not getattr(p, "_vcg_skip_wgrad", False)


# ----------------------------------------------------------------------------- plan description
class Act:
    """A finished activation: raw tensor (conv output / external / z) + pending norm, act, residual."""
    _next = 0

    def __init__(self, c, c_log, h, w, producer=None, norm=False, act=L.ACT_NONE, res=None, kind="conv"):
        self.id = Act._next
        Act._next += 1
        self.c, self.c_log, self.h, self.w = c, c_log, h, w
        self.producer, self.norm, self.act, self.res, self.kind = producer, norm, act, res, kind
        self.consumers = []      # nodes reading this act
        self.passthrough = []    # acts that add this act as their residual
        if res is not None:
            res.passthrough.append(self)


class ConvNode:
    def __init__(self, holder, spec, inp, mode, pad, pre_act):
        self.holder, self.spec, self.inp, self.mode, self.pad, self.pre_act = holder, spec, inp, mode, pad, pre_act
        self.out_acts = []


class ReparamNode:
    def __init__(self, mu, lv, z):
        self.mu, self.lv, self.z = mu, lv, z


class HeadNode:
    def __init__(self, holder, inp):
        self.holder, self.inp = holder, inp


class PlanBuilder:
    def __init__(self, n):
        self.n = n
        self.nodes = []
        self.inputs = []
        self.outputs = []     # (act, kind) kind in {"image", "mu", "logvar", "score"}

    def input(self, c_log, h, w):
        a = Act(ops.rup(c_log, 8), c_log, h, w, kind="ext")
        self.inputs.append(a)
        return a

    def conv(self, holder, inp, mode=L.MODE_PLAIN, pad=1, pre_act=L.ACT_NONE, norm=False, act=L.ACT_NONE):
        co, ci, kh, kw = holder.weight.shape
        if mode == L.MODE_PLAIN:
            wmap, c_phys, hs, ws = L.WMAP_PLAIN, inp.c, inp.h, inp.w
        elif mode == L.MODE_SHUFFLE:
            wmap, c_phys, hs, ws = L.WMAP_PLAIN, inp.c // 4, inp.h * 2, inp.w * 2
        elif mode == L.MODE_UNSHUFFLE:
            wmap, c_phys, hs, ws = L.WMAP_UNSHUFFLE, inp.c * 4, inp.h // 2, inp.w // 2
        else:
            wmap, c_phys, hs, ws = L.WMAP_S2D, inp.c * 4, None, None
        spec = ops.ConvSpec(co, ci, kh, kw, wmap, c_phys)
        if mode == L.MODE_PAD_S2D:
            ho, wo = (inp.h + 2 * pad) // 2 - spec.pkh + 1, (inp.w + 2 * pad) // 2 - spec.pkw + 1
        else:
            ho, wo = hs + 2 * pad - kh + 1, ws + 2 * pad - kw + 1
        node = ConvNode(holder, spec, inp, mode, pad, pre_act)
        out = Act(spec.out_c, co, ho, wo, producer=node, norm=norm, act=act)
        node.out_acts.append(out)
        inp.consumers.append(node)
        self.nodes.append(node)
        return out

    def residual(self, main, res):
        """main + res, where main is a conv output with its pending norm (R block, Networks.py:108-116)."""
        a = Act(main.c, main.c_log, main.h, main.w, producer=main.producer, norm=main.norm, act=main.act, res=res)
        main.producer.out_acts.append(a)
        return a

    def reparam(self, mu, lv):
        z = Act(mu.c, mu.c_log, mu.h, mu.w, kind="z")
        node = ReparamNode(mu, lv, z)
        z.producer = node
        mu.consumers.append(node)
        lv.consumers.append(node)
        self.nodes.append(node)
        return z

    def head(self, holder, inp):
        node = HeadNode(holder, inp)
        inp.consumers.append(node)
        self.nodes.append(node)
        self.outputs.append((node, "score"))
        return node

    def output(self, act, kind="image"):
        self.outputs.append((act, kind))


# ----------------------------------------------------------------------------- parameter caches
_TABLES = {}      # cached device-side job tables of the multi-tensor pack / unpack launches (keyed by pointers)
_PENDING = {}     # id -> PackedConv whose fp32 accumulators hold weight gradient not yet added to .grad
_FLUSH_QUEUED = [False]
_TRACKERS = []    # optim._Tracker objects of the backward pass that is running (FusedAdam.track)


def _cached_table(key, build):
    """Device tables are never evicted: their addresses may be baked into a captured CUDA graph (graph.py).  A table
    is a few KB and a training run creates a few dozen (one per bucket and layout)."""
    t = _TABLES.get(key)
    if t is None:
        t = _TABLES[key] = build()
    return t


def has_pending():
    return bool(_PENDING)


def _mark_grad_written(param):
    opt = getattr(param, "_vcg_opt", None)
    if opt is not None:
        opt.mark_dirty()


class PackedConv:
    """Kernel-layout copies of one conv's master weight, refreshed when the master changes
    (tensor._version is bumped by load_state_dict, _vcg_epoch by the fused optimiser), and the fp32
    accumulators the weight-gradient GEMMs add into.  Accumulators are zero-initialised once; flush_grads()
    adds them to .grad and re-zeroes them in the same launch, so no per-step clearing is needed."""

    def __init__(self, holder, spec):
        self.holder, self.spec = holder, spec
        self.key = None
        self.wk = self.wkT = self.dw = self.dbias = None
        self.pending = self.pending_bias = False

    def _key(self, dtype):
        w = self.holder.weight
        return (w._version, getattr(w, "_vcg_epoch", 0), w.data_ptr(), dtype, w.device)

    def stale(self, dtype):
        return self._key(dtype) != self.key

    def buffers(self, dtype):
        w = self.holder.weight
        if self.wk is None or self.wk.dtype != dtype or self.wk.device != w.device:
            # zero-filled: the padding rows / columns are never written by the multi-tensor pack
            self.wk = torch.zeros(self.spec.packed_shape(False), dtype=dtype, device=w.device)
            self.wkT = torch.zeros(self.spec.packed_shape(True), dtype=dtype, device=w.device)
        return self.wk, self.wkT

    def refresh(self, dtype):
        if self.stale(dtype):
            refresh_many([self], dtype)
        return self

    def grad_buffers(self):
        dev = self.holder.weight.device
        if self.dw is None or self.dw.device != dev:
            self.dw = torch.zeros(self.spec.packed_shape(False), dtype=torch.float32, device=dev)
            self.dbias = torch.zeros(ops.rup(self.spec.co, 8), dtype=torch.float32, device=dev)
        return self.dw, self.dbias

    def mark_pending(self, bias):
        if not self.pending:
            self.pending = True
            _PENDING[id(self)] = self
        self.pending_bias = self.pending_bias or bias


def refresh_many(pcs, dtype):
    """Re-pack every stale filter of `pcs` (both layouts) with ONE multi-tensor launch."""
    stale, seen = [], set()
    for pc in pcs:
        if id(pc) not in seen and pc.stale(dtype):
            seen.add(id(pc))
            stale.append(pc)
    if not stale:
        return
    entries, cacheable = [], True
    for pc in stale:
        w = pc.holder.weight.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w, cacheable = w.float().contiguous(), False
        wk, wkT = pc.buffers(dtype)
        entries.append((pc.spec, w, wk, wkT))
    dev = entries[0][1].device
    # (pointers AND geometry: a freed model's addresses can be handed to differently shaped tensors later)
    key = ("pack", dtype, tuple((e[0], e[1].data_ptr(), e[2].data_ptr(), e[3].data_ptr()) for e in entries))
    table = _cached_table(key, lambda: ops.wjob_table(entries, dev)) if cacheable else ops.wjob_table(entries, dev)
    ops.wpack_multi(table, dtype)
    for pc in stale:
        pc.key = pc._key(dtype)


def refresh_holders(holders, dtype=None):
    """re-pack the stale filters of the given parameter holders (one launch)"""
    pcs = [pc for h in holders for pc in h.__dict__.get("_vcg_packed", {}).values()]
    if pcs:
        refresh_many(pcs, dtype or _STATE["dtype"])


def flush_grads(holders=None):
    """Add pending weight / bias gradient accumulators to their parameters' .grad (and re-zero the accumulators):
    two multi-tensor launches instead of one unpack per convolution.  holders=None: everything that is pending,
    except what a running FusedAdam.track() will hand over bucket by bucket; else: only these holders."""
    if not _PENDING:
        return
    if holders is None:
        pcs = [pc for pc in _PENDING.values() if not any(t.tracks(pc.holder) for t in _TRACKERS)]
    else:
        pcs = [pc for h in holders for pc in h.__dict__.get("_vcg_packed", {}).values() if pc.pending]
    if not pcs:
        return
    went, bent = [], []
    for pc in pcs:
        _PENDING.pop(id(pc), None)
        w = pc.holder.weight
        if w.grad is None:
            w.grad = torch.zeros_like(w, memory_format=torch.contiguous_format)
        _mark_grad_written(w)
        went.append((pc.spec, w.grad, pc.dw, None))
        if pc.pending_bias:
            b = pc.holder.bias
            if b.grad is None:
                b.grad = torch.zeros_like(b)
            bent.append((pc.dbias, b.grad, pc.spec.co))
        pc.pending = pc.pending_bias = False
    dev = went[0][1].device
    key = ("unpack", tuple((e[0], e[1].data_ptr(), e[2].data_ptr()) for e in went))
    ops.wunpack_multi(_cached_table(key, lambda: ops.wjob_table(went, dev)))
    if bent:
        key = ("vec", tuple((e[0].data_ptr(), e[1].data_ptr(), e[2]) for e in bent))
        ops.vecflush_multi(_cached_table(key, lambda: ops.vecjob_table(bent, dev)))


def queue_flush():
    """Called from inside an autograd backward: flush once, when the whole backward pass has finished."""
    if _FLUSH_QUEUED[0]:
        return
    _FLUSH_QUEUED[0] = True

    def cb():
        _FLUSH_QUEUED[0] = False
        from . import lanes
        lanes.join_dirty()          # backward nodes ran on the lanes of their forward passes (lanes.py)
        flush_grads()
    torch.autograd.Variable._execution_engine.queue_callback(cb)


def packed_for(holder, spec):
    cache = holder.__dict__.setdefault("_vcg_packed", {})
    pc = cache.get(spec)
    if pc is None:
        pc = cache[spec] = PackedConv(holder, spec)
    return pc


def _accumulate_grad(param, fn_write):
    """fn_write(grad_tensor, accumulate: bool) fills/accumulates param.grad in place (no autograd
    AccumulateGrad round trip: the wgrad GEMM result is unpacked straight into .grad)."""
    if param.grad is None:
        param.grad = torch.empty_like(param, memory_format=torch.contiguous_format)
        fn_write(param.grad, False)
    else:
        fn_write(param.grad, True)


# ----------------------------------------------------------------------------- execution
class Run:
    """Tensors of one forward execution that the backward needs (saved activations)."""

    def __init__(self):
        self.Y, self.MR, self.XP = {}, {}, {}
        self.stats_hw = {}     # conv node -> pixel count when MR holds raw sums (bf16 mode), absent = finalised
        self.xp_of_node = {}
        self.plain = {}        # act id -> (padded tensor, pad): where a residual can be read from
        self.ext = {}          # act id -> dense NHWC tensor (external inputs, z)
        self.reparam = {}      # node -> (eps,)
        self.head = {}         # node -> (x dense, w_khwc, wnorm2)


class Plan:
    def __init__(self, builder):
        self.n = builder.n
        self.nodes, self.inputs, self.outputs = builder.nodes, builder.inputs, builder.outputs

    # ------------------------------------------------------------------ forward
    def _raw(self, run, act):
        return run.Y[act.producer] if act.kind == "conv" else run.ext[act.id]

    def _materialize(self, run, act, mode, pad, dtype):
        key = (act.id, mode, pad)
        xp = run.XP.get(key)
        if xp is None:
            src = self._raw(run, act)
            shape = ops.xform_dst_shape(self.n, act.h, act.w, act.c, mode, pad)
            xp = torch.empty(shape, dtype=dtype, device=src.device)
            mr = run.MR[act.producer] if act.norm else None
            resbuf, off = run.plain[act.res.id] if act.res is not None else (None, 0)
            ops.xform_fwd(src, act.c, xp, mode, pad, mr, act.act, resbuf, off,
                          stats_hw=run.stats_hw.get(act.producer, 0) if act.norm else 0)
            run.XP[key] = xp
            if mode == L.MODE_PLAIN and act.id not in run.plain:
                run.plain[act.id] = (xp, pad)
        return xp

    def run_forward(self, inputs, eps_list=(), keep=True):
        """inputs: NCHW fp32 CUDA tensors (one per plan input).  Returns (outputs, run)."""
        dtype = _STATE["dtype"]
        run = Run()
        n = self.n
        dev = inputs[0].device
        for a, x in zip(self.inputs, inputs):
            x = x.detach()
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.float().contiguous()
            t = torch.empty(n, a.h, a.w, a.c, dtype=dtype, device=dev)
            ops.pack_nchw(x, t)
            run.ext[a.id] = t
        eps_iter = iter(eps_list)
        outs = []
        node_out = {}
        out_nodes = {id(o[0]): o[1] for o in self.outputs}
        convs = [nd for nd in self.nodes if isinstance(nd, ConvNode)]
        refresh_many([packed_for(nd.holder, nd.spec) for nd in convs], dtype)      # one launch when weights changed
        # InstanceNorm sum / sum-of-squares accumulators of all layers: one buffer, one clear
        stat_off, total = {}, 0
        if dtype == torch.bfloat16:
            for nd in convs:
                if nd.out_acts[0].norm:
                    stat_off[nd] = total
                    total += n * nd.spec.co * 2
        stat_arena = ops.zero_(torch.empty(total, dtype=torch.float32, device=dev)) if total else None
        for node in self.nodes:
            if isinstance(node, ConvNode):
                xp = self._materialize(run, node.inp, node.mode, node.pad, dtype)
                run.xp_of_node[node] = xp
                pc = packed_for(node.holder, node.spec).refresh(dtype)
                a0 = node.out_acts[0]
                # the final image layer (conv -> Identity, no norm, Networks.py:192) is stored in fp32
                # and so are the bottleneck's mu / logvar convs (Networks.py:217-218): logvar reaches +-10 and
                # feeds exp(), where a bf16 ulp (0.0625 at |x|>=8) would cost 3% on sigma
                plain = (dtype != torch.float32 and len(node.out_acts) == 1 and not a0.norm
                         and a0.act == L.ACT_NONE and node.pre_act == L.ACT_NONE)
                final_image = plain and id(a0) in out_nodes and not a0.consumers
                bottleneck = plain and bool(a0.consumers) and all(isinstance(c, ReparamNode) for c in a0.consumers)
                y = torch.empty(n, a0.h, a0.w, node.spec.out_c,
                                dtype=torch.float32 if (final_image or bottleneck) else dtype, device=dev)
                bias = node.holder.bias.detach()
                if a0.norm and dtype == torch.bfloat16:
                    acc = stat_arena[stat_off[node]:stat_off[node] + n * node.spec.co * 2]
                    ops.conv_fwd(node.spec, xp, pc.wk, bias, y, acc, node.pre_act)
                    # raw {sum, sum^2}: the consumers derive mean / rstd on load (stats_hw), no finalize launch
                    run.MR[node] = acc
                    run.stats_hw[node] = a0.h * a0.w
                else:
                    ops.conv_fwd(node.spec, xp, pc.wk, bias, y, None, node.pre_act)
                    if a0.norm:
                        mr = torch.empty(n * node.spec.co * 6, dtype=torch.float32, device=dev)
                        run.MR[node] = ops.in_stats(y, node.spec.co, mr)
                run.Y[node] = y
            elif isinstance(node, ReparamNode):
                eps = next(eps_iter)
                eps = eps.detach().float().contiguous()
                mu_t, lv_t = self._raw(run, node.mu), self._raw(run, node.lv)
                c = node.mu.c
                z = torch.empty(n, node.mu.h, node.mu.w, c, dtype=dtype, device=dev)
                mu_o = torch.empty(n, c, node.mu.h, node.mu.w, dtype=torch.float32, device=dev)
                lv_o = torch.empty_like(mu_o)
                ops.reparam_fwd(mu_t, 0, lv_t, 0, eps, c, z, mu_o, lv_o, None)
                run.ext[node.z.id] = z
                run.reparam[node] = (eps,)
                node_out[node] = (mu_o, lv_o)
            else:   # HeadNode
                x = self._materialize(run, node.inp, L.MODE_PLAIN, 0, dtype)
                wk = head_weight(node.holder)
                score = torch.empty(n, dtype=torch.float32, device=dev)
                wn2 = torch.empty(1, dtype=torch.float32, device=dev)
                ops.dhead_fwd(x, wk, node.holder.bias.detach(), score, wn2)
                # NOTE: tensors returned to autograd must not be stored in `run` (ctx -> run -> output ->
                # grad_fn -> ctx would be a reference cycle that only the cyclic GC frees: GBs per step)
                run.head[node] = (x, wk, wn2)
                node_out[node] = (score,)
        for obj, kind in self.outputs:
            if kind == "score":
                outs.append(node_out[obj][0])
            elif kind == "mu":
                outs.append(node_out[obj][0])
            elif kind == "logvar":
                outs.append(node_out[obj][1])
            else:
                outs.append(self._export(run, obj, dtype))
        if not keep:
            run = None
        return outs, run

    def _export(self, run, act, dtype):
        """finished activation -> NCHW fp32 (API boundary)."""
        if act.kind == "conv" and not act.norm and act.act == L.ACT_NONE and act.res is None:
            src = run.Y[act.producer]
        elif act.kind != "conv":
            src = run.ext[act.id]
        else:
            src = self._materialize(run, act, L.MODE_PLAIN, 0, dtype)
        out = torch.empty(self.n, act.c_log, act.h, act.w, dtype=torch.float32, device=src.device)
        return ops.unpack_nchw(src, act.c_log, out)

    # ------------------------------------------------------------------ backward
    def run_backward(self, run, grads, need_input_grad=(False,), defer_flush=False):
        """grads: one NCHW fp32 tensor (or None) per plan output, in output order.
        Accumulates parameter gradients into .grad (through the fp32 accumulators of the weight-gradient GEMMs,
        flushed here unless defer_flush: the autograd wrapper flushes once at the end of the whole backward
        pass); returns input gradients (NCHW fp32 or None)."""
        dtype = _STATE["dtype"]
        n = self.n
        gs_off, gs_total = {}, 0
        for nd in self.nodes:
            if isinstance(nd, ConvNode) and nd.out_acts[0].norm:
                gs_off[nd] = gs_total
                gs_total += n * nd.spec.co * 2
        gs_arena = None
        wside, wkeep, cur_stream = None, [], None       # weight gradients on a side stream (set_wgrad_side)
        dense = {}        # act id -> list of dense NHWC grad tensors
        dxp = {}          # conv node -> padded-input gradient
        ext_mu, ext_lv, gscores = {}, {}, {}
        dev = None
        for (obj, kind), g in zip(self.outputs, grads):
            if g is None:
                continue
            dev = g.device
            g = g.detach()
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            if kind == "score":
                gscores[obj] = g
            elif kind == "mu":
                ext_mu[obj] = g
            elif kind == "logvar":
                ext_lv[obj] = g
            else:
                t = torch.empty(n, obj.h, obj.w, obj.c, dtype=dtype, device=dev)
                ops.pack_nchw(g, t)
                dense.setdefault(obj.id, []).append(t)
        if dev is None:
            return [None] * len(self.inputs)

        def sources(act, seen=None):
            s = [(t, L.MODE_PLAIN, 0) for t in dense.get(act.id, ())]
            for c in act.consumers:
                if isinstance(c, ConvNode) and c in dxp:
                    s.append((dxp[c], c.mode, c.pad, True))      # halo already folded (see conv_dgrad below)
            for p in act.passthrough:
                s.extend(sources(p))
            return s

        def needs_dx(act):
            if act.kind == "ext":
                return need_input_grad[self.inputs.index(act)]
            return True

        for node in reversed(self.nodes):
            if isinstance(node, HeadNode):
                gs = gscores.get(node)
                if gs is None:
                    continue
                x, wk, wn2 = run.head[node]
                k = wk.numel()
                dx = torch.empty_like(x)
                scratch = torch.empty(k + 8, dtype=torch.float32, device=dev)
                h = node.holder
                want_w = _wants_grad(h.weight_orig)
                gw = gb = None
                if want_w:
                    # the head's gradient is written straight into .grad in OIHW order (no temporaries, no torch ops)
                    if h.weight_orig.grad is None:
                        h.weight_orig.grad = torch.zeros_like(h.weight_orig, memory_format=torch.contiguous_format)
                    if h.bias.grad is None:
                        h.bias.grad = torch.zeros_like(h.bias)
                    gw, gb = h.weight_orig.grad, h.bias.grad
                ops.dhead_bwd(x, wk, wn2, gs, dx, gw, gb, scratch, dw_c=h.weight_orig.shape[1] if want_w else 0)
                if want_w:
                    _mark_grad_written(h.weight_orig)
                    for t in _TRACKERS:
                        t.contributed(h)
                dense.setdefault(node.inp.id, []).append(dx)
            elif isinstance(node, ReparamNode):
                srcs = sources(node.z)
                gm, gl = ext_mu.get(node), ext_lv.get(node)
                if not srcs and gm is None and gl is None:
                    continue
                c, h, w = node.mu.c, node.mu.h, node.mu.w
                dz = torch.empty(n, h, w, c, dtype=dtype, device=dev)
                if srcs:
                    ops.xform_bwd_gather(srcs, None, n, h, w, c, dz, 0)
                else:
                    ops.zero_(dz)
                dmu, dlv = torch.empty_like(dz), torch.empty_like(dz)
                eps = run.reparam[node][0]
                ops.reparam_bwd(self._raw(run, node.mu), 0, self._raw(run, node.lv), 0, eps, dz, c, dmu, dlv, gm, gl, 0.0)
                dense.setdefault(node.mu.id, []).append(dmu)
                dense.setdefault(node.lv.id, []).append(dlv)
            else:
                srcs = []
                for a in node.out_acts:
                    srcs.extend(sources(a))
                if not srcs:
                    continue          # dead branch: nothing downstream needs this layer's gradient
                a0, spec, holder = node.out_acts[0], node.spec, node.holder
                halo = spec.pkh - 1
                # the halo ring of dY is cleared by the gather kernel itself (clear_halo)
                dy = torch.empty(n, a0.h + 2 * halo, a0.w + 2 * halo, spec.out_c, dtype=dtype, device=dev)
                y = run.Y[node]
                if y.dtype != dtype:          # fp32-stored final image in bf16 mode: no norm/act there
                    y = None
                pc = packed_for(holder, spec)
                want_w = _wants_grad(holder.weight)
                want_b = want_w and _wants_grad(holder.bias)
                dw, dbias = pc.grad_buffers() if want_w else (None, None)
                if not want_b:
                    dbias = None
                if a0.norm:
                    mr = run.MR[node]
                    if gs_arena is None:      # per-(n,c) reduction buffers of all normalised layers: one clear
                        gs_arena = ops.zero_(torch.empty(gs_total, dtype=torch.float32, device=dev))
                    gs = gs_arena[gs_off[node]:gs_off[node] + n * spec.co * 2]
                    shw = run.stats_hw.get(node, 0)
                    ops.xform_bwd_gather(srcs, y, n, a0.h, a0.w, spec.out_c, dy, halo, mr, a0.act, node.pre_act, gs, None,
                                         stats_hw=shw, clear_halo=True)
                    ops.xform_bwd_norm(y, n, a0.h, a0.w, spec.out_c, dy, halo, mr, gs, node.pre_act, dbias, stats_hw=shw)
                else:
                    ops.xform_bwd_gather(srcs, y, n, a0.h, a0.w, spec.out_c, dy, halo, None, a0.act, node.pre_act, None, dbias,
                                         clear_halo=True)
                xp = run.xp_of_node[node]
                if want_w:
                    # accumulates; added to .grad by flush_grads().  Maps below 8x8 (inputs smaller than the 256x256 the
                    # networks are built for) have no tensor-core weight-gradient kernel: SIMT kernel, on request
                    if _STATE["wgrad_side"]:
                        from . import lanes
                        if wside is None:
                            cur_stream = torch.cuda.current_stream(dev)
                            wside = lanes.wgrad_stream(cur_stream)
                        ready = torch.cuda.Event()
                        ready.record(cur_stream)                  # dY is complete here
                        wside.wait_event(ready)
                        with torch.cuda.stream(wside):
                            ops.conv_wgrad(spec, xp, dy, dw, allow_simt=a0.h * a0.w < 64)
                        wkeep.append(dy)                          # alive until the side stream is joined below
                    else:
                        ops.conv_wgrad(spec, xp, dy, dw, allow_simt=a0.h * a0.w < 64)
                    pc.mark_pending(want_b)
                if needs_dx(node.inp):
                    pc.refresh(dtype)
                    g = torch.empty_like(xp)
                    ops.conv_dgrad(spec, dy, pc.wkT, g)
                    # adjoint of the reflect padding: fold the halo into the interior once, here, so that every
                    # gather of this gradient (it can feed two activations through a residual) reads one position
                    # (skipping this launch for small planes and letting the gather read the mirrors itself measured
                    #  SLOWER: batch 8 13.19 -> 13.69 ms, batch 64 81.0 -> 84.1 ms; the single-position gather is the fast one)
                    ops.fold_halo_(g, node.mode, node.pad, node.inp.h, node.inp.w, node.inp.c)
                    dxp[node] = g
                if want_w:
                    # FusedAdam.track: when this was the bucket's last contribution, its tail (unpack, all-reduce, Adam,
                    # RE-PACK of the filters) starts on the side stream behind the event recorded here -- i.e. behind
                    # this layer's weight-gradient AND data-gradient GEMM, the last readers of the packed filter
                    for t in _TRACKERS:
                        t.contributed(holder, wside)
        if wside is not None:
            cur_stream.wait_stream(wside)       # the pass's activations and dY buffers are released after this point
            wkeep.clear()
        if not defer_flush:
            flush_grads()
        res = []
        for a, need in zip(self.inputs, need_input_grad):
            srcs = sources(a) if need else []
            if not srcs:
                res.append(None)
                continue
            g = torch.empty(n, a.h, a.w, a.c, dtype=dtype, device=dev)
            ops.xform_bwd_gather(srcs, None, n, a.h, a.w, a.c, g, 0)
            out = torch.empty(n, a.c_log, a.h, a.w, dtype=torch.float32, device=dev)
            res.append(ops.unpack_nchw(g, a.c_log, out))
        return res


def head_state(holder, iterate=False):
    """Derived state of the spectral-normalised head (Networks.py:248): the (h, w, c)-ordered fp32 filter vector the
    head kernels read and aux = {sigma, |W|}, in PERSISTENT buffers (stable addresses: CUDA-graph safe), refreshed by
    one vcg_dhead_prepare launch when weight_orig changed; iterate=True also runs the reference's power iteration
    (in place on weight_u / weight_v) unless it already ran on these very W, u, v -- with a 1 x K matrix the
    iteration is idempotent (u = +-1 exactly), so the reference's repeated iterations within a step are skipped."""
    w = holder.weight_orig
    base = (w._version, getattr(w, "_vcg_epoch", 0), w.data_ptr(), w.device)
    cache = holder.__dict__.setdefault("_vcg_head", {})
    need_copy = cache.get("key") != base
    it_key = (base, holder.weight_u._version, holder.weight_v._version, holder.weight_u.data_ptr())
    need_iter = iterate and cache.get("iter_key") != it_key
    if need_copy or need_iter:
        if cache.get("w") is None or cache["w"].device != w.device:
            cache["w"] = torch.empty(w[0].numel(), dtype=torch.float32, device=w.device)
            cache["aux"] = torch.empty(2, dtype=torch.float32, device=w.device)
            cache["scratch"] = torch.empty(3, dtype=torch.float64, device=w.device)
        wd = w.detach()
        if wd.dtype != torch.float32 or not wd.is_contiguous():
            raise RuntimeError("spectral head: weight_orig must be contiguous fp32")
        ops.dhead_prepare(wd, holder.weight_u, holder.weight_v, cache["w"], cache["aux"], need_iter, cache["scratch"])
        cache["key"] = base
        if need_iter:
            cache["iter_key"] = it_key
    return cache


def head_weight(holder):
    """(1,512,16,16) weight_orig -> fp32 vector in (h,w,c) order"""
    return head_state(holder)["w"]
