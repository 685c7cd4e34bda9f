"""The reference's network API (its Networks.py:57-2150) on B200 kernels.

Same class names, constructor arguments, forward arities / return orders, state_dict keys and
training_step / validation_step metric keys as the reference, so train.py, utils.py-style checkpoint
code and user scripts keep working; underneath, every module *emits* nodes into a plan (plan.py)
that runs hand-written sm_100a kernels through the C ABI.  No torch conv / norm / loss op is ever
called and there is no CPU fallback.

Reference quirks that are reproduced on purpose (parity, not a bug fix):
  * ReLU comes BEFORE InstanceNorm in D / U / R.conv1 (Networks.py:94,111,129);
  * VAEGAN's discriminator loss uses DGx.detach(), so only D(y) trains D (Networks.py:1280);
  * CycleVAEGAN's generator loss uses only the *fake* LSGAN terms, CycleAEGAN's uses real+fake
    (Networks.py:2012-2014 vs 1743-1748);
  * the constructor RNG order (nested re-initialisation) so that seeded construction matches."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lanes as _lanes
from . import lib as _L
from .Losses import (CycleConsistencyLoss, GANLossDiscriminator, GANLossGenerator, IdentityLoss,
                     KLDivergenceLoss, TranslationLoss)
from .functions import require_cuda, run_plan
from . import functions as _fn
from .optim import FusedAdam
from .plan import Plan, PlanBuilder, no_wgrad, _STATE
from . import plan as _plan

_ACTS = {"ReLU": _L.ACT_RELU, "LeakyReLU": _L.ACT_LEAKY, "Identity": _L.ACT_NONE, "Tanh": _L.ACT_TANH,
         "Sigmoid": _L.ACT_SIGMOID}


# ====================================================================== parameter holders
class ConvParams(nn.Module):
    """Holds one convolution's master parameters (fp32 OIHW weight + bias) under the names the
    reference's nn.Conv2d uses, initialised with the same RNG consumption (torch/nn/modules/conv.py
    reset_parameters: kaiming_uniform_(a=sqrt(5)) on the weight, then the bias)."""
    _vcg_holder = True

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
        nn.init.uniform_(self.bias, -bound, bound)

    def init_weight(self):
        return self.weight


class SpectralConvParams(nn.Module):
    """spectral_norm(nn.Conv2d(512, 1, 16)) of the reference (Networks.py:248): parameters `bias`,
    `weight_orig`, buffers `weight_u`, `weight_v` in the reference's state_dict order."""
    _vcg_holder = True

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        b = torch.empty(out_channels)
        bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
        nn.init.uniform_(b, -bound, bound)
        self.bias = nn.Parameter(b)
        self.weight_orig = nn.Parameter(w)
        self.register_buffer("weight_u", F.normalize(torch.empty(out_channels).normal_(0, 1), dim=0, eps=1e-12))
        self.register_buffer("weight_v", F.normalize(torch.empty(w[0].numel()).normal_(0, 1), dim=0, eps=1e-12))

    def init_weight(self):
        return self.weight_orig

    @torch.no_grad()
    def power_iteration(self):
        """One power iteration as every training forward of the reference does
        (torch/nn/utils/spectral_norm.py:92-114), by one kernel launch (csrc/losses.cu, vcg_dhead_prepare).  With a
        1 x K matrix it converges in one step: v = +-W/|W|, u = +-1, so sigma = |W| and the head kernels' unit-vector
        form is exact.  u is a normalised 1-vector, i.e. exactly +-1, so repeating the iteration on unchanged W, u, v
        reproduces the same bits: the reference's further iterations within one step (one per discriminator call,
        Networks.py:1916-1919, 2032-2035) are skipped (plan.head_state keeps the key)."""
        _plan.head_state(self, iterate=True)


def _kaiming_init(module, nonlinearity="relu", a=0.0):
    if isinstance(module, (ConvParams, SpectralConvParams)):
        nn.init.kaiming_normal_(module.init_weight(), a=a, mode="fan_out", nonlinearity=nonlinearity)
        nn.init.zeros_(module.bias)


# ====================================================================== plan plumbing
class _PlanModule(nn.Module):
    """Base: caches one plan per (entry point, input shapes) and runs it."""

    def _plan(self, key, shapes, build):
        cache = self.__dict__.setdefault("_vcg_plans", {})
        k = (key, tuple(shapes))
        p = cache.get(k)
        if p is None:
            b = PlanBuilder(shapes[0][0])
            build(b, [b.input(s[1], s[2], s[3]) for s in shapes])
            p = cache[k] = Plan(b)
        return p

    def _run(self, key, inputs, build, eps=()):
        for x in inputs:
            if x.dim() != 4:
                raise ValueError("expected NCHW input")
        return run_plan(self._plan(key, [tuple(x.shape) for x in inputs], build), list(inputs), eps)

    def forward(self, x):          # single-input single-output blocks
        def build(b, ins):
            b.output(self.emit(b, ins[0]))
        return self._run("fwd", [x], build)[0]


# ====================================================================== atomic blocks
class CaSb(_PlanModule):
    """conv -> [InstanceNorm] -> activation (Networks.py:57-81)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=3, activation="ReLU", use_norm=True):
        super().__init__()
        if activation not in ("ReLU", "LeakyReLU", "Tanh", "Sigmoid", "Identity"):
            raise NotImplementedError("Activation not implemented")
        if stride not in (1, 2) or (stride == 2 and kernel_size % 2):
            raise NotImplementedError("CaSb: stride must be 1, or 2 with an even kernel")
        self.conv = ConvParams(in_channels, out_channels, kernel_size)
        self.stride, self.padding, self.act, self.use_norm = stride, padding, _ACTS[activation], use_norm

    def emit(self, b, a):
        mode = _L.MODE_PLAIN if self.stride == 1 else _L.MODE_PAD_S2D
        if self.use_norm:
            return b.conv(self.conv, a, mode, self.padding, _L.ACT_NONE, True, self.act)
        if self.act in (_L.ACT_TANH, _L.ACT_SIGMOID):
            # not fused in the conv epilogues (no shipped network uses them): applied by the consumer's transform pass
            return b.conv(self.conv, a, mode, self.padding, _L.ACT_NONE, False, self.act)
        return b.conv(self.conv, a, mode, self.padding, self.act, False, _L.ACT_NONE)


class D(_PlanModule):
    """PixelUnshuffle(2) -> conv3x3 -> ReLU -> InstanceNorm (Networks.py:83-96)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = ConvParams(in_channels * 4, out_channels, 3)

    def emit(self, b, a):
        return b.conv(self.conv, a, _L.MODE_UNSHUFFLE, 1, _L.ACT_RELU, True)


class R(_PlanModule):
    """conv -> ReLU -> IN -> conv -> IN -> + input (Networks.py:98-116)."""

    def __init__(self, out_channels):
        super().__init__()
        self.conv1 = ConvParams(out_channels, out_channels, 3)
        self.conv2 = ConvParams(out_channels, out_channels, 3)

    def emit(self, b, a):
        h = b.conv(self.conv1, a, _L.MODE_PLAIN, 1, _L.ACT_RELU, True)
        h = b.conv(self.conv2, h, _L.MODE_PLAIN, 1, _L.ACT_NONE, True)
        return b.residual(h, a)


class U(_PlanModule):
    """PixelShuffle(2) -> conv3x3 -> ReLU -> InstanceNorm (Networks.py:118-131)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = ConvParams(in_channels // 4, out_channels, 3)

    def emit(self, b, a):
        return b.conv(self.conv, a, _L.MODE_SHUFFLE, 1, _L.ACT_RELU, True)


class S(_PlanModule):
    """bare reflect-padded conv3x3 (Networks.py:133-140)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = ConvParams(in_channels, out_channels, 3)

    def emit(self, b, a):
        return b.conv(self.conv, a, _L.MODE_PLAIN, 1)


class L(S):
    """identical to S in the reference (Networks.py:142-149)."""


# ====================================================================== molecular networks
class Encoder(_PlanModule):
    def __init__(self):
        super().__init__()
        self.model = nn.Sequential(CaSb(3, 64, kernel_size=7, stride=1), D(64, 128), D(128, 256), D(256, 512),
                                   D(512, 1024), R(1024))
        self.apply(_kaiming_init)

    def emit(self, b, a):
        for m in self.model:
            a = m.emit(b, a)
        return a


class Decoder(_PlanModule):
    def __init__(self):
        super().__init__()
        self.model = nn.Sequential(R(1024), U(1024, 512), U(512, 256), U(256, 128), U(128, 64),
                                   CaSb(64, 3, kernel_size=7, stride=1, activation="Identity", use_norm=False))
        self.apply(_kaiming_init)

    def emit(self, b, a):
        for m in self.model:
            a = m.emit(b, a)
        return a


class VariationalEncoderBlock(_PlanModule):
    def __init__(self, in_channels, latent_dim=64):
        super().__init__()
        self.muConv = L(in_channels, latent_dim)
        self.logvarConv = nn.Sequential(S(in_channels, latent_dim), S(latent_dim, latent_dim))

    def emit(self, b, a):
        mu = self.muConv.emit(b, a)
        lv = self.logvarConv[1].emit(b, self.logvarConv[0].emit(b, a))
        z = b.reparam(mu, lv)
        return z, z.producer

    def forward(self, x):
        """-> (z, mu, clamped logvar); eps is drawn with torch.randn like the reference (Networks.py:225)."""
        def build(b, ins):
            z, node = self.emit(b, ins[0])
            b.output(z)
            b.output(node, "mu")
            b.output(node, "logvar")
        n, _, h, w = x.shape
        eps = torch.randn(n, self.muConv.conv.weight.shape[0], h, w, device=x.device)
        return self._run("fwd", [x], build, [eps])


class VariationalDecoderBlock(_PlanModule):
    def __init__(self, latent_dim=64, out_channels=1024):
        super().__init__()
        self.conv = S(latent_dim, out_channels)

    def emit(self, b, a):
        return self.conv.emit(b, a)


class Discriminator(_PlanModule):
    """Networks.py:240-269: 4 x (k4 s2 reflect conv [+IN] + LeakyReLU 0.2) + spectral-normed 16x16 head."""

    def __init__(self):
        super().__init__()
        self.model = nn.Sequential(
            CaSb(3, 64, kernel_size=4, stride=2, padding=1, activation="LeakyReLU", use_norm=False),
            CaSb(64, 128, kernel_size=4, stride=2, padding=1, activation="LeakyReLU"),
            CaSb(128, 256, kernel_size=4, stride=2, padding=1, activation="LeakyReLU"),
            CaSb(256, 512, kernel_size=4, stride=2, padding=1, activation="LeakyReLU"),
            SpectralConvParams(512, 1, 16))
        self.apply(lambda m: _kaiming_init(m, "leaky_relu", 0.2))

    def emit(self, b, a):
        for m in list(self.model)[:4]:
            a = m.emit(b, a)
        head = self.model[4]
        if (a.h, a.w) != tuple(head.weight_orig.shape[2:]):
            raise ValueError(f"Discriminator: the {head.weight_orig.shape[2]}x{head.weight_orig.shape[3]} head needs a "
                             f"256x256 input (got a {a.h}x{a.w} feature map)")
        return b.head(head, a)

    def forward(self, x):
        head = self.model[4]
        if self.training:
            head.power_iteration()
        out = self._run("fwd", [x], lambda b, ins: self.emit(b, ins[0]))[0]
        if not self.training:
            # eval: sigma from the stale u, v of the last training forward (spectral_norm.py:125-130); the head kernel
            # divided by |W|: rescale by |W| / sigma (both from vcg_dhead_prepare)
            aux = _plan.head_state(head)["aux"]
            out = head.bias.detach() + (out - head.bias.detach()) * (aux[1] / aux[0])
        return out


# ====================================================================== composite base
class _Composite(_PlanModule):
    """Shared training-loop contract of the reference's composites (Networks.py:9-44)."""
    _two_optimizers = False

    def _adam(self, params, lr, betas, owners=None):
        """owners: the sub-networks whose backward passes training_step announces to FusedAdam.track() -- their
        gradient buckets are unpacked / all-reduced / applied while the rest of the step still runs (optim.py)."""
        opt = FusedAdam(list(params), lr=lr, betas=betas)
        if owners:
            opt.set_owners(owners)
        return opt

    def save_optimizer_states(self):
        if self._two_optimizers:
            if self.optimizer_G is None or self.optimizer_D is None:
                raise ValueError("Optimizers have not been configured yet.")
            return {"optimizer_G": self.optimizer_G.state_dict(), "optimizer_D": self.optimizer_D.state_dict()}
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        return {"optimizer": self.optimizer.state_dict()}

    def load_optimizer_states(self, states):
        names = ("optimizer_G", "optimizer_D") if self._two_optimizers else ("optimizer",)
        for n in names:
            if getattr(self, n) is None:
                raise ValueError("Optimizer has not been configured yet.")
        for n in names:
            if n not in states:
                raise KeyError(f"{n} state not found in states")
            getattr(self, n).load_state_dict(states[n])

    def enable_debug_mode(self, enabled=True):
        self.debug_mode = enabled

    def _noise_calls(self, x, y=None):
        """[(generator, input)] of the bottleneck-noise draws one forward() makes, in the reference's order"""
        return []

    def skip_forward_noise(self, x, y=None):
        """Consume the random stream exactly like forward(x, y) would, without computing it (train.py's epoch loop
        skips the reference's per-batch display forward, train.py:112-117, but keeps its RNG consumption)."""
        _draw_eps(self._noise_calls(x, y))

    def _prepack(self):
        """Refresh every derived weight copy of the model on the CALLING stream, before the lanes fork (lanes.py):
        one multi-tensor pack launch for all stale filters, the discriminator heads' (h, w, c) weight vectors and
        their spectral-norm power iteration.  Returns True when every convolution has been planned before, i.e.
        when no pass of this step will have to create (and pack) a kernel-layout copy on a lane."""
        self._finish_steps()            # the previous step's updates and re-packs ran on the optimisers' side streams
        mods = self.__dict__.get("_vcg_param_mods")
        if mods is None:
            mods = self.__dict__["_vcg_param_mods"] = [m for m in self.modules() if isinstance(m, (ConvParams, SpectralConvParams))]
        warm, pcs = True, []
        for m in mods:
            if isinstance(m, SpectralConvParams):
                if m.weight_orig.is_cuda:
                    _plan.head_state(m, iterate=self.training)
                continue
            cache = m.__dict__.get("_vcg_packed")
            if not cache:
                warm = False
            else:
                pcs.extend(cache.values())
        if pcs and pcs[0].holder.weight.is_cuda:
            _plan.refresh_many(pcs, _STATE["dtype"])
        return warm

    def _dual(self, device, fresh=True):
        """lanes.dual region for this model's passes.  fresh=True (forward): prepack first; lanes are used from the
        second step on (the first one creates and packs the kernel-layout copies inside the passes).
        fresh=False (backward of the same step): same setting as the forward."""
        if fresh:
            self.__dict__["_vcg_dual"] = self._prepack() and torch.device(device).type == "cuda"
        return _lanes.dual(device, self.__dict__.get("_vcg_dual", False))

    def _finish_steps(self):
        """complete optimiser steps whose gradient exchange was left running (data-parallel overlap, dist.py)"""
        for n in ("optimizer", "optimizer_G", "optimizer_D"):
            o = getattr(self, n, None)
            if o is not None and hasattr(o, "finish"):
                o.finish()

    def _items(self, named):
        """dict of 0-d tensors -> dict of python floats with ONE device synchronisation (the reference
        pays one .item() sync per metric, Networks.py:2054-2076); averaged over ranks when data-parallel."""
        self._finish_steps()
        _fn.end_step()
        keys = list(named)
        vals = torch.stack([named[k].detach().float().reshape(()) for k in keys])
        sync = getattr(self, "_vcg_sync", None)
        if sync is not None and sync.world > 1:
            import torch.distributed as dist
            dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=sync.group)
            vals = vals / sync.world
        sink = getattr(self, "_vcg_capture", None)
        if sink is not None:            # CUDA-graph capture (graph.py): no host read inside the captured region
            sink["keys"], sink["vals"] = keys, vals
            return {}
        return dict(zip(keys, vals.tolist()))

    @staticmethod
    def _xy(batch):
        x, y = batch["x"], batch["y"]
        require_cuda(x, "batch['x']")
        _fn.begin_step(x.device)          # one zero-fill for all loss accumulators of the step (functions.py)
        return x, y


_EPS_SOURCE = [None]


def set_eps_source(fn):
    """Replace the noise draw of every VAE bottleneck: fn(shape, device) -> eps tensor (or None to
    restore torch.randn on the device).  Parity tests feed the reference's CPU-generator noise."""
    _EPS_SOURCE[0] = fn


def _vae_eps(vae, x):
    n, _, h, w = x.shape
    shape = (n, vae.latent_dim, h // 16, w // 16)
    if _EPS_SOURCE[0] is not None:
        return _EPS_SOURCE[0](shape, x.device)
    return torch.randn(*shape, device=x.device)


# ====================================================================== single generators
class Autoencoder(_Composite):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.optimizer = None
        self.loss_fn = None
        self.apply(_kaiming_init)

    def emit(self, b, a):
        return self.decoder.emit(b, self.encoder.emit(b, a))

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999), decoder_only=False):
        self.optimizer = self._adam(self.decoder.parameters() if decoder_only else self.parameters(), lr, betas, [self])
        return self.optimizer

    def configure_loss(self, **kwargs):
        self.loss_fn = TranslationLoss()

    def training_step(self, batch):
        if self.loss_fn is None:
            raise ValueError("Loss function has not been configured yet.")
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        x, y = self._xy(batch)
        loss = self.loss_fn(self(x), y)
        # the reference's NaN/Inf guard (Networks.py:357-372) is a host synchronisation: it is kept in eager mode and
        # skipped while the step is being captured into a CUDA graph (graph.GraphedStep), where no host read is legal
        capturing = getattr(self, "_vcg_capture", None) is not None
        if not capturing and not bool(torch.isfinite(loss)):
            self.optimizer.zero_grad()
            nan = float("nan")
            return {"nan_detected": True, "G_loss": nan, "loss_trans": nan, "total_loss": nan}
        self.optimizer.zero_grad()
        with self.optimizer.track({self: 1}):
            loss.backward()
        self.optimizer.step()
        return self._items({"G_loss": loss, "loss_trans": loss, "total_loss": loss})

    def validation_step(self, batch):
        if self.loss_fn is None:
            raise ValueError("Loss function has not been configured yet.")
        with torch.no_grad():
            x, y = self._xy(batch)
            out = self(x)
            v = self.loss_fn(out, y).item()
            return {"G_loss": v, "total_loss": v, "loss_trans": v, "Gx": out}


class VariationalAutoencoder(_Composite):
    def __init__(self, latent_dim=64):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder = Encoder()
        self.variational_encoder_block = VariationalEncoderBlock(in_channels=1024, latent_dim=latent_dim)
        self.variational_decoder_block = VariationalDecoderBlock(latent_dim=latent_dim, out_channels=1024)
        self.decoder = Decoder()
        self.optimizer = None
        self.loss_trans_fn = None
        self.loss_kl_fn = None
        self.lambda_kl = 0
        self.apply(_kaiming_init)

    def emit(self, b, a):
        """-> (Gx act, reparam node)"""
        z, node = self.variational_encoder_block.emit(b, self.encoder.emit(b, a))
        return self.decoder.emit(b, self.variational_decoder_block.emit(b, z)), node

    def forward(self, x, eps=None):
        """-> (Gx, mu, logvar).  `eps` (optional, shape (B, latent, H/16, W/16)) replaces the
        torch.randn draw so parity tests can feed the reference's noise."""
        def build(b, ins):
            g, node = self.emit(b, ins[0])
            b.output(g)
            b.output(node, "mu")
            b.output(node, "logvar")
        if eps is None:
            eps = _vae_eps(self, x)
        return self._run("fwd", [x], build, [eps])

    def _noise_calls(self, x, y=None):
        return [(self, x)]

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999)):
        self.optimizer = self._adam(self.parameters(), lr, betas, [self])
        return self.optimizer

    def configure_loss(self, **kwargs):
        self.loss_trans_fn = TranslationLoss()
        self.loss_kl_fn = KLDivergenceLoss()
        self.lambda_kl = kwargs.get("lambda_kl", 1e-5)

    def _check(self):
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        if self.loss_trans_fn is None:
            raise ValueError("Translation loss function has not been configured yet.")
        if self.loss_kl_fn is None:
            raise ValueError("KL divergence loss function has not been configured yet.")

    def _losses(self, x, y):
        out, mu, lv = self(x)
        lt, lk = self.loss_trans_fn(out, y), self.loss_kl_fn(mu, lv)
        return out, lt, lk, lt + self.lambda_kl * lk

    def training_step(self, batch):
        self._check()
        x, y = self._xy(batch)
        _, lt, lk, G_loss = self._losses(x, y)
        self.optimizer.zero_grad()
        with self.optimizer.track({self: 1}):
            G_loss.backward()
        self.optimizer.step()
        return self._items({"G_loss": G_loss, "loss_trans": lt, "loss_kl": lk})

    def validation_step(self, batch):
        self._check()
        with torch.no_grad():
            x, y = self._xy(batch)
            out, lt, lk, total = self._losses(x, y)
            m = self._items({"G_loss": total, "loss_trans": lt, "loss_kl": lk})
            m["Gx"] = out
            return m


def _gen_call(g, x, eps=None):
    """(Gx, mu, logvar) for a VAE generator, (Gx,) for an AE generator.  `eps`: pre-drawn bottleneck noise."""
    out = g(x) if eps is None else g(x, eps=eps)
    return out if isinstance(out, tuple) else (out,)


def _draw_eps(calls):
    """Bottleneck noise of several VAE passes, drawn up front IN THE REFERENCE'S ORDER (one torch.randn_like per
    pass, Networks.py:223-226 as called from :1909-1914), so that the passes themselves can be issued in any order
    and on any lane.  calls: [(generator, input tensor)]; AE generators draw nothing."""
    return [_vae_eps(g, x) if getattr(g, "latent_dim", None) else None for g, x in calls]


# ====================================================================== generator + discriminator
class _PairedGAN(_Composite):
    _two_optimizers = True

    def configure_optimizers(self, lr=2e-4, betas=(0.5, 0.999)):
        self.optimizer_G = self._adam(self.G.parameters(), lr, betas, [self.G])
        self.optimizer_D = self._adam(self.D.parameters(), lr, betas, [self.D])
        return self.optimizer_G, self.optimizer_D

    def _require(self):
        if self.optimizer_G is None or self.optimizer_D is None:
            raise ValueError("Optimizers have not been configured yet.")


class AEGAN(_PairedGAN):
    def __init__(self):
        super().__init__()
        self.G = Autoencoder()
        self.D = Discriminator()
        self.apply(_kaiming_init)
        self.optimizer_G = self.optimizer_D = None
        self.loss_trans_fn = self.loss_gan_gen_fn = self.loss_gan_disc_fn = self.loss_identity_fn = None
        self.lambda_gan = self.lambda_identity = 0

    def forward(self, x, y):
        with self._dual(x.device) as d:          # lane 0: G(x), D(Gx), D(y); lane 1: G(y)
            with d.lane(0):
                Gx = self.G(x)
            with d.lane(1):
                Gy = self.G(y)
            with d.lane(0):
                DGx, Dy = self.D(Gx), self.D(y)
        return Gx, Gy, DGx, Dy

    def configure_loss(self, **kwargs):
        self.loss_trans_fn = TranslationLoss()
        self.loss_gan_gen_fn = GANLossGenerator()
        self.loss_gan_disc_fn = GANLossDiscriminator()
        self.loss_identity_fn = TranslationLoss()
        self.lambda_gan = kwargs.get("lambda_gan", 1.0)
        self.lambda_identity = kwargs.get("lambda_identity", 5.0)

    def _g_losses(self, x, y):
        Gx, Gy, DGx, Dy = self(x, y)
        lt = self.loss_trans_fn(Gx, y)
        lg, lg_real, lg_fake = self.loss_gan_gen_fn(Dy, DGx)
        lid = self.loss_identity_fn(Gy, y)
        return Gx, DGx, Dy, lt, lg, lg_real, lg_fake, lid, lt + self.lambda_gan * lg + self.lambda_identity * lid

    def training_step(self, batch):
        self._require()
        if None in (self.loss_trans_fn, self.loss_gan_gen_fn, self.loss_gan_disc_fn, self.loss_identity_fn):
            raise ValueError("Loss functions have not been configured yet.")
        x, y = self._xy(batch)
        self.optimizer_G.zero_grad()
        Gx, DGx, Dy, lt, lg, _, _, lid, G_loss = self._g_losses(x, y)
        # (D grads of the G step are discarded by the reference; G(x) and G(y) both produce G's gradients)
        with self._dual(x.device, fresh=False), no_wgrad(self.D), self.optimizer_G.track({self.G: 2}):
            G_loss.backward(retain_graph=True)
        self.optimizer_G.step()
        # D step: the reference re-runs D(Gx.detach()) and D(y) (Networks.py:1110-1112); D's weights have
        # not changed, so the activations of the first forward are reused and only D's backward runs again
        self.optimizer_D.zero_grad()
        D_loss, D_real, D_fake = self.loss_gan_disc_fn(Dy, DGx)
        with _stop_at_inputs(), self.optimizer_D.track({self.D: 2}):
            D_loss.backward()
        self.optimizer_D.step()
        return self._items({"G_loss": G_loss, "D_loss": D_loss, "D_loss_real": D_real, "D_loss_fake": D_fake,
                            "loss_trans": lt, "loss_gan_g": lg, "loss_identity": lid,
                            "d_y_mean": Dy.mean(), "d_gx_mean": DGx.mean()})

    def validation_step(self, batch):
        self._require()
        with torch.no_grad():
            x, y = self._xy(batch)
            Gx, DGx, Dy, lt, lg, lg_real, lg_fake, lid, G_loss = self._g_losses(x, y)
            D_loss, D_real, D_fake = self.loss_gan_disc_fn(Dy, DGx)
            m = self._items({"total_loss": G_loss + D_loss, "G_loss": G_loss, "D_loss": D_loss, "D_loss_real": D_real,
                             "D_loss_fake": D_fake, "loss_trans": lt, "loss_gan_g": lg, "loss_gan_g_real": lg_real,
                             "loss_gan_g_fake": lg_fake, "loss_identity": lid})
            m["Gx"] = Gx
            return m


class VAEGAN(_PairedGAN):
    def __init__(self, latent_dim=64):
        super().__init__()
        self.G = VariationalAutoencoder(latent_dim)
        self.D = Discriminator()
        self.latent_dim = latent_dim
        self.debug_mode = False
        self.debug_info = {}
        self.optimizer_G = self.optimizer_D = None

    def _noise_calls(self, x, y=None):
        return [(self.G, x), (self.G, y)]

    def forward(self, x, y):
        e = _draw_eps(self._noise_calls(x, y))
        with self._dual(x.device) as d:          # lane 0: G(x), D(Gx), D(y); lane 1: G(y)
            with d.lane(0):
                Gx, mu, lv = self.G(x, eps=e[0])
            with d.lane(1):
                Gy, mu_y, lv_y = self.G(y, eps=e[1])
            with d.lane(0):
                DGx, Dy = self.D(Gx), self.D(y)
        return Gx, mu, lv, Gy, mu_y, lv_y, DGx, Dy

    def configure_loss(self, **kwargs):
        self.translation_loss = TranslationLoss()
        self.gan_loss_gen = GANLossGenerator()
        self.gan_loss_disc = GANLossDiscriminator()
        self.identity_loss = TranslationLoss()
        self.kl_loss = KLDivergenceLoss()
        self.lambda_gan = kwargs.get("lambda_gan", 1.0)
        self.lambda_identity = kwargs.get("lambda_identity", 5.0)
        self.lambda_kl = kwargs.get("lambda_kl", 1e-5)
        self.lambda_recon = kwargs.get("lambda_recon", 1.0)

    def _g_losses(self, x, y):
        Gx, mu, lv, Gy, _, _, DGx, Dy = self(x, y)
        lt = self.translation_loss(Gx, y)
        lg, lg_real, lg_fake = self.gan_loss_gen(Dy, DGx)
        lid = self.identity_loss(Gy, y)
        lk = self.kl_loss(mu, lv)
        G_loss = self.lambda_recon * lt + self.lambda_gan * lg + self.lambda_identity * lid + self.lambda_kl * lk
        return Gx, DGx, Dy, lt, lg_real, lg_fake, lid, lk, G_loss

    def training_step(self, batch):
        self._require()
        x, y = self._xy(batch)
        Gx, DGx, Dy, lt, lg_real, lg_fake, lid, lk, G_loss = self._g_losses(x, y)
        D_loss, D_real, D_fake = self.gan_loss_disc(Dy, DGx.detach())     # only D(y) trains D (Networks.py:1280)
        self.optimizer_G.zero_grad()
        # (D grads of the G step are zeroed by optimizer_D.zero_grad() in the reference)
        with self._dual(x.device, fresh=False), no_wgrad(self.D), self.optimizer_G.track({self.G: 2}):
            G_loss.backward(retain_graph=True)
        self.optimizer_G.step()
        self.optimizer_D.zero_grad()
        with _stop_at_inputs(), self.optimizer_D.track({self.D: 1}):     # DGx is detached: only D(y) contributes
            D_loss.backward()
        self.optimizer_D.step()
        m = self._items({"G_loss": G_loss, "D_loss": D_loss, "loss_gan_disc_real": D_real, "loss_gan_disc_fake": D_fake,
                         "loss_trans": lt, "loss_gan_real": lg_real, "loss_gan_fake": lg_fake, "loss_identity": lid,
                         "loss_kl": lk})
        if self.debug_mode:
            m["debug_info"] = self.debug_info
        return m

    def validation_step(self, batch):
        self._require()
        with torch.no_grad():
            x, y = self._xy(batch)
            Gx, DGx, Dy, lt, lg_real, lg_fake, lid, lk, G_loss = self._g_losses(x, y)
            D_loss, _, _ = self.gan_loss_disc(Dy, DGx)
            m = self._items({"total_loss": G_loss + D_loss, "G_loss": G_loss, "D_loss": D_loss, "loss_trans": lt,
                             "loss_gan_real": lg_real, "loss_gan_fake": lg_fake, "loss_identity": lid, "loss_kl": lk})
            m["Gx"] = Gx
            return m


class _stop_at_inputs:
    """Discriminator step: run the discriminators' backward again on the saved activations but do not
    propagate into the generators (equivalent to the reference's .detach() + re-forward)."""

    def __enter__(self):
        self.prev = _STATE.get("input_grads", True)
        _STATE["input_grads"] = False

    def __exit__(self, *exc):
        _STATE["input_grads"] = self.prev


# ====================================================================== cycle models
class _Cycle(_Composite):
    pass


class CycleAE(_Cycle):
    _vae = False

    def __init__(self, paired=True, latent_dim=64):
        super().__init__()
        if self._vae:
            self.F = VariationalAutoencoder(latent_dim)
            self.G = VariationalAutoencoder(latent_dim)
        else:
            self.F = Autoencoder()
            self.G = Autoencoder()
        self.paired = paired
        self.optimizer = None
        self.loss_cycle = None
        self.loss_trans = None
        self.loss_kl = None
        self.lambda_cycle = 0

    def _noise_calls(self, x, y=None):
        # the reference's order: G(x), F(Gx), F(y), G(Fy) (Networks.py:1489-1494); Gx has x's shape, Fy has y's
        return [(self.G, x), (self.F, x), (self.F, y), (self.G, y)]

    def forward(self, x, y):
        e = _draw_eps(self._noise_calls(x, y))
        with self._dual(x.device) as d:          # lane 0: x -> G -> F; lane 1: y -> F -> G
            with d.lane(0):
                gx = _gen_call(self.G, x, e[0])
            with d.lane(1):
                fy = _gen_call(self.F, y, e[2])
            with d.lane(0):
                fgx = _gen_call(self.F, gx[0], e[1])
            with d.lane(1):
                gfy = _gen_call(self.G, fy[0], e[3])
        if not self._vae:
            return gx[0], fgx[0], fy[0], gfy[0]
        return (gx[0], fgx[0], fy[0], gfy[0], gx[1], gx[2], fgx[1], fgx[2], fy[1], fy[2], gfy[1], gfy[2])

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999)):
        self.optimizer = self._adam(self.parameters(), lr, betas, [self.F, self.G])
        return self.optimizer

    def configure_loss(self, **kwargs):
        self.loss_cycle = CycleConsistencyLoss()
        if self.paired:
            self.loss_trans = TranslationLoss()
        if self._vae:
            self.loss_kl = KLDivergenceLoss()
            self.lambda_kl = kwargs.get("lambda_kl", 1e-5)
        self.lambda_cycle = kwargs.get("lambda_cycle", 10.0)

    def _losses(self, x, y):
        if self.loss_cycle is None or (self._vae and self.loss_kl is None):
            raise ValueError("Loss functions have not been configured yet.")
        if self.paired and self.loss_trans is None:
            raise ValueError("Translation loss not configured for paired mode.")
        o = self(x, y)
        Gx, FGx, Fy, GFy = o[:4]
        lc = self.loss_cycle(x, y, FGx, GFy)
        total = self.lambda_cycle * lc
        named = {"loss_cycle": lc}
        if self._vae:
            lk = (self.loss_kl(o[4], o[5]) + self.loss_kl(o[6], o[7]) + self.loss_kl(o[8], o[9]) + self.loss_kl(o[10], o[11]))
            total = total + self.lambda_kl * lk
            named["loss_kl"] = lk
        if self.paired:
            lt = self.loss_trans(Gx, y) + self.loss_trans(Fy, x)
            total = total + lt
            named["loss_trans"] = lt
        named["total_loss"] = named["G_loss"] = total
        return Gx, Fy, total, named

    def training_step(self, batch):
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        x, y = self._xy(batch)
        _, _, total, named = self._losses(x, y)
        self.optimizer.zero_grad()
        with self._dual(x.device, fresh=False), self.optimizer.track({self.F: 2, self.G: 2}):
            total.backward()
        self.optimizer.step()
        return self._items(named)

    def validation_step(self, batch):
        with torch.no_grad():
            x, y = self._xy(batch)
            Gx, Fy, _, named = self._losses(x, y)
            m = self._items(named)
            m["Gx"], m["Fy"] = Gx.detach(), Fy.detach()
            return m


class CycleVAE(CycleAE):
    _vae = True

    def __init__(self, latent_dim=64, paired=True):
        super().__init__(paired=paired, latent_dim=latent_dim)


class CycleAEGAN(_Cycle):
    _two_optimizers = True
    _vae = False

    def __init__(self, paired=True, latent_dim=64):
        super().__init__()
        if self._vae:
            self.F = VariationalAutoencoder(latent_dim)
            self.G = VariationalAutoencoder(latent_dim)
        else:
            self.F = Autoencoder()
            self.G = Autoencoder()
        self.DX = Discriminator()
        self.DY = Discriminator()
        self.paired = paired
        self.apply(_kaiming_init)
        self.debug_mode = False
        self.debug_info = {}
        self.optimizer_G = self.optimizer_D = None
        self.loss_cycle = self.loss_gan_gen = self.loss_gan_disc = self.loss_kl = self.loss_identity = None
        # unpaired mode: G(y) and F(x) are computed by the reference but never used (Networks.py:1911,
        # 1914 vs 2016); skip_dead_passes drops them (their eps is still drawn, so RNG order is kept)
        self.skip_dead_passes = True

    def forward(self, x, y):
        return self._forward(x, y, skip_dead=False)

    def _noise_calls(self, x, y=None):
        # the reference's order: G(x), G(y), F(Gx), F(y), F(x), G(Fy) (Networks.py:1909-1914)
        return [(self.G, x), (self.G, y), (self.F, x), (self.F, y), (self.F, x), (self.G, y)]

    def _forward(self, x, y, skip_dead):
        dead = skip_dead and not self.paired
        # the dead passes' draws are made (and dropped) so that the RNG stream is the reference's
        e = _draw_eps(self._noise_calls(x, y))
        none3 = (None, None, None)
        with self._dual(x.device) as d:
            # lane 0: x -> G -> F, F(x), both DY passes; lane 1: y -> F -> G, G(y), both DX passes
            with d.lane(0):
                gx = _gen_call(self.G, x, e[0])
            with d.lane(1):
                fy = _gen_call(self.F, y, e[3])
            with d.lane(0):
                fgx = _gen_call(self.F, gx[0], e[2])
            with d.lane(1):
                gfy = _gen_call(self.G, fy[0], e[5])
            with d.lane(1):
                gy = none3 if dead else _gen_call(self.G, y, e[1])
            with d.lane(0):
                fx = none3 if dead else _gen_call(self.F, x, e[4])
            with d.lane(0):
                DYGx, DYy = self.DY(gx[0]), self.DY(y)
            with d.lane(1):
                DXFy, DXx = self.DX(fy[0]), self.DX(x)
        if not self._vae:
            return gx[0], fgx[0], fy[0], gfy[0], DYGx, DXFy, DXx, DYy, gy[0], fx[0]
        return (gx[0], fgx[0], fy[0], gfy[0], gx[1], gx[2], fgx[1], fgx[2], fy[1], fy[2], gfy[1], gfy[2],
                DYGx, DXFy, DXx, DYy, gy[0], fx[0])

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999)):
        self.optimizer_G = self._adam(list(self.F.parameters()) + list(self.G.parameters()), lr, betas, [self.F, self.G])
        self.optimizer_D = self._adam(list(self.DX.parameters()) + list(self.DY.parameters()), lr, betas, [self.DX, self.DY])
        return self.optimizer_G, self.optimizer_D

    def configure_loss(self, **kwargs):
        self.loss_cycle = CycleConsistencyLoss()
        self.loss_gan_gen = GANLossGenerator()
        self.loss_gan_disc = GANLossDiscriminator()
        if self.paired:
            self.loss_identity = IdentityLoss()
        if self._vae:
            self.loss_kl = KLDivergenceLoss()
            self.lambda_kl = kwargs.get("lambda_kl", 1e-5)
        self.lambda_gan = kwargs.get("lambda_gan", 1.0)
        self.lambda_identity = kwargs.get("lambda_identity", 5.0)
        self.lambda_cycle = kwargs.get("lambda_cycle", 10.0)

    def _g_losses(self, x, y, skip_dead):
        if self.loss_cycle is None or self.loss_gan_gen is None or self.loss_gan_disc is None or (self._vae and self.loss_kl is None):
            raise ValueError("Loss functions have not been configured yet.")
        if self.paired and self.loss_identity is None:
            raise ValueError("Identity loss not configured for paired mode.")
        o = self._forward(x, y, skip_dead)
        Gx, FGx, Fy, GFy = o[:4]
        DYGx, DXFy, DXx, DYy, Gy, Fx = o[-6:]
        lc = self.loss_cycle(x, y, FGx, GFy)
        lgx, lgx_r, lgx_f = self.loss_gan_gen(DXx, DXFy)
        lgy, lgy_r, lgy_f = self.loss_gan_gen(DYy, DYGx)
        named = {"loss_cycle": lc, "loss_gan_g_x_real": lgx_r, "loss_gan_g_x_fake": lgx_f,
                 "loss_gan_g_y_real": lgy_r, "loss_gan_g_y_fake": lgy_f}
        if self._vae:       # fake terms only (Networks.py:2006-2014)
            lg = lgx_f + lgy_f
            lk = (self.loss_kl(o[4], o[5]) + self.loss_kl(o[6], o[7]) + self.loss_kl(o[8], o[9]) + self.loss_kl(o[10], o[11]))
            G_loss = self.lambda_cycle * lc + self.lambda_gan * lg + self.lambda_kl * lk
            named["loss_kl"] = lk
        else:               # real + fake (Networks.py:1741-1748)
            lg = lgx + lgy
            G_loss = self.lambda_cycle * lc + self.lambda_gan * lg
        named["loss_gan_g"] = lg
        if self.paired:
            lid = self.loss_identity(x, y, Fx, Gy)
            G_loss = G_loss + self.lambda_identity * lid
            named["loss_identity"] = lid
        named["G_loss"] = G_loss
        return Gx, Fy, (DYGx, DXFy, DXx, DYy), G_loss, named

    def _d_losses(self, d, named):
        DYGx, DXFy, DXx, DYy = d
        ldx, dxr, dxf = self.loss_gan_disc(DXx, DXFy)
        ldy, dyr, dyf = self.loss_gan_disc(DYy, DYGx)
        D_loss = ldx + ldy
        named.update(D_loss=D_loss, D_loss_x_real=dxr, D_loss_x_fake=dxf, D_loss_y_real=dyr, D_loss_y_fake=dyf)
        return D_loss

    def training_step(self, batch):
        if self.optimizer_G is None or self.optimizer_D is None:
            raise ValueError("Optimizers have not been configured yet.")
        x, y = self._xy(batch)
        self.optimizer_G.zero_grad()
        _, _, d, G_loss, named = self._g_losses(x, y, self.skip_dead_passes)
        n_gen = 3 if self.paired else 2         # passes per generator that reach the loss (G(y), F(x): identity term only)
        with self._dual(x.device, fresh=False), no_wgrad(self.DX, self.DY), self.optimizer_G.track({self.F: n_gen, self.G: n_gen}):
            G_loss.backward(retain_graph=True)
        self.optimizer_G.step()
        # discriminators: same weights, same inputs => the four D forwards of Networks.py:2032-2035 would
        # reproduce the activations already saved; run only their backward, stopping at the D inputs
        self.optimizer_D.zero_grad()
        D_loss = self._d_losses(d, named)
        with self._dual(x.device, fresh=False), _stop_at_inputs(), self.optimizer_D.track({self.DX: 2, self.DY: 2}):
            D_loss.backward()
        self.optimizer_D.step()
        DYGx, DXFy, DXx, DYy = d
        named.update(total_loss=G_loss + D_loss, d_x_real_mean=DXx.mean(), d_x_fake_mean=DXFy.mean(),
                     d_y_real_mean=DYy.mean(), d_y_fake_mean=DYGx.mean())
        return self._items(named)

    def validation_step(self, batch):
        with torch.no_grad():
            x, y = self._xy(batch)
            Gx, Fy, d, G_loss, named = self._g_losses(x, y, False)
            D_loss = self._d_losses(d, named)
            named["total_loss"] = G_loss + D_loss
            m = self._items(named)
            m["Gx"], m["Fy"] = Gx.detach(), Fy.detach()
            return m


class CycleVAEGAN(CycleAEGAN):
    _vae = True

    def __init__(self, latent_dim=64, paired=True):
        super().__init__(paired=paired, latent_dim=latent_dim)


# ====================================================================== shared-encoder pretraining models
class DoubleAutoencoder(_Composite):
    """Shared encoder, two decoders (Networks.py:415-605)."""

    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder_A = Decoder()
        self.decoder_B = Decoder()
        self.optimizer = None
        self.loss_fn = None

    def _pass(self, x, dec, key):
        return self._run(key, [x], lambda b, ins: b.output(dec.emit(b, self.encoder.emit(b, ins[0]))))[0]

    def forward(self, x, y):
        return self._pass(x, self.decoder_A, "A"), self._pass(y, self.decoder_B, "B")

    def translate_A_to_B(self, x):
        return self._pass(x, self.decoder_B, "B")

    def translate_B_to_A(self, y):
        return self._pass(y, self.decoder_A, "A")

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999)):
        self.optimizer = self._adam(self.parameters(), lr, betas)
        return self.optimizer

    def configure_loss(self, **kwargs):
        self.loss_fn = TranslationLoss()

    def _losses(self, x, y):
        if self.loss_fn is None:
            raise ValueError("Loss function has not been configured yet.")
        Gx, Gy = self(x, y)
        la, lb = self.loss_fn(Gx, x), self.loss_fn(Gy, y)
        return la + lb, {"G_loss": la + lb, "loss_recon_A": la, "loss_recon_B": lb, "total_loss": la + lb}

    def training_step(self, batch):
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        x, y = self._xy(batch)
        total, named = self._losses(x, y)
        self.optimizer.zero_grad()
        total.backward()
        self.optimizer.step()
        return self._items(named)

    def validation_step(self, batch):
        with torch.no_grad():
            x, y = self._xy(batch)
            _, named = self._losses(x, y)
            m = self._items(named)
            m["Gx"], m["Fy"] = self.translate_A_to_B(x), self.translate_B_to_A(y)
            return m

    def create_cycle_ae(self):
        c = CycleAE()
        c.G.encoder.load_state_dict(self.encoder.state_dict())
        c.G.decoder.load_state_dict(self.decoder_B.state_dict())
        c.F.encoder.load_state_dict(self.encoder.state_dict())
        c.F.decoder.load_state_dict(self.decoder_A.state_dict())
        return c.to(next(self.parameters()).device)


class DoubleVariationalAutoencoder(_Composite):
    """Shared encoder, two VAE bottlenecks, two decoders (Networks.py:608-852)."""

    def __init__(self, latent_dim=64):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder = Encoder()
        self.vae_encoder_block_A = VariationalEncoderBlock(in_channels=1024, latent_dim=latent_dim)
        self.vae_encoder_block_B = VariationalEncoderBlock(in_channels=1024, latent_dim=latent_dim)
        self.vae_decoder_block_A = VariationalDecoderBlock(latent_dim=latent_dim, out_channels=1024)
        self.vae_decoder_block_B = VariationalDecoderBlock(latent_dim=latent_dim, out_channels=1024)
        self.decoder_A = Decoder()
        self.decoder_B = Decoder()
        self.optimizer = None
        self.loss_trans_fn = self.loss_kl_fn = None
        self.lambda_kl = 0
        self.apply(_kaiming_init)

    def _pass(self, x, which):
        veb = getattr(self, "vae_encoder_block_" + which)
        vdb = getattr(self, "vae_decoder_block_" + which)
        dec = getattr(self, "decoder_" + which)

        def build(b, ins):
            z, node = veb.emit(b, self.encoder.emit(b, ins[0]))
            b.output(dec.emit(b, vdb.emit(b, z)))
            b.output(node, "mu")
            b.output(node, "logvar")
        return self._run(which, [x], build, [_vae_eps(self, x)])

    def _noise_calls(self, x, y=None):
        return [(self, x), (self, y)]

    def forward(self, x, y):
        Gx, mu_x, lv_x = self._pass(x, "A")
        Gy, mu_y, lv_y = self._pass(y, "B")
        return Gx, Gy, mu_x, lv_x, mu_y, lv_y

    def translate_A_to_B(self, x):
        return self._pass(x, "B")[0]

    def translate_B_to_A(self, y):
        return self._pass(y, "A")[0]

    def configure_optimizers(self, lr=1e-4, betas=(0.5, 0.999)):
        self.optimizer = self._adam(self.parameters(), lr, betas)
        return self.optimizer

    def configure_loss(self, **kwargs):
        self.loss_trans_fn = TranslationLoss()
        self.loss_kl_fn = KLDivergenceLoss()
        self.lambda_kl = kwargs.get("lambda_kl", 1e-5)

    def _losses(self, x, y):
        if self.loss_trans_fn is None or self.loss_kl_fn is None:
            raise ValueError("Loss functions have not been configured yet.")
        Gx, Gy, mu_x, lv_x, mu_y, lv_y = self(x, y)
        la, lb = self.loss_trans_fn(Gx, x), self.loss_trans_fn(Gy, y)
        ka, kb = self.loss_kl_fn(mu_x, lv_x), self.loss_kl_fn(mu_y, lv_y)
        total = la + lb + self.lambda_kl * (ka + kb)
        return total, {"G_loss": total, "loss_recon_A": la, "loss_recon_B": lb, "loss_kl": ka + kb, "loss_kl_A": ka,
                       "loss_kl_B": kb, "total_loss": total}

    def training_step(self, batch):
        if self.optimizer is None:
            raise ValueError("Optimizer has not been configured yet.")
        x, y = self._xy(batch)
        total, named = self._losses(x, y)
        self.optimizer.zero_grad()
        total.backward()
        self.optimizer.step()
        return self._items(named)

    def validation_step(self, batch):
        with torch.no_grad():
            x, y = self._xy(batch)
            _, named = self._losses(x, y)
            m = self._items(named)
            m["Gx"], m["Fy"] = self.translate_A_to_B(x), self.translate_B_to_A(y)
            return m

    def create_cycle_vae(self):
        c = CycleVAE(latent_dim=self.latent_dim)
        for g, w in ((c.G, "B"), (c.F, "A")):
            g.encoder.load_state_dict(self.encoder.state_dict())
            g.variational_encoder_block.load_state_dict(getattr(self, "vae_encoder_block_" + w).state_dict())
            g.variational_decoder_block.load_state_dict(getattr(self, "vae_decoder_block_" + w).state_dict())
            g.decoder.load_state_dict(getattr(self, "decoder_" + w).state_dict())
        return c.to(next(self.parameters()).device)
