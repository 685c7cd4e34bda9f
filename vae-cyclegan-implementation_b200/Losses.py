"""Loss modules with the reference's names, signatures and return tuples (Losses.py:14-121 of the
reference), computed by the fused sm_100a loss kernels (csrc/losses.cu).

The reference's composite loss classes (its Losses.py:126-379) are dead code there -- never
referenced by Networks.py/train.py and with stale signatures -- and are
intentionally not reproduced."""
from __future__ import annotations

import torch.nn as nn

from .functions import KlFn, L1MeanFn, MseConstFn


class TranslationLoss(nn.Module):
    """L_trans = mean |G(x) - y|."""

    def forward(self, generated, target):
        return L1MeanFn.apply(generated, target)


class CycleConsistencyLoss(nn.Module):
    """L_cycle = mean|F(G(x)) - x| + mean|G(F(y)) - y|."""

    def forward(self, x, y, FGx, GFy):
        return L1MeanFn.apply(FGx, x) + L1MeanFn.apply(GFy, y)


class IdentityLoss(nn.Module):
    """L_id = mean|F(x) - x| + mean|G(y) - y|."""

    def forward(self, x, y, Fx, Gy):
        return L1MeanFn.apply(Fx, x) + L1MeanFn.apply(Gy, y)


class GANLossGenerator(nn.Module):
    """LSGAN generator side: real -> 0, fake -> 1; returns (total, real, fake)."""

    def forward(self, D_real, D_fake):
        real = MseConstFn.apply(D_real, 0.0)
        fake = MseConstFn.apply(D_fake, 1.0)
        return real + fake, real, fake


class GANLossDiscriminator(nn.Module):
    """LSGAN discriminator side: real -> 1, fake -> 0; returns (total, real, fake)."""

    def forward(self, D_real, D_fake):
        real = MseConstFn.apply(D_real, 1.0)
        fake = MseConstFn.apply(D_fake, 0.0)
        return real + fake, real, fake


class KLDivergenceLoss(nn.Module):
    """KL(q(z|x) || N(0, I)) with logvar clamped to [-10, 10], mean over all elements."""

    def forward(self, mu, logvar):
        return KlFn.apply(mu, logvar)
