"""torch-backed stand-in for the tensorflow_probability symbols on the reference's
hot path (model.py:395, signals.py:265-267).  TEST INFRASTRUCTURE ONLY -- see
oracle/tf_shim/tensorflow/__init__.py."""
import types

import numpy as np
import scipy.stats as _st
import tensorflow as tf
import torch


def _clip_by_value_preserve_gradient(t, clip_value_min, clip_value_max):
    # tfp.math.clip_by_value_preserve_gradient: t + stop_gradient(clip(t) - t)
    clipped = torch.clamp(t, clip_value_min, clip_value_max)
    return t + (clipped - t).detach()


math = types.SimpleNamespace(clip_by_value_preserve_gradient=_clip_by_value_preserve_gradient)


class _TruncatedNormal:
    def __init__(self, loc, scale, low, high):
        self.loc, self.scale, self.low, self.high = loc, scale, low, high

    def sample(self, shape):
        n = int(np.prod(shape))
        u = torch.rand(n, dtype=torch.float64)
        tf.random.LOG.append(('truncnorm_u', u.clone()))
        a, b = (self.low - self.loc) / self.scale, (self.high - self.loc) / self.scale
        x = _st.truncnorm.ppf(u.numpy(), a, b, loc=self.loc, scale=self.scale)
        return torch.as_tensor(x, dtype=torch.float32).reshape(tuple(shape))


class _StudentT:
    def __init__(self, df, loc, scale):
        self.df, self.loc, self.scale = float(df), loc, scale

    def log_prob(self, x):
        import math as _m
        df = self.df
        y = (x - self.loc) / self.scale
        c = _m.lgamma(0.5 * (df + 1.0)) - _m.lgamma(0.5 * df) - 0.5 * _m.log(df * _m.pi)
        return c - torch.log(self.scale) - 0.5 * (df + 1.0) * torch.log1p(y * y / df)


class _LogitNormal:
    """Only kl_divergence is used (model.py:695-698).  TFP evaluates the KL of a bijected distribution as the KL of
    the base Normals (_kl_normal_normal): 0.5*((mu_a-mu_b)/s_b)^2 + 0.5*expm1(2*log(s_a/s_b)) - log(s_a/s_b)."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def kl_divergence(self, other):
        d = torch.log(self.scale) - torch.log(other.scale)
        return 0.5 * (self.loc / other.scale - other.loc / other.scale) ** 2 + 0.5 * torch.expm1(2.0 * d) - d


class _InverseGamma:
    def __init__(self, concentration, scale):
        self.a, self.b = tf._t(concentration).float(), tf._t(scale).float()

    def log_prob(self, x):
        return self.a * torch.log(self.b) - torch.lgamma(self.a) - (self.a + 1.0) * torch.log(x) - self.b / x


distributions = types.SimpleNamespace(TruncatedNormal=_TruncatedNormal, StudentT=_StudentT, LogitNormal=_LogitNormal,
                                      InverseGamma=_InverseGamma)
layers = types.SimpleNamespace()
