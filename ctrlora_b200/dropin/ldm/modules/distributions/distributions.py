"""Drop-in for the reference's `ldm/modules/distributions/distributions.py`: the VAE posterior.

`sample()` consumes the host RNG exactly like the reference (`torch.randn(shape)` on the CPU generator, then moved to the
device, reference :35-37); the arithmetic mean + exp(0.5 * clamp(logvar)) * noise (and the latent scale factor, when the
LDM passes it) is one kernel, `ctrlora_gaussian_sample`."""
import numpy as np
import torch

from ctrlora_b200 import ops


class AbstractDistribution:
    def sample(self):
        raise NotImplementedError()

    def mode(self):
        raise NotImplementedError()


class DiracDistribution(AbstractDistribution):
    def __init__(self, value):
        self.value = value

    def sample(self):
        return self.value

    def mode(self):
        return self.value


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self._std = self._var = None

    # std / var are only materialised when somebody asks (kl / nll are training-time VAE losses, not on the CtrLoRA path)
    @property
    def std(self):
        if self._std is None:
            self._std = torch.zeros_like(self.mean) if self.deterministic else torch.exp(0.5 * self.logvar)
        return self._std

    @property
    def var(self):
        if self._var is None:
            self._var = torch.zeros_like(self.mean) if self.deterministic else torch.exp(self.logvar)
        return self._var

    def sample(self, scale=1.0):
        noise = torch.randn(self.mean.shape).to(device=self.parameters.device)
        if self.deterministic:
            return self.mode(scale)
        if self.parameters.is_cuda:
            return ops.gaussian_sample(self.parameters, noise, scale)
        return scale * (self.mean + self.std * noise)

    def mode(self, scale=1.0):
        if self.parameters.is_cuda and scale != 1.0:
            return ops.gaussian_sample(self.parameters, None, scale)
        return self.mean if scale == 1.0 else scale * self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar
                               + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.])
        logtwopi = np.log(2.0 * np.pi)
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)
