"""Spectral normalisation as a forward-pre-hook (reference: spectral_norm.py:9-261, the vendored copy of
torch.nn.utils.spectral_norm).  Same API and state layout: `<name>_orig` parameter, `<name>_u` /
`<name>_v` buffers, version-1 state_dict metadata; `spectral_norm(module)` / `remove_spectral_norm`.

The power iteration (two mat-vecs + two normalisations + sigma) and the W/sigma scaling run as CUDA
kernels over the fp32 master weight; every rank performs the identical iteration on identical
weights, so no collective is needed (the property spectral_norm.py:57-60 relies on)."""
import torch
from torch.nn.functional import normalize

from . import ops

__all__ = ["SpectralNorm", "spectral_norm", "remove_spectral_norm"]


class SpectralNorm(object):
    _version = 1

    def __init__(self, name="weight", n_power_iterations=1, dim=0, eps=1e-12):
        if n_power_iterations <= 0:
            raise ValueError("Expected n_power_iterations to be positive, but got n_power_iterations={}".format(n_power_iterations))
        self.name, self.dim, self.n_power_iterations, self.eps = name, dim, n_power_iterations, eps

    def reshape_weight_to_matrix(self, weight):
        w = weight
        if self.dim != 0:
            w = w.permute(self.dim, *[d for d in range(w.dim()) if d != self.dim])
        return w.reshape(w.size(0), -1)

    def compute_weight(self, module, do_power_iteration):
        weight = getattr(module, self.name + "_orig")
        u = getattr(module, self.name + "_u")
        v = getattr(module, self.name + "_v")
        if self.dim != 0:
            raise NotImplementedError("spectral_norm: only dim=0 (Conv2d / Linear) is on the seg-GAN path")
        # u, v are updated in place by the kernel (spectral_norm.py:73-80); clones are kept for backward
        return ops.spectral_weight(weight, u, v, self.eps, bool(do_power_iteration), self.n_power_iterations)

    def remove(self, module):
        with torch.no_grad():
            weight = self.compute_weight(module, do_power_iteration=False)
        delattr(module, self.name)
        delattr(module, self.name + "_u")
        delattr(module, self.name + "_v")
        delattr(module, self.name + "_orig")
        module.register_parameter(self.name, torch.nn.Parameter(weight.detach()))

    def __call__(self, module, inputs):
        setattr(module, self.name, self.compute_weight(module, do_power_iteration=module.training))

    def _solve_v_and_rescale(self, weight_mat, u, target_sigma):
        v = torch.linalg.multi_dot([weight_mat.t().mm(weight_mat).pinverse(), weight_mat.t(), u.unsqueeze(1)]).squeeze(1)
        return v.mul_(target_sigma / torch.dot(u, torch.mv(weight_mat, v)))

    @staticmethod
    def apply(module, name, n_power_iterations, dim, eps):
        for hook in module._forward_pre_hooks.values():
            if isinstance(hook, SpectralNorm) and hook.name == name:
                raise RuntimeError("Cannot register two spectral_norm hooks on the same parameter {}".format(name))
        fn = SpectralNorm(name, n_power_iterations, dim, eps)
        weight = module._parameters[name]
        with torch.no_grad():
            h, w = fn.reshape_weight_to_matrix(weight).size()
            u = normalize(weight.new_empty(h).normal_(0, 1), dim=0, eps=fn.eps)
            v = normalize(weight.new_empty(w).normal_(0, 1), dim=0, eps=fn.eps)
        delattr(module, fn.name)
        module.register_parameter(fn.name + "_orig", weight)
        setattr(module, fn.name, weight.data)
        module.register_buffer(fn.name + "_u", u)
        module.register_buffer(fn.name + "_v", v)
        module.register_forward_pre_hook(fn)
        module._register_state_dict_hook(SpectralNormStateDictHook(fn))
        module._register_load_state_dict_pre_hook(SpectralNormLoadStateDictPreHook(fn))
        return fn


class SpectralNormLoadStateDictPreHook(object):
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        fn = self.fn
        version = local_metadata.get("spectral_norm", {}).get(fn.name + ".version", None)
        if version is None or version < 1:
            with torch.no_grad():
                weight_orig = state_dict[prefix + fn.name + "_orig"]
                weight = state_dict.pop(prefix + fn.name)
                sigma = (weight_orig / weight).mean()
                weight_mat = fn.reshape_weight_to_matrix(weight_orig)
                u = state_dict[prefix + fn.name + "_u"]
                state_dict[prefix + fn.name + "_v"] = fn._solve_v_and_rescale(weight_mat, u, sigma)


class SpectralNormStateDictHook(object):
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, module, state_dict, prefix, local_metadata):
        meta = local_metadata.setdefault("spectral_norm", {})
        key = self.fn.name + ".version"
        if key in meta:
            raise RuntimeError("Unexpected key in metadata['spectral_norm']: {}".format(key))
        meta[key] = self.fn._version


def spectral_norm(module, name="weight", n_power_iterations=1, eps=1e-12, dim=None):
    if dim is None:
        dim = 1 if isinstance(module, (torch.nn.ConvTranspose1d, torch.nn.ConvTranspose2d, torch.nn.ConvTranspose3d)) else 0
    SpectralNorm.apply(module, name, n_power_iterations, dim, eps)
    return module


def remove_spectral_norm(module, name="weight"):
    for k, hook in module._forward_pre_hooks.items():
        if isinstance(hook, SpectralNorm) and hook.name == name:
            hook.remove(module)
            del module._forward_pre_hooks[k]
            return module
    raise ValueError("spectral_norm of '{}' not found in {}".format(name, module))


def apply_spectral_norm_to_discriminator(disc):
    """north_star's "spectral-norm discriminator": wrap every conv / linear of `Discriminator`."""
    from .nn_layers import Conv2d, Linear
    for m in disc.modules():
        if isinstance(m, (Conv2d, Linear)) and "weight" in m._parameters:
            spectral_norm(m)
    return disc
