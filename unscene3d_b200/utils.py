"""Small host-side helpers shared by bench.py, smoke() and the examples."""
import zlib

import torch


class BackboneConfig:
    """The two fields the backbones read from the hydra node (conf/model/mask3d.yaml:36-47)."""

    def __init__(self, bn_momentum=0.02, conv1_kernel_size=3, dilations=(1, 1, 1, 1)):
        self.bn_momentum = bn_momentum
        self.conv1_kernel_size = conv1_kernel_size
        self.dilations = list(dilations)


def seeded_state(module: torch.nn.Module, seed: int = 0):
    """Name-keyed deterministic weights: every tensor is drawn from its own generator seeded by
    (seed, crc32(name)), so two implementations with the same state-dict names get identical
    weights regardless of construction order.  Convolution kernels ~ U(+-1.4*sqrt(3/fan_in)),
    BatchNorm affine ~ U(0.5,1.5) / U(-0.2,0.2), running stats perturbed."""
    out = {}
    for name, t in module.state_dict().items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros_like(t)
        elif name.endswith("running_mean"):
            out[name] = (torch.rand(t.shape, generator=g) - 0.5) * 0.2
        elif name.endswith("running_var"):
            out[name] = 0.5 + torch.rand(t.shape, generator=g)
        elif ".bn.weight" in name or name.endswith("norm.weight"):
            out[name] = 0.5 + torch.rand(t.shape, generator=g)
        elif ".bn.bias" in name or name.endswith("norm.bias"):
            out[name] = (torch.rand(t.shape, generator=g) - 0.5) * 0.4
        else:
            fan = t.shape[-2] * (t.shape[0] if t.ndim == 3 else 1) if t.ndim >= 2 else max(t.numel(), 1)
            bound = (3.0 / fan) ** 0.5 * 1.4
            out[name] = (torch.rand(t.shape, generator=g) * 2 - 1) * bound
    return out


def conv_layer_bytes(n_in, n_out, kvol, cin, cout, kind="fwd", elem=4):
    """Algorithmic HBM bytes of one sparse-conv launch:
    fwd/dgrad: read the input rows once, write the output rows once, read the weights once;
    wgrad: read X and dY once, write dW once.  BN/ReLU/residual/cat and index traffic count zero."""
    if kind in ("fwd", "dgrad"):
        return elem * (n_in * cin + n_out * cout) + 4 * kvol * cin * cout
    return elem * (n_in * cin + n_out * cout) + 4 * kvol * cin * cout


def conv_layer_flops(pairs, cin, cout):
    return 2 * pairs * cin * cout
