"""Image encoders: the 2-D trunk (boundary of the hot path, stays torch/cuDNN) and a
parameter-only stand-in for the spconv 3-D encoder.

``SpatialEncoder`` mirrors ``lib/encoder.py:186-306`` of the reference (ResNet34 conv1 +
bn1 + relu + layer1 on the half-resolution image, concatenated).  ``index`` is served by the
CUDA gather kernel (csrc/gather.cu); the torch version here is the standalone-module
behaviour (bilinear, border, align_corners=True == the reference's hand-written gather).
"""
import torch
import torch.nn.functional as F
from torch import nn


def grid_sample(image, optical):
    """Reference-compatible entry (lib/encoder.py:12-62): optical in [-1,1], (N,H,W,2)."""
    return F.grid_sample(image, optical, mode="bilinear", padding_mode="border", align_corners=True)


class SpatialEncoder(nn.Module):
    def __init__(self, backbone="resnet34", pretrained=True, num_layers=4, index_interp="bilinear",
                 feature_scale=0.5, use_first_pool=False):
        super().__init__()
        import torchvision
        self.feature_scale = feature_scale
        self.use_first_pool = use_first_pool
        weights = None
        if pretrained:
            # ImageNet weights only when they are already cached locally (no download attempt
            # offline); otherwise random init -- checkpoints overwrite them anyway
            import os
            try:
                w = torchvision.models.get_model_weights(backbone).DEFAULT
                cached = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
                if os.path.exists(cached):
                    self.model = torchvision.models.get_model(backbone, weights=w)
                    weights = w
            except Exception:
                weights = None
        if weights is None:
            self.model = torchvision.models.get_model(backbone, weights=None)
        self.latent_size = [0, 64, 128, 256, 512, 1024][num_layers]
        self.num_layers = num_layers
        self.index_interp = index_interp
        self.latent = None

    def forward(self, x):
        if self.feature_scale != 1.0:
            x = F.interpolate(x, scale_factor=self.feature_scale,
                              mode="bilinear" if self.feature_scale > 1.0 else "area",
                              align_corners=True if self.feature_scale > 1.0 else None,
                              recompute_scale_factor=True)
        m = self.model
        x = m.relu(m.bn1(m.conv1(x)))
        latents = [x]
        stages = [m.layer1, m.layer2, m.layer3, m.layer4]
        for i in range(self.num_layers - 1):
            if i == 0 and self.use_first_pool:
                x = m.maxpool(x)
            x = stages[i](x)
            latents.append(x)
        sz = latents[0].shape[-2:]
        # same-size align_corners=True resize is the identity; deeper levels are upsampled
        latents = [l if l.shape[-2:] == sz else F.interpolate(l, sz, mode="bilinear", align_corners=True)
                   for l in latents]
        self.latent = torch.cat(latents, dim=1)
        return self.latent

    def index(self, uv, image_size=(512, 512)):
        """uv (B,N,2) pixels -> (B,L,N) features of the last encoded images."""
        size = torch.as_tensor(image_size, dtype=torch.float32, device=uv.device)
        g = 2.0 * uv.unsqueeze(2).float() / size - 1.0
        return grid_sample(self.latent.float(), g)[:, :, :, 0]


class _SpConvWeight(nn.Module):
    """Holds the ``weight`` of one spconv 3x3x3 convolution (spconv-2 layout: out, k, k, k, in)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, 3, 3, 3, cin).normal_(0, 0.02))


def _sp_seq(cin, cout, n):
    """Same child indices as the reference's SparseSequential: (conv, BatchNorm1d, ReLU) * n."""
    mods = []
    for i in range(n):
        mods += [_SpConvWeight(cin if i == 0 else cout, cout), nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01), nn.ReLU()]
    return nn.Sequential(*mods)


class SparseConvNet(nn.Module):
    """State-dict stand-in for the reference's spconv pyramid (lib/encoder.py:367-527).

    Only reached when correction_field or skinning_field is set, which both shipped configs
    disable (configs/canonical_transformer.txt:25,49); kept constructible so that checkpoints
    load (keys ``conv0.0.weight``, ``conv0.1.running_mean`` ...).
    """

    def __init__(self, num_layers=2):
        super().__init__()
        self.num_layers = num_layers
        self.conv0 = _sp_seq(3, 16, 2)
        self.down0 = _sp_seq(16, 32, 1)
        self.conv1 = _sp_seq(32, 32, 2)
        self.down1 = _sp_seq(32, 64, 1)
        self.conv2 = _sp_seq(64, 64, 3)
        self.down2 = _sp_seq(64, 128, 1)
        self.conv3 = _sp_seq(128, 128, 3)
        self.down3 = _sp_seq(128, 128, 1)
        self.conv4 = _sp_seq(128, 128, 3)

    def forward(self, *a, **k):
        raise NotImplementedError("SparseConvNet is outside the render hot path (correction_field/skinning_field = 0)")
