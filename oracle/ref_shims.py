"""Import shims that let the UNMODIFIED reference (/root/reference) run on this CPU box.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).  None of the shims
changes reference arithmetic except ``pytorch3d.ops.knn.knn_points``, which is
absent offline and is restated (brute force, d2 = (dx*dx+dy*dy)+dz*dz, ties ->
lowest index).  Recipe: SURVEY.md Appendix B.
"""
import argparse
import os
import pickle
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MPSNERF_REF", "/root/reference")
KNN_LOG = []          # (query, points, dist, idx) of every knn_points call, for stage capture


def _knn_points(p1, p2, K=1, **kw):
    q, v = p1[0].float(), p2[0].float()
    dists, idxs = [], []
    for s in range(0, len(q), 8192):
        c = q[s:s + 8192]
        dx, dy, dz = c[:, 0:1] - v[None, :, 0], c[:, 1:2] - v[None, :, 1], c[:, 2:3] - v[None, :, 2]
        d2 = dx * dx
        d2 += dy * dy
        d2 += dz * dz
        if K == 1:
            m, i = torch.min(d2, dim=1)
            m, i = m[:, None], i[:, None]
        else:
            m, i = torch.topk(d2, K, dim=1, largest=False, sorted=True)
        dists.append(m)
        idxs.append(i)
    d, i = torch.cat(dists)[None], torch.cat(idxs)[None]
    KNN_LOG.append((p1[0].detach().clone(), d[0, :, 0].clone(), i[0, :, 0].clone()))
    return d, i, None


class _ConfigArgParser(argparse.ArgumentParser):
    """argparse + ``is_config_file`` + ``key = value`` files (configargparse subset)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._cfg_dest = None

    def add_argument(self, *a, **k):
        if k.pop("is_config_file", False):
            act = super().add_argument(*a, **k)
            self._cfg_dest = act.dest
            return act
        return super().add_argument(*a, **k)

    def parse_known_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        pre = argparse.ArgumentParser(add_help=False)
        pre.add_argument("--config")
        cfg = pre.parse_known_args(args)[0].config
        file_args = []
        if cfg:
            for line in open(cfg):
                line = line.split("#")[0].strip()
                if not line or "=" not in line:
                    continue
                key, val = (s.strip() for s in line.split("=", 1))
                if val in ("True", "true"):
                    file_args.append("--" + key)
                elif val in ("False", "false"):
                    continue
                else:
                    file_args += ["--" + key, val]
        return super().parse_known_args(file_args + args, namespace)


def install(workdir, smpl_np):
    """Install every shim; write synthetic SMPL pickles under ``workdir/assets``; chdir there."""
    import numpy.lib.npyio as npyio
    if not hasattr(npyio, "save"):
        npyio.save = np.save
    poly = types.ModuleType("numpy.lib.polynomial")
    poly.roots = np.roots
    sys.modules.setdefault("numpy.lib.polynomial", poly)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("memory_profiler", profile=lambda f=None, **k: f if f else (lambda g: g))
    mod("pytorch3d")
    mod("pytorch3d.ops")
    mod("pytorch3d.ops.knn", knn_points=_knn_points)

    class _Conv(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    class _Seq(torch.nn.Sequential):
        pass

    sp_attrs = dict(SubMConv3d=_Conv, SparseConv3d=_Conv, SparseSequential=_Seq, SparseConvTensor=object)
    spc = mod("spconv", **sp_attrs)
    spc.pytorch = mod("spconv.pytorch", **sp_attrs)
    mod("configargparse", ArgumentParser=_ConfigArgParser)
    mod("trimesh")
    mod("imageio")
    mod("skimage")
    mod("skimage.measure", compare_ssim=lambda *a, **k: 0.0)
    if "torch.utils.tensorboard" not in sys.modules:
        try:
            import importlib
            importlib.import_module("torch.utils.tensorboard")
        except Exception:
            mod("torch.utils.tensorboard", SummaryWriter=object)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import torchvision
    _r34 = torchvision.models.resnet34
    torchvision.models.resnet34 = lambda pretrained=False, **k: _r34(weights=None, **k)

    os.makedirs(os.path.join(workdir, "assets"), exist_ok=True)
    for name in ("basicmodel_m_lbs_10_207_0_v1.0.0.pkl", "basicmodel_f_lbs_10_207_0_v1.0.0.pkl", "SMPL_NEUTRAL.pkl"):
        with open(os.path.join(workdir, "assets", name), "wb") as fh:
            pickle.dump(smpl_np, fh)
    os.makedirs(os.path.join(workdir, "logs", "THuman_3_view"), exist_ok=True)
    os.chdir(workdir)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load_reference(n_samples=64, config="canonical_transformer.txt", extra=()):
    """Import run_nerf_batch from the reference (it parses sys.argv at import)."""
    sys.argv = ["x", "--config", os.path.join(REF, "configs", config), "--N_samples", str(n_samples), *extra]
    import run_nerf_batch as R
    return R


class ScatterLike(torch.nn.Module):
    """Gives the net a ``.module`` and does what DataParallel's scatter / gather do around it (the reference's
    only multi-subject mode, run_nerf_batch.py:344-350): every tensor of the input dicts and the points are
    split along the batch dim into chunks of one subject, the module runs per chunk on a private copy of the
    dicts (the reference's ``sequeeze_0`` mutates them in place), and the outputs are concatenated."""

    def __init__(self, net):
        super().__init__()
        self.module = net

    @staticmethod
    def _take(d, b):
        out = {}
        for k, v in d.items():
            if isinstance(v, dict):
                out[k] = ScatterLike._take(v, b)
            elif torch.is_tensor(v) and v.dim() > 0:
                out[k] = v[b:b + 1].clone()
            else:
                out[k] = v
        return out

    def forward(self, sp, tp, pts, dirs):
        outs = []
        for b in range(pts.shape[0]):
            outs.append(self.module(self._take(sp, b), self._take(tp, b), pts[b:b + 1], None if dirs is None else dirs[b:b + 1]))
        return torch.cat(outs, 0)
