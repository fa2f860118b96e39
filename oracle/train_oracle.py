"""CPU ORACLE for the training step (BASELINE config 4).  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Restates what the reference's autograd does for one optimisation step of run_nerf_batch.py:544-570 under the shipped
configs with the smooth term off: render in training mode (BatchNorm of the encoder trunk on batch statistics) ->
loss = img2mse(rgb, target) [+ img2mse(bkgd_msk, acc)] -> gradients of every live parameter.  Index stages (mask,
nearest vertices, canonical points, pixel coordinates) come from the pinned numpy stages of oracle/oracle.py and
carry no gradient -- no parameter sits upstream of them (skinning_field = correction_field = 0).
Pinned against the reference's own autograd by oracle/make_golden_train.py -> tests/golden/train_grads.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O

TRUNK_KEYS = ["encoder_2d.model.conv1.weight", "encoder_2d.model.bn1.weight", "encoder_2d.model.bn1.bias"] + \
             [f"encoder_2d.model.layer1.{b}.{n}" for b in range(3) for n in ("conv1.weight", "bn1.weight", "bn1.bias",
                                                                            "conv2.weight", "bn2.weight", "bn2.bias")]


def dense_keys():
    return ([f"transformer.layers.{l}.{k}" for l in range(2) for k in (
        "0.fn.norm.weight", "0.fn.norm.bias", "0.fn.fn.to_qkv.weight", "0.fn.fn.to_out.0.weight",
        "0.fn.fn.to_out.0.bias", "1.fn.norm.weight", "1.fn.norm.bias", "1.fn.fn.net.0.weight",
        "1.fn.fn.net.0.bias", "1.fn.fn.net.3.weight", "1.fn.fn.net.3.bias")]
        + [f"pts_linears.{i}.{k}" for i in range(8) for k in ("weight", "bias")]
        + ["alpha_linear.weight", "alpha_linear.bias", "feature_linear.weight", "feature_linear.bias",
           "views_linear.weight", "views_linear.bias", "rgb_linear.weight", "rgb_linear.bias"])


def encode_images_train(img_all, sd):
    """encoder trunk (ref lib/encoder.py:256-306) with BatchNorm in training mode: batch statistics, eps 1e-5."""
    p = "encoder_2d.model."
    x = F.interpolate(img_all, scale_factor=0.5, mode="area", recompute_scale_factor=True)

    def bn(x, n):
        return F.batch_norm(x, None, None, sd[p + n + ".weight"], sd[p + n + ".bias"], True, 0.0, 1e-5)

    x = F.relu(bn(F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3), "bn1"))
    lat0 = x
    for b in range(3):
        y = F.relu(bn(F.conv2d(x, sd[p + f"layer1.{b}.conv1.weight"], None, 1, 1), f"layer1.{b}.bn1"))
        y = bn(F.conv2d(y, sd[p + f"layer1.{b}.conv2.weight"], None, 1, 1), f"layer1.{b}.bn2")
        x = F.relu(x + y)
    return torch.cat([lat0, x], dim=1)


def loss_and_grads(smpl, sd, sp_b, tp_b, rays_o, rays_d, near, far, S, target_rgb, bkgd_msk=None, u=None, eval_bn=False):
    """-> (loss float, {name: grad numpy} over dense_keys() + TRUNK_KEYS, forward outputs dict)."""
    sp, tp = O.squeeze_inputs(sp_b, tp_b)
    names = dense_keys() + TRUNK_KEYS
    par = {k: (v.clone().float().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    c = O.frame_constants(smpl, sp, tp)
    latent = O.encode_images(sp["img_all"], par) if eval_bn else encode_images_train(sp["img_all"], par)
    z = O.sample_z(near.reshape(-1), far.reshape(-1), S, u)
    pts = O.sample_points(rays_o.astype(np.float32), rays_d.astype(np.float32), z).reshape(-1, 3)
    q_all = O.world_to_smpl(pts, c["Th_tp"], c["R_tp"])
    d2, idx_all = O.knn1(q_all, c["verts_smpl"])
    mask = d2 < O.THRESH
    act = np.nonzero(mask)[0]
    xc = O.target2canonical(q_all[act], idx_all[act], c)
    _, _, xw, _ = O.canonical2source(xc, c)
    uv = O.projection(xw, sp["R_all"], sp["T_all"], sp["K_all"])
    H, W = sp["img_all"].shape[-2:]
    feat = O.bilinear_border(latent, uv, (W, H))
    rgbs = O.bilinear_border(sp["img_all"], uv, (W, H))
    V = rgbs.shape[0]
    tok = torch.cat((feat, O.posenc(rgbs.reshape(-1, 3), 4).reshape(V, -1, 27)), dim=-1).transpose(0, 1).contiguous()
    lin = O._Lin(False)
    tout = O.transformer(tok, par, lin)
    rgb, alpha = O.nerf_mlp(torch.from_numpy(xc), tout[:, 0], tout[:, 1], par, lin)
    raw = torch.full((len(pts), 4), -80.0)
    raw = raw.index_put((torch.from_numpy(act),), torch.cat([rgb, alpha], -1))
    N = len(rays_o)
    rgb_map, disp, acc, w, depth = O.raw2outputs(raw.reshape(N, S, 4), torch.from_numpy(z), torch.from_numpy(rays_d.astype(np.float32)))
    loss = torch.mean((rgb_map - torch.as_tensor(target_rgb)) ** 2)
    if bkgd_msk is not None:
        loss = loss + torch.mean((torch.as_tensor(bkgd_msk) - acc) ** 2)
    grads = torch.autograd.grad(loss, [par[k] for k in names], allow_unused=True)
    out = {"rgb_map": rgb_map.detach().numpy(), "acc_map": acc.detach().numpy(), "n_active": len(act),
           "d_latent_norm": None}
    return float(loss.detach()), {k: (None if g is None else g.numpy()) for k, g in zip(names, grads)}, out
