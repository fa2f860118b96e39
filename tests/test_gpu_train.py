"""GPU parity of the training step (BASELINE config 4): the hand-written backward kernels against torch autograd of
the oracle's functions, and the whole step against the reference's own autograd (tests/golden/train_grads.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda_dict(d):
    return {k: (v.cuda() if torch.is_tensor(v) else _cuda_dict(v) if isinstance(v, dict) else v) for k, v in d.items()}


def _net(scene, sd):
    from test_gpu_parity import make_net
    return make_net(scene, sd, "fp32")


@pytest.mark.parametrize("S,occupancy", [(1, 0), (7, 0), (64, 0), (64, 1), (130, 0)])
def test_composite_backward_against_autograd(S, occupancy):
    from mpsnerf_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    g = torch.Generator().manual_seed(S + occupancy)
    N = 200
    raw = torch.randn(N, S, 4, generator=g) * 4
    raw[:30] = -80.0                                   # empty rays
    raw[30:60, ::2] = -80.0                            # masked-out samples between active ones
    raw[60:80, :, 3] = 50.0                            # opaque
    z = torch.sort(torch.rand(N, S, generator=g) * 2 + 1, dim=-1).values
    d = torch.randn(N, 3, generator=g)
    d_rgb, d_acc = torch.randn(N, 3, generator=g), torch.randn(N, generator=g)
    x = raw.clone().requires_grad_(True)
    rgb, disp, acc, w, depth = O.raw2outputs(x, z, d, occupancy=bool(occupancy))
    ((rgb * d_rgb).sum() + (acc * d_acc).sum()).backward()
    want = x.grad.clone()
    if not occupancy:
        want[raw[..., 3] == -80.0] = 0.0               # the -80 fill of masked-out points is a constant, not a leaf
    rays8 = torch.zeros(N, 8)
    rays8[:, 3:6] = d
    got = torch.empty(N, S, 4, device="cuda")
    raw_d, rays_d, z_d, drgb_d, dacc_d = raw.cuda(), rays8.cuda(), z.cuda(), d_rgb.cuda(), d_acc.cuda()     # (kept alive)
    _lib.check(lib.mpsnerf_composite_bwd(_lib.ptr(raw_d), _lib.ptr(rays_d), N, S, None, None, _lib.ptr(z_d),
                                         occupancy, _lib.ptr(drgb_d), _lib.ptr(dacc_d), _lib.ptr(got), None), "composite_bwd")
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=2e-5 * max(scale, 1.0), rtol=2e-4)


def test_gather_backward_against_autograd():
    from mpsnerf_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    V, Hf, Wf, H, W, n = 3, 17, 23, 64, 96, 500
    lat = torch.randn(V, 128, Hf, Wf, generator=g, requires_grad=True)
    uv = torch.rand(V, n, 2, generator=g) * torch.tensor([W + 10.0, H + 10.0]) - 5.0          # some outside: border taps
    d_tok = torch.randn(n, V, 155, generator=g)
    feat = O.bilinear_border(lat, uv, (W, H))                                                 # (V, n, 128)
    (feat * d_tok[:, :, :128].transpose(0, 1)).sum().backward()
    want = lat.grad.permute(0, 2, 3, 1).contiguous()
    fr = _lib.Frame()
    fr.n_views, fr.img_w, fr.img_h, fr.feat_w, fr.feat_h = V, W, H, Wf, Hf
    frame = torch.frombuffer(bytearray(bytes(fr)), dtype=torch.uint8).cuda()
    got = torch.zeros(V, Hf, Wf, 128, device="cuda")
    uv_d = uv.transpose(0, 1).contiguous().cuda()                                             # (n, V, 2)
    dtok_d = d_tok.cuda()
    _lib.check(lib.mpsnerf_gather_tokens_bwd(_lib.ptr(uv_d), n, V, _lib.ptr(frame), _lib.ptr(dtok_d), 155, _lib.ptr(got), None),
               "gather_tokens_bwd")
    torch.cuda.synchronize()
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("V,n", [(3, 700), (2, 65), (4, 130)])
def test_dense_train_forward_backward_against_autograd(V, n):
    """mpsnerf_dense_train_fwd / _bwd on random tokens against torch autograd of the oracle's transformer + MLP."""
    import ctypes
    from mpsnerf_b200 import _lib, synthetic
    from mpsnerf_b200.engine import DENSE_FP32_ORDER
    from oracle import oracle as O
    lib = _lib.load()
    sd = synthetic.seeded_state_dict(5, 30.0)
    g = torch.Generator().manual_seed(n)
    tok = torch.randn(n, V, 155, generator=g)
    xc = torch.randn(n, 3, generator=g) * 0.5
    d_out = torch.randn(n, 4, generator=g)
    par = {k: sd[k].clone().requires_grad_(True) for k in DENSE_FP32_ORDER}
    t = tok.clone().requires_grad_(True)
    lin = O._Lin(False)
    tout = O.transformer(t, par, lin)
    rgb, alpha = O.nerf_mlp(xc, tout[:, 0], tout[:, 1], par, lin)
    out_ref = torch.cat([rgb, alpha], -1)
    (out_ref * d_out).sum().backward()
    w_dev = [sd[k].cuda().contiguous() for k in DENSE_FP32_ORDER]
    g_dev = [torch.zeros_like(w) for w in w_dev]
    wt = (ctypes.c_void_p * 46)(*[w.data_ptr() for w in w_dev])
    gt = (ctypes.c_void_p * 46)(*[x.data_ptr() for x in g_dev])
    ws = torch.empty(lib.mpsnerf_dense_train_workspace(n, V), dtype=torch.uint8, device="cuda")
    out4 = torch.empty(n, 4, device="cuda")
    tok_d, xc_d = tok.cuda(), xc.cuda()
    _lib.check(lib.mpsnerf_dense_train_fwd(_lib.ptr(tok_d), 155, _lib.ptr(xc_d), n, V, wt, _lib.ptr(out4), _lib.ptr(ws), None), "fwd")
    d_tok = torch.empty(n, V, 155, device="cuda")
    dout_d = d_out.cuda()
    _lib.check(lib.mpsnerf_dense_train_bwd(_lib.ptr(dout_d), n, V, wt, gt, _lib.ptr(d_tok), _lib.ptr(ws), None), "bwd")
    torch.cuda.synchronize()
    scale = float(out_ref.abs().max())
    np.testing.assert_allclose(out4.cpu().numpy(), out_ref.detach().numpy(), atol=2e-4 * max(1.0, scale))
    np.testing.assert_allclose(d_tok.cpu().numpy(), t.grad.numpy(), atol=2e-4 * float(t.grad.abs().max()) + 1e-6)
    for k, got in zip(DENSE_FP32_ORDER, g_dev):
        want = par[k].grad
        err = float((got.cpu() - want).abs().max()) / max(float(want.abs().max()), 1e-20)
        assert err <= 2e-3, (k, err)        # fp32 sums over the points in a different order (atomics, tiles)
    # gradients accumulate: a second backward doubles them
    _lib.check(lib.mpsnerf_dense_train_bwd(_lib.ptr(dout_d), n, V, wt, gt, _lib.ptr(d_tok), _lib.ptr(ws), None), "bwd")
    torch.cuda.synchronize()
    k0 = DENSE_FP32_ORDER.index("pts_linears.3.weight")
    np.testing.assert_allclose(g_dev[k0].cpu().numpy(), 2 * par["pts_linears.3.weight"].grad.numpy(),
                               atol=4e-3 * float(par["pts_linears.3.weight"].grad.abs().max()))


@pytest.fixture
def strict_fp32_convs():
    """cuDNN convolutions default to TF32 (forward and backward); the reference gradients were produced on CPU in
    fp32, so the trunk runs in true fp32 for the comparison."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _train_setup(spec=None):
    from golden_cases import build_train_case
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    from mpsnerf_b200.parser_config import config_parser
    scene, sd, ids, S, u, target, msk = build_train_case(spec)
    net = _net(scene, sd).train()
    R.configure(config_parser().parse_args(["--smooth_loss", "0"]))
    handle = R.NetworkHandle(net)
    rays, near, far = synthetic.rays_tensor(scene, ids, device="cuda")
    kw = dict(rays=rays, near=near, far=far, sp_input=_cuda_dict(scene.sp_input), tp_input=_cuda_dict(scene.tp_input),
              N_samples=S, perturb=1.0, perturb_u=torch.from_numpy(u)[None].cuda(), use_viewdirs=True)
    return R, net, handle, kw, torch.from_numpy(target)[None].cuda(), torch.from_numpy(msk)[None, :, None].cuda()


def test_training_step_gradients_against_reference_autograd(strict_fp32_convs):
    """render() in training mode + loss.backward(): every live parameter's gradient against what the UNMODIFIED
    reference's autograd produced for the same step (and the forward outputs against its render)."""
    from test_train_oracle import GOLD, check_grads_against_golden
    from oracle import train_oracle as TO
    g = np.load(GOLD)
    R, net, handle, kw, target, msk = _train_setup()
    bucket = net.train_engine().bucket
    bucket.begin_step(1)
    rgb, disp, acc, extras = R.render(network_fn=handle, **kw)
    assert rgb.requires_grad and acc.requires_grad and not extras["raw"].requires_grad
    assert int(extras["pts_mask"].sum()) == int(g["n_active"])
    np.testing.assert_allclose(rgb[0].detach().cpu().numpy(), g["rgb_map"], atol=5e-4)
    np.testing.assert_allclose(acc[0].detach().cpu().numpy(), g["acc_map"], atol=2e-4)
    loss = torch.mean((rgb - target) ** 2) + torch.mean((msk.squeeze(2) - acc) ** 2)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    loss.backward()
    bucket.finish()
    named = dict(net.named_parameters())
    grads = {k: named[k].grad for k in TO.dense_keys() + TO.TRUNK_KEYS}
    assert all(v is not None for v in grads.values())
    check_grads_against_golden(grads, g, rtol_norm=2e-3, rtol_val=2e-2)
    # nothing else received a gradient (dead branches of the shipped configs stay dead)
    live = set(grads)
    assert not [k for k, p in named.items() if p.grad is not None and k not in live]


def test_training_step_in_the_reference_loop_without_trainstep(strict_fp32_convs):
    """The reference's loop body unchanged (run_nerf_batch.py:544-563): render -> loss -> loss.backward(), no
    TrainStep and no bucket calls.  The gradient bucket opens its step at the first render node and closes it from a
    callback on the autograd engine, so every live parameter has its .grad when backward() returns -- the same
    gradients as the reference's autograd -- and a second step after zero_grad() reproduces them."""
    from test_train_oracle import GOLD, check_grads_against_golden
    from oracle import train_oracle as TO
    g = np.load(GOLD)
    R, net, handle, kw, target, msk = _train_setup()
    named = dict(net.named_parameters())
    opt = torch.optim.Adam(list(net.parameters()), lr=0.0)
    for _ in range(2):
        opt.zero_grad()
        rgb, disp, acc, extras = R.render(network_fn=handle, **kw)
        loss = torch.mean((rgb - target) ** 2) + torch.mean((msk.squeeze(2) - acc) ** 2)
        loss.backward()
        grads = {k: named[k].grad for k in TO.dense_keys() + TO.TRUNK_KEYS}
        assert all(v is not None for v in grads.values())
        check_grads_against_golden(grads, g, rtol_norm=2e-3, rtol_val=2e-2)
        opt.step()


def test_train_step_optimises():
    """TrainStep on one fixed batch: Adam steps through the CUDA path reduce the loss, parameters and BN statistics move."""
    from mpsnerf_b200.train import TrainStep
    R, net, handle, kw, target, msk = _train_setup()
    # (the seeded weights have a density head of gain 300: a small step size keeps the descent monotone)
    opt = torch.optim.Adam([p for p in net.parameters()], lr=2e-5, betas=(0.9, 0.999))
    ts = TrainStep(handle, opt, acc_loss=True)
    w0 = net.pts_linears[3].weight.detach().clone()
    rm0 = net.encoder_2d.model.bn1.running_mean.clone()
    losses = [float(ts.step(R.render, target_rgb=target, bkgd_msk=msk, **kw)) for _ in range(8)]
    assert losses[-1] < losses[0] and all(np.isfinite(losses)), losses
    assert not torch.equal(w0, net.pts_linears[3].weight.detach())
    assert not torch.equal(rm0, net.encoder_2d.model.bn1.running_mean)
    # back to inference: eval() + no_grad goes through the production path again
    net.eval()
    with torch.no_grad():
        out = R.render(network_fn=handle, **dict(kw, perturb=False, perturb_u=None))
    assert not out[0].requires_grad and torch.isfinite(out[0]).all()


def test_smooth_step_against_reference_double_backward(strict_fp32_convs):
    """An interval step of the shipped configs (smooth_loss = 1, global_step % smooth_interval == 0): render() in
    training mode returns the two normal-smoothness terms in extras['other_loss'] and their gradients -- K1 / K3
    locate the sample points and their perturbed copies, smooth.py differentiates the active points twice -- against
    the UNMODIFIED reference's own double backward (tests/golden/smooth_grads.npz, oracle/make_golden_smooth.py).
    Tolerances: see tests/test_smooth_cpu.py (the terms are functions of normalised gradients)."""
    from test_smooth_cpu import GOLD, check_smooth_grads, smooth_keys
    from mpsnerf_b200.parser_config import config_parser
    from golden_cases import SMOOTH_CASE
    g = np.load(GOLD)
    R, net, handle, kw, target, msk = _train_setup(SMOOTH_CASE)
    R.configure(config_parser().parse_args(["--smooth_loss", "1"]))
    sp = dict(kw["sp_input"])
    sp["global_step"] = torch.zeros(1, dtype=torch.long)
    sp["smooth_interval"] = torch.full((1,), 4, dtype=torch.long)
    kw = dict(kw, sp_input=sp, smooth_delta=torch.from_numpy(g["delta"])[None].cuda())
    bucket = net.train_engine().bucket
    bucket.begin_step(1)
    rgb, disp, acc, extras = R.render(network_fn=handle, **kw)
    other = extras["other_loss"]
    assert other.shape == (1, 4) and other.requires_grad
    np.testing.assert_allclose(other.detach().cpu().numpy().reshape(4), g["other_loss"], rtol=5e-4, atol=1e-7)
    other[0][0].backward()
    bucket.absorb_autograd()
    bucket.finish()
    named = dict(net.named_parameters())
    grads = {k: named[k].grad for k in smooth_keys()}
    assert all(v is not None for v in grads.values())
    check_smooth_grads(grads, g, rtol_norm=2e-2, rtol_val=0.15, min_cos=0.995)
    # a step that is not on the interval carries no smooth terms
    sp["global_step"] = torch.ones(1, dtype=torch.long)
    bucket.begin_step(1)
    extras = R.render(network_fn=handle, **kw)[3]
    assert not extras["other_loss"].requires_grad and float(extras["other_loss"].abs().sum()) == 0.0


def test_train_step_with_smooth_terms_optimises():
    """TrainStep over four consecutive global steps with the shipped smooth_loss = 1: step 0 takes the smooth path,
    steps 1..3 the plain one; the loss stays finite and the parameters move."""
    from mpsnerf_b200.train import TrainStep
    from mpsnerf_b200.parser_config import config_parser
    R, net, handle, kw, target, msk = _train_setup()
    R.configure(config_parser().parse_args(["--smooth_loss", "1"]))
    opt = torch.optim.Adam([p for p in net.parameters()], lr=2e-5, betas=(0.9, 0.999))
    ts = TrainStep(handle, opt, acc_loss=True)
    w0 = net.alpha_linear.weight.detach().clone()
    losses = []
    for step in range(4):
        sp = dict(kw["sp_input"])
        sp["global_step"] = torch.full((1,), step, dtype=torch.long)
        sp["smooth_interval"] = torch.full((1,), 4, dtype=torch.long)
        losses.append(float(ts.step(R.render, target_rgb=target, bkgd_msk=msk, **dict(kw, sp_input=sp))))
    assert all(np.isfinite(losses)), losses
    assert not torch.equal(w0, net.alpha_linear.weight.detach())


def test_smooth_step_batch_of_two_equals_the_per_subject_steps(strict_fp32_convs):
    """B = 2 subjects (male / female tables, different poses and rays) in one smooth step: the terms are means over
    both subjects' sample points (the reference evaluates them on the gathered DataParallel output, ref :60-79), so
    other_loss and its gradients must equal the average of the two B = 1 steps -- which
    test_smooth_step_against_reference_double_backward pins to the reference."""
    from conftest import load_batch_case
    from test_smooth_cpu import smooth_keys
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    from mpsnerf_b200.lib import skinnning_batch as SB
    from mpsnerf_b200.parser_config import config_parser
    from golden_cases import SMOOTH_CASE
    scenes, _, sp, tp, models, g = load_batch_case("batch2")
    sd = synthetic.seeded_state_dict(scenes[0].seed, SMOOTH_CASE["alpha_gain"], SMOOTH_CASE["alpha_bias"])
    SB.set_default_smpl_models(models)
    torch.manual_seed(0)
    net = SB.SKinningBatch(human_sample=1, use_f2d=1, use_trans=1, smooth_loss=1, num_instances=25, mean_shape=0,
                           correction_field=0, skinning_field=0, data_set_type="THuman_B", append_rgb=1, with_viewdirs=0,
                           precision="fp32")
    net.load_state_dict(sd, strict=False)
    net = net.cuda().train()
    R.configure(config_parser().parse_args(["--smooth_loss", "1"]))
    handle = R.NetworkHandle(net)
    S = int(g["S"])
    parts = [synthetic.rays_tensor(sc, g["ray_ids"][b], device="cuda") for b, sc in enumerate(scenes)]
    rays, near, far = (torch.cat([p[k] for p in parts], 0) for k in range(3))
    n = rays.shape[2]
    gen = torch.Generator().manual_seed(11)
    delta = (0.01 * torch.randn(2, n * S, 3, generator=gen)).cuda()
    u = torch.rand(2, n, S, generator=gen).cuda()
    step = dict(global_step=torch.zeros(2, dtype=torch.long), smooth_interval=torch.full((2,), 4, dtype=torch.long))
    sp_d, tp_d = dict(_cuda_dict(sp), **step), _cuda_dict(tp)
    cut = lambda d, b: {k: (cut(v, b) if isinstance(v, dict) else v[b:b + 1]) for k, v in d.items()}
    keys = smooth_keys()
    named = dict(net.named_parameters())
    bucket = net.train_engine().bucket

    def run(sp_in, tp_in, sl):
        for p in net.parameters():
            p.grad = None
        bucket.begin_step(1)
        extras = R.render(rays=rays[sl], near=near[sl], far=far[sl], sp_input=sp_in, tp_input=tp_in, network_fn=handle,
                          N_samples=S, perturb=1.0, perturb_u=u[sl], smooth_delta=delta[sl], use_viewdirs=True)[3]
        other = extras["other_loss"]
        other[0][0].backward()
        bucket.absorb_autograd()
        bucket.finish()
        return other.detach().clone(), {k: named[k].grad.detach().clone() for k in keys}

    both, g_both = run(sp_d, tp_d, slice(0, 2))
    singles = [run(cut(sp_d, b), cut(tp_d, b), slice(b, b + 1)) for b in range(2)]
    assert float(singles[0][0][0, 1]) != float(singles[1][0][0, 1])          # the subjects really differ
    want = (singles[0][0] + singles[1][0]) / 2
    assert torch.allclose(both, want, rtol=1e-5, atol=1e-8), (both, want)
    # gradients: the same ill-conditioned sums as in tests/test_smooth_cpu.py (one fp32 ulp on x_c moves single entries
    # by ~4 %), and the order of the active points -- K1 compacts with atomics -- differs between the runs, so this is a
    # consistency check of the batching (a subject dropped, doubled or mis-averaged would be off by 50 - 100 %), with
    # generous bounds on the largest-entry error and the direction; numerically-zero gradients (the MLP biases) skipped
    scale = max(float((singles[0][1][k] + singles[1][1][k]).norm()) / 2 for k in keys)
    for k in keys:
        w = ((singles[0][1][k] + singles[1][1][k]) / 2).double().reshape(-1)
        got = g_both[k].double().reshape(-1)
        if float(w.norm()) < 1e-6 * scale:
            assert float(got.norm()) < 1e-5 * scale, k
            continue
        err = float((got - w).abs().max()) / max(float(w.abs().max()), 1e-30)
        cos = float(got @ w) / max(float(got.norm() * w.norm()), 1e-300)
        assert err <= 0.3 and cos >= 0.97, (k, err, cos)
