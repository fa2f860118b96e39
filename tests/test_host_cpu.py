"""CPU-side checks: config parsing, state-dict surface, weight packing, C-ABI symbols."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_config_files_parse_like_reference():
    from mpsnerf_b200.parser_config import config_parser
    a = config_parser().parse_args(["--config", os.path.join(ROOT, "configs", "canonical_transformer.txt")])
    assert a.model == "skinning_batch" and a.use_trans == 1 and a.append_rgb == 1 and a.N_samples == 128
    assert a.chunk == 12000 and a.use_viewdirs is True and a.white_bkgd is False and a.human_sample == 1
    assert a.correction_field == 0 and a.skinning_field == 0 and a.mean_shape == 0 and a.num_instance == 25
    # CLI wins over the file
    b = config_parser().parse_args(["--config", os.path.join(ROOT, "configs", "canonical_transformer.txt"), "--N_samples", "64"])
    assert b.N_samples == 64
    # h36m.txt uses the abbreviation i_test -> --i_testset (parser_config.py:101)
    h = config_parser().parse_args(["--config", os.path.join(ROOT, "configs", "h36m.txt")])
    assert h.i_testset == 6000 and h.data_set_type == "H36M_P" and h.num_instance == 6


def _make_net():
    from mpsnerf_b200 import synthetic
    from mpsnerf_b200.lib import skinnning_batch as SB
    from mpsnerf_b200.parser_config import config_parser
    from mpsnerf_b200.model_selection import return_model
    SB.set_default_smpl_models(synthetic.make_smpl("n", 0))
    args = config_parser().parse_args(["--config", os.path.join(ROOT, "configs", "canonical_transformer.txt")])
    return return_model(args)


def test_state_dict_surface():
    net = _make_net()
    sd = net.state_dict()
    shapes = {
        "pts_linears.0.weight": (256, 194), "pts_linears.5.weight": (256, 450), "pts_linears.7.bias": (256,),
        "alpha_linear.weight": (1, 256), "feature_linear.weight": (256, 256), "views_linear.weight": (128, 411),
        "rgb_linear.weight": (3, 128), "transformer.layers.0.0.fn.norm.weight": (155,),
        "transformer.layers.1.0.fn.fn.to_qkv.weight": (768, 155),
        "transformer.layers.0.0.fn.fn.to_out.0.weight": (155, 256), "transformer.layers.0.0.fn.fn.to_out.0.bias": (155,),
        "transformer.layers.1.1.fn.fn.net.0.weight": (128, 155), "transformer.layers.1.1.fn.fn.net.3.weight": (155, 128),
        "latent_codes.weight": (25, 128), "pos_enc._freqs": (1, 12, 1), "view_enc._freqs": (1, 8, 1),
        "encoder_2d.model.conv1.weight": (64, 3, 7, 7), "encoder_2d.model.layer4.2.bn2.running_var": (512,),
        "encoder_3d.conv0.0.weight": (16, 3, 3, 3, 3), "encoder_3d.down3.1.running_mean": (128,),
        "forward_deform.output_linear.weight": (3, 256), "backward_deform.pts_time_linears.3.weight": (256, 256),
    }
    for k, s in shapes.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == s, (k, tuple(sd[k].shape))
    assert "transformer.layers.0.0.fn.fn.to_qkv.bias" not in sd
    from mpsnerf_b200 import synthetic
    missing = net.load_state_dict(synthetic.seeded_state_dict(0), strict=False)
    assert not missing.unexpected_keys


def test_unsupported_configs_fail_loudly():
    net = _make_net()
    net.mean_shape = 1
    with pytest.raises(NotImplementedError):
        net._check_supported()
    from mpsnerf_b200.model_selection import return_model
    from mpsnerf_b200.parser_config import config_parser
    with pytest.raises(NotImplementedError):
        return_model(config_parser().parse_args(["--model", "direct_deform"]))


def test_no_cpu_path():
    from mpsnerf_b200 import synthetic
    net = _make_net().eval()
    sc = synthetic.make_scene("thuman", seed=0, H=64, W=64)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        net(sc.sp_input, sc.tp_input, torch.zeros(1, 8, 3), None)


def test_pack_kmajor_sw128_matches_address_formula():
    from mpsnerf_b200.engine import pack_kmajor_sw128
    g = torch.Generator().manual_seed(1)
    for N, K in ((256, 194), (768, 155), (16, 64), (160, 256)):
        w = torch.randn(N, K, generator=g)
        blob = pack_kmajor_sw128(w).numpy().view(np.uint16)
        n_pad, k_pad = (N + 15) // 16 * 16, (K + 63) // 64 * 64
        wp = torch.zeros(n_pad, k_pad, dtype=torch.bfloat16)
        wp[:N, :K] = w.bfloat16()
        ref = wp.view(torch.int16).numpy().view(np.uint16)
        rng = np.random.RandomState(0)
        for _ in range(500):
            r, k = rng.randint(n_pad), rng.randint(k_pad)
            chunk, j16, e = k // 64, (k % 64) // 8, k % 8
            byte = chunk * n_pad * 128 + (r // 8) * 1024 + (r % 8) * 128 + ((j16 ^ (r % 8)) << 4) + 2 * e
            assert blob[byte // 2] == ref[r, k]


def test_cabi_exports_every_declared_symbol():
    from mpsnerf_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "mpsnerf.h")).read()
    declared = set(re.findall(r"\b(mpsnerf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mpsnerf_abi_version() == 1
    assert ctypes.sizeof(_lib.Frame) == 4 * (3 + 9 + 9 + 3 + 4 * 288 + 72 + 24 + 72 + 8)


def test_mesh_grid_points_device_form_matches_numpy_meshgrid():
    """extract_thuman_mesh.grid_points: the broadcast form used on the device equals the reference's
    np.stack(np.meshgrid(...)).astype(float32) bit for bit (checked here on the CPU device)."""
    import numpy as np
    import torch
    from mpsnerf_b200 import extract_thuman_mesh as X
    for can in (False, True):
        ref, S0, Z0, R0 = X.grid_points(can, 16)
        got, S1, Z1, R1 = X.grid_points(can, 16, device=torch.device("cpu"))
        assert ref.dtype == np.float32 and tuple(got.shape) == ref.shape
        assert np.array_equal(got.numpy(), ref) and np.array_equal(R0, R1) and np.array_equal(S0, S1)


def test_hot_input_selection_for_host_dicts():
    """render() uploads only the dict entries the path reads (run_nerf_batch.HOT_KEYS_*): the selection keeps
    nested params, drops e.g. the target view's images, and hot_input_bytes counts exactly the kept tensors."""
    import torch
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    sc = synthetic.make_scene("thuman", seed=0, H=64, W=64)
    sp = R._upload_hot(sc.sp_input, R.HOT_KEYS_SP, torch.device("cpu"))
    tp = R._upload_hot(sc.tp_input, R.HOT_KEYS_TP, torch.device("cpu"))
    assert set(sp) == set(R.HOT_KEYS_SP) and set(tp) == set(R.HOT_KEYS_TP)
    assert "img_all" not in tp and set(tp["params"]) == set(sc.tp_input["params"])
    assert torch.equal(sp["img_all"], sc.sp_input["img_all"]) and torch.equal(tp["vertices"], sc.tp_input["vertices"])
    count = lambda d: sum(v.numel() * v.element_size() if torch.is_tensor(v) else count(v) if isinstance(v, dict) else 0 for v in d.values())
    assert R.hot_input_bytes(sc.sp_input, sc.tp_input) == count(sp) + count(tp)
    assert R.hot_input_bytes(sc.sp_input, sc.tp_input) < count(sc.sp_input) + count(sc.tp_input)
    # every key the engine / frame cache reads is in the selection
    import inspect
    from mpsnerf_b200 import engine
    src = inspect.getsource(engine.RenderEngine._prepare_frame)
    for key in ("img_all", "R_all", "T_all", "K_all", "t_vertices", "params"):
        assert f'sp["{key}"]' in src and key in R.HOT_KEYS_SP
    assert 'tp["vertices"]' in src and "vertices" in R.HOT_KEYS_TP


def test_assemble_latent_rows_matches_the_row_split():
    """engine.assemble_latent_rows: rank-major, padded bands of latent rows (what the all-gather of the sharded encoder
    trunk returns) -> the full (V, Hf, Wf, C) latent, for even and ragged splits."""
    import torch
    from mpsnerf_b200.engine import assemble_latent_rows
    from mpsnerf_b200.parallel import ray_block
    g = torch.Generator().manual_seed(0)
    for Hf, world in ((250, 8), (128, 8), (7, 3), (5, 8), (16, 2)):
        full = torch.randn(3, Hf, 4, 6, generator=g)
        mr = (Hf + world - 1) // world
        bands = torch.zeros(world, 3, mr, 4, 6)
        for r in range(world):
            a, b = ray_block(Hf, r, world)
            bands[r, :, :b - a] = full[:, a:b]
        assert torch.equal(assemble_latent_rows(bands, Hf), full)


def test_too_many_views_raise_before_anything_is_enqueued():
    """the tensor-core path takes 2..4 input views (fp32: 2..8); a larger V must fail up front with a clear
    message, not after K1 / K3 / K4 have been queued."""
    import pytest
    import torch
    from mpsnerf_b200 import synthetic
    from mpsnerf_b200.engine import RenderEngine
    from mpsnerf_b200.lib import skinnning_batch as SB
    from mpsnerf_b200.run_nerf_batch import _select
    scene = synthetic.make_scene("thuman", seed=0, H=32, W=32, n_views=5)
    SB.set_default_smpl_models(scene.smpl)
    net = SB.SKinningBatch(human_sample=1, use_f2d=1, use_trans=1, smooth_loss=1, num_instances=2, mean_shape=0,
                           correction_field=0, skinning_field=0, data_set_type="THuman_B", append_rgb=1, with_viewdirs=0)
    sp, tp = _select(scene.sp_input, 0), _select(scene.tp_input, 0)
    with pytest.raises(ValueError, match="input views"):
        RenderEngine(net, precision="bf16").prepare_frame(sp, tp, net._smpl_for(sp["gender"]))


def test_build_records_the_source_hash():
    """a stale .so must not load silently -- the library carries the hash of the sources it was built from."""
    from mpsnerf_b200 import build as B
    assert B.is_current() and len(B.source_hash()) == 64
    with open(B.HASH_FILE) as fh:
        assert fh.read().strip() == B.source_hash()


def test_bench_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract keys, the same `config` object
    our arm prints, and `cpu_baseline.kind` = "reference" when a reference tree is present, "port" otherwise."""
    import argparse
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from oracle import run_reference as RR
    monkeypatch.setenv("MPSNERF_CPU_ARM_BUDGET_S", "2")
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    cwd = os.getcwd()
    try:
        bench.run_reference(argparse.Namespace(gpus=1, steps=1, warmup=0, workload=None, mode="auto"))
    finally:
        os.chdir(cwd)
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0
    assert d["config"] == bench.workload_config("thuman", 1, False)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == ("reference" if RR.find_reference() else "port")
    # N > 1: both arms name the strong-scaling workload (BASELINE configs[2])
    monkeypatch.setenv("WORLD_SIZE", "8")
    a = argparse.Namespace(gpus=8, steps=1, warmup=0, workload=None, mode="auto")
    assert bench.resolve_mode(a) == (8, "h36m", True)


def test_header_is_plain_c_and_a_c_host_links_against_the_library(tmp_path):
    """include/mpsnerf.h is the drop-in boundary: it must compile as C99 (no C++ in the signatures), and a C host that
    binds the entry points the way INTEGRATION.md describes must link against the in-tree .so and reach the functions
    that need no GPU (ABI version, argument validation with the thread-local error string, workspace sizes)."""
    import subprocess
    from mpsnerf_b200 import build
    lib = build.build()
    libdir, root = os.path.dirname(lib), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "mpsnerf.h"
int main(void) {
  if (mpsnerf_abi_version() != MPSNERF_ABI_VERSION) return 1;
  /* a call with a missing pointer is refused on the host, before any CUDA work, and says why */
  if (mpsnerf_allreduce_mean(NULL, NULL, 4, NULL) != MPSNERF_EINVAL) return 2;
  if (strstr(mpsnerf_last_error(), "mpsnerf_allreduce_mean") == NULL) return 3;
  if (mpsnerf_allreduce_mean(NULL, NULL, 0, NULL) != MPSNERF_OK) return 4;      /* empty bucket: nothing to do */
  if (mpsnerf_grid_bytes(6890) == 0) return 5;
  if (mpsnerf_dense_train_workspace(1000, 3) == 0 || mpsnerf_render_rays_workspace(1024, 64, 3, 65536) == 0) return 6;
  printf("abi %d frame %zu bytes\n", mpsnerf_abi_version(), sizeof(mpsnerf_frame));
  return 0;
}
''')
    exe = tmp_path / "host"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-lmpsnerf_b200", "-Wl,-rpath," + libdir])
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([str(exe)], env=env, capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    from mpsnerf_b200 import _lib
    import ctypes
    assert out.stdout.split() == ["abi", "1", "frame", str(ctypes.sizeof(_lib.Frame)), "bytes"]


def test_create_nerf_resumes_from_the_last_checkpoint(tmp_path):
    """create_nerf (ref run_nerf_batch.py:301-366): same five return values and render_kwargs keys; resumes from the
    lexicographically last ``*.tar`` under basedir/expname in the reference's checkpoint format
    ({'global_step', 'network_fn_state_dict'}, :606-617) unless --no_reload; --ft_path names one explicitly."""
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    from mpsnerf_b200.lib import skinnning_batch as SB
    from mpsnerf_b200.parser_config import config_parser
    SB.set_default_smpl_models(synthetic.make_smpl("n", 0))
    exp = tmp_path / "exp"
    exp.mkdir()
    cfg = ["--config", os.path.join(ROOT, "configs", "canonical_transformer.txt"), "--basedir", str(tmp_path), "--expname", "exp"]
    cpu = torch.device("cpu")
    tr, te, start, grad_vars, opt = R.create_nerf(config_parser().parse_args(cfg), device=cpu)
    assert start == 0
    assert set(tr) == {"network_query_fn", "perturb", "N_samples", "network_fn", "use_viewdirs", "N_importance"}
    assert set(te) == set(tr) and te["perturb"] is False and tr["perturb"] == 1.0 and tr["network_fn"] is te["network_fn"]
    net = R._net_of(tr["network_fn"])
    assert len(grad_vars) == len(list(net.parameters())) and isinstance(opt, torch.optim.Adam)
    assert opt.param_groups[0]["lr"] == config_parser().parse_args(cfg).lrate
    # two checkpoints: the later one wins; the earlier one can be named with --ft_path
    sd_a = {k: v.clone() for k, v in net.state_dict().items()}
    sd_b = {k: (v + 1.0 if v.is_floating_point() else v.clone()) for k, v in sd_a.items()}
    torch.save({"global_step": 1000, "network_fn_state_dict": sd_a}, str(exp / "001000.tar"))
    torch.save({"global_step": 2000, "network_fn_state_dict": sd_b}, str(exp / "002000.tar"))
    key = "pts_linears.3.weight"
    tr2, _, start2, _, _ = R.create_nerf(config_parser().parse_args(cfg), device=cpu)
    assert start2 == 2000 and torch.equal(R._net_of(tr2["network_fn"]).state_dict()[key], sd_b[key])
    tr3, _, start3, _, _ = R.create_nerf(config_parser().parse_args(cfg + ["--ft_path", "001000.tar"]), device=cpu)
    assert start3 == 1000 and torch.equal(R._net_of(tr3["network_fn"]).state_dict()[key], sd_a[key])
    _, _, start4, _, _ = R.create_nerf(config_parser().parse_args(cfg + ["--no_reload"]), device=cpu)
    assert start4 == 0
