"""CPU suite: the oracle restatement against the committed golden vectors (generated from the unmodified reference by
oracle/make_golden.py) and, when /root/reference is present, against the reference modules themselves."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, restormer as O, weights as W
from oracle import nafnet as ON
from oracle import vit as OV
from oracle.make_golden import VIT_CASES
from oracle.make_golden import (GUIDED_CASES, NAF_GUIDED_CASES, NAFNET_CASES, RESTORMER_CASES, denoise_inputs,
                                guided_inputs)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return json.loads(str(z["meta"])), torch.from_numpy(z["out"])


def _shapes(module_ctor, cfg):
    return {k: v.shape for k, v in module_ctor(**cfg).state_dict().items()}


@pytest.mark.parametrize("name", list(RESTORMER_CASES))
def test_restormer_oracle_matches_golden(name):
    from textualdegremoval_b200.archs.restormer_b200_arch import Restormer
    meta, ref = _load(name)
    assert meta == json.loads(json.dumps(RESTORMER_CASES[name]))
    sd = W.seeded_state_dict(_shapes(Restormer, meta["cfg"]), meta["seed"])
    x = W.seeded_image("x", meta["shape"], meta["seed"])
    with torch.no_grad():
        y = O.restormer_forward(sd, x, meta["cfg"]["heads"])
    assert (y - ref).abs().max().item() < 2e-5          # fp32 vs fp32, different op order only


@pytest.mark.parametrize("name", list(GUIDED_CASES))
def test_guided_oracle_matches_golden(name):
    from textualdegremoval_b200.archs.restormer_b200_arch import RestormerRefFusion
    meta, ref = _load(name)
    sd = W.seeded_state_dict(_shapes(RestormerRefFusion, meta["cfg"]), meta["seed"])
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        y = O.restormer_ref_fusion_forward(sd, lq, rf, meta["cfg"]["heads"])
    assert (y - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", list(NAFNET_CASES))
def test_nafnet_oracle_matches_golden(name):
    from textualdegremoval_b200.archs.nafnet_b200_arch import NAFNet
    meta, ref = _load(name)
    sd = W.seeded_state_dict(_shapes(NAFNet, meta["cfg"]), meta["seed"])
    lq, _ = denoise_inputs(meta)
    with torch.no_grad():
        y = ON.nafnet_forward(sd, lq)
    assert (y - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", list(NAF_GUIDED_CASES))
def test_guided_nafnet_oracle_matches_golden(name):
    from textualdegremoval_b200.archs.nafnet_b200_arch import NAFNetRefFusion
    meta, ref = _load(name)
    sd = W.seeded_state_dict(_shapes(NAFNetRefFusion, meta["cfg"]), meta["seed"])
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        y = ON.nafnet_ref_fusion_forward(sd, lq, rf)
    assert (y - ref).abs().max().item() < 2e-5


def test_config1_nafnet_tiny_plumbing():
    """BASELINE.json configs[0] (CPU plumbing): registry -> NAFNet-tiny, structural checksum 1,136,625 params
    (SURVEY 8c iii), oracle forward on the sigma=15 gray tile, L1 loss against the clean tile is finite."""
    from textualdegremoval_b200 import define_network
    meta = NAFNET_CASES["nafnet_tiny_gray64"]
    net = define_network(dict(type="NAFNet", **meta["cfg"]))
    assert sum(p.numel() for p in net.parameters()) == 1136625
    sd = W.load_seeded(net, meta["seed"])
    lq, gt = denoise_inputs(meta)
    with torch.no_grad():
        loss = (ON.nafnet_forward(sd, lq) - gt).abs().mean().item()
    assert 0 < loss < 10


def test_vit_oracles_match_golden():
    """DINOv2 ViT (reference models/dino), CLIP tower (installed transformers; parity with 4.31.0 unpinned) and the
    mapper MLPs (classes extracted from the reference script) against their fixtures."""
    from textualdegremoval_b200.archs import vit_b200 as VB
    meta, ref = _load("dino_vit_tiny")
    sd = W.seeded_state_dict({k: v.shape for k, v in VB.DinoVisionTransformer(**meta["cfg"]).state_dict().items()}, meta["seed"])
    with torch.no_grad():
        y = OV.dino_vit_forward(sd, W.seeded_image("x", meta["shape"], meta["seed"]), heads=meta["cfg"]["num_heads"])
    assert (y - ref).abs().max().item() < 2e-5
    meta, ref = _load("clip_vit_tiny")
    c = meta["cfg"]
    sd = W.seeded_state_dict({k: v.shape for k, v in VB.CLIPVisionTower(**c).state_dict().items()}, meta["seed"])
    with torch.no_grad():
        y = OV.clip_vision_forward(sd, W.seeded_image("x", meta["shape"], meta["seed"]), c["num_attention_heads"], c["patch_size"])
    assert (y - ref).abs().max().item() < 2e-5
    z = np.load(os.path.join(GOLD, "mappers_tiny.npz"))
    meta = json.loads(str(z["meta"]))
    c = meta["cfg"]
    m, cm = VB.Mapper(c["input_dim"], c["mid_dim"], c["num_words"]), VB.CleanMapper(c["mid_dim"], c["mid_dim"], c["num_words"])
    sd1 = W.seeded_state_dict({k: v.shape for k, v in m.state_dict().items()}, meta["seed"])
    sd2 = W.seeded_state_dict({k: v.shape for k, v in cm.state_dict().items()}, meta["seed"] + 1)
    emb = W.seeded_image("emb", meta["shape"], meta["seed"]) * 2 - 1
    with torch.no_grad():
        w1 = OV.mapper_forward(sd1, emb, c["num_words"])
        w2 = OV.clean_mapper_forward(sd2, w1, c["num_words"])
    assert (w1 - torch.from_numpy(z["out"])).abs().max().item() < 2e-5
    assert (w2 - torch.from_numpy(z["out2"])).abs().max().item() < 2e-5


def test_vit_structural_checksums():
    """Parameter counts of the full-size encoders / mappers (SURVEY 8c iii)."""
    from textualdegremoval_b200.archs import vit_b200 as VB
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(VB.vit_base(img_size=518, patch_size=14, init_values=1.0, ffn_layer="mlp", block_chunks=0)) == 86580480
    with torch.device("meta"):
        assert n(VB.CLIPVisionTower()) == 630766080 + 0
        assert n(VB.Mapper(1280, 1024, 20)) == 249538560
        assert n(VB.CleanMapper(1024, 1024, 20)) == 118215680


def test_identity_at_zero_alpha():
    """Known-answer anchor (SURVEY 8c i): with alpha = 0 the fusion blocks are identities on [:, :C], so the guided net
    equals the unguided Restormer built from the same non-masa weights."""
    from textualdegremoval_b200.archs.restormer_b200_arch import RestormerRefFusion
    meta = GUIDED_CASES["guided_restormer_128"]
    sd = W.seeded_state_dict(_shapes(RestormerRefFusion, meta["cfg"]), meta["seed"])
    for k in sd:
        if k.endswith(".alpha"):
            sd[k] = torch.zeros_like(sd[k])
    lq, rf = guided_inputs(meta)
    with torch.no_grad():
        y = O.restormer_ref_fusion_forward(sd, lq, rf)
        y0 = O.restormer_forward({k: v for k, v in sd.items() if "masa" not in k}, lq)
    assert (y - y0).abs().max().item() < 1e-5


def test_transfer_overlap_counts():
    """transfer (:698-715): a constant window must come back as att (bilinear of a constant) exactly -- i.e. the
    overlap-add is divided by the right 4..9 patch count everywhere."""
    m, c, k, d, s = 2, 3, 8, 13, 4
    win = torch.ones(m, c, (d + 2) * s, (d + 2) * s)
    index = torch.randint(0, d * d, (m, k, k))
    att = torch.full((m, k, k), 0.5)
    out = O.transfer(win, index, att, s, d)
    assert torch.allclose(out, torch.full_like(out, 0.5), atol=1e-6)


def test_window_origin_clamps():
    idx = torch.tensor([[0, 15, 16 * 15 + 15, 16 * 8 + 8]])
    y1, x1 = O.window_origin(idx, 16, 16, 13, 13)
    assert y1.tolist() == [[0, 0, 1, 1]] and x1.tolist() == [[0, 1, 1, 1]]
    assert int((y1 + 14).max()) <= 15 and int((x1 + 14).max()) <= 15


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference():
    cfg = dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8], nf=16,
               ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1], LayerNorm_type="BiasFree")
    net = ref_loader.restormer_ref_fusion(**cfg)
    sd = W.load_seeded(net, 5)
    lq = W.seeded_image("a", (1, 3, 128, 192), 5)
    rf = W.seeded_image("b", (1, 3, 192, 128), 5)
    with torch.no_grad():
        assert (net(lq, rf) - O.restormer_ref_fusion_forward(sd, lq, rf)).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", ["nafnet_rgb_ragged", "guided_nafnet_256"])
def test_nafnet_oracle_autograd_matches_reference_gradients(name):
    """Same pin for the NAFNet family: the reference differentiates LayerNorm2d with its hand-written
    LayerNormFunction.backward (nafnet_arch_utils.py:277-289); the oracle's plain autograd must agree with it."""
    from oracle.make_golden import GRAD_CASES, grad_probe
    from textualdegremoval_b200.archs.nafnet_b200_arch import NAFNet, NAFNetRefFusion
    z = np.load(os.path.join(GOLD, name + "_grad.npz"))
    meta = json.loads(str(z["meta"]))
    guided = GRAD_CASES[name] == "guided_nafnet"
    sd = W.seeded_state_dict(_shapes(NAFNetRefFusion if guided else NAFNet, meta["cfg"]), meta["seed"])
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if guided:
        lq, rf = guided_inputs(meta)
        gt = W.seeded_image("gt", meta["lq"], meta["seed"])
        y = ON.nafnet_ref_fusion_forward(sdg, lq, rf)
    else:
        x, gt = denoise_inputs(meta)
        y = ON.nafnet_forward(sdg, x)
    loss = (y - gt).abs().mean()
    loss.backward()
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-6
    total = float(np.sqrt((z["norms"] ** 2).sum()))
    assert len(z["names"]) == len(sdg)
    for n, norm, probe in zip(z["names"].tolist(), z["norms"].tolist(), z["probes"].tolist()):
        g = sdg[n].grad if sdg[n].grad is not None else torch.zeros_like(sdg[n])
        a, b = grad_probe(n, g)
        tol = 2e-4 * max(norm, 1e-3 * total)
        assert abs(a - norm) < tol, (n, a, norm)
        assert abs(b - probe) < tol * max(1.0, float(g.numel()) ** 0.5), (n, b, probe)


@pytest.mark.parametrize("name", ["restormer_withbias", "guided_restormer_128"])
def test_oracle_autograd_matches_reference_gradients(name):
    """The oracle's autograd (the checker of the CUDA backward schedule) against per-parameter gradient fingerprints of
    the unmodified reference modules (oracle/make_golden.py main_grads)."""
    from oracle.make_golden import GRAD_CASES, grad_probe
    from textualdegremoval_b200.archs.restormer_b200_arch import Restormer, RestormerRefFusion
    z = np.load(os.path.join(GOLD, name + "_grad.npz"))
    meta = json.loads(str(z["meta"]))
    guided = GRAD_CASES[name] == "guided"
    sd = W.seeded_state_dict(_shapes(RestormerRefFusion if guided else Restormer, meta["cfg"]), meta["seed"])
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if guided:
        lq, rf = guided_inputs(meta)
        gt = W.seeded_image("gt", meta["lq"], meta["seed"])
        y = O.restormer_ref_fusion_forward(sdg, lq, rf, meta["cfg"]["heads"])
    else:
        x = W.seeded_image("x", meta["shape"], meta["seed"])
        gt = W.seeded_image("gt", meta["shape"], meta["seed"])
        y = O.restormer_forward(sdg, x, meta["cfg"]["heads"])
    loss = (y - gt).abs().mean()
    loss.backward()
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-6
    total = float(np.sqrt((z["norms"] ** 2).sum()))
    for n, norm, probe in zip(z["names"].tolist(), z["norms"].tolist(), z["probes"].tolist()):
        g = sdg[n].grad if sdg[n].grad is not None else torch.zeros_like(sdg[n])
        a, b = grad_probe(n, g)
        tol = 2e-4 * max(norm, 1e-3 * total)
        assert abs(a - norm) < tol, (n, a, norm)
        assert abs(b - probe) < tol * max(1.0, float(g.numel()) ** 0.5), (n, b, probe)


def test_input_pipeline_oracle_matches_reference_fixture():
    """oracle/input_pipeline.py vs tensors produced by the unmodified reference chain (padding, paired_random_crop,
    random_augmentation, img2tensor, normalize; oracle/make_golden_input.py): bit-exact, all 8 augmentation modes,
    frames smaller than the patch (single and multiple reflections)."""
    import numpy as np
    from oracle import input_pipeline as IP
    z = np.load(os.path.join(GOLD, "input_pipeline.npz"))
    modes, padded = set(), 0
    for k in range(int(z["n"])):
        top, left, mode, size, norm = [int(v) for v in z[f"s{k}_dec"]]
        modes.add(mode)
        padded += int(z[f"s{k}_gt_frame"].shape[0] < size or z[f"s{k}_gt_frame"].shape[1] < size)
        mean, std = (z["mean"], z["std"]) if norm else (None, None)
        for which in ("gt", "lq"):
            got = IP.prepare_patch(z[f"s{k}_{which}_frame"], top, left, mode, size, mean=mean, std=std)
            ref = z[f"s{k}_{which}"]
            assert got.shape == ref.shape and got.dtype == ref.dtype
            assert np.array_equal(got, ref), f"sample {k} {which}: max diff {np.abs(got - ref).max()}"
    assert modes == set(range(8)) and padded >= 8


def test_psnr_oracle_matches_reference_fixture():
    """oracle/metrics.py vs the doubles returned by the unmodified tensor2img + calculate_psnr (oracle/make_golden_metrics.py):
    identical float64, incl. crop_border, clamping, round-half-to-even ties, the max_value = 1 branch and mse == 0."""
    from oracle import metrics as M
    from oracle.make_golden_metrics import CASES, make_pair
    ref = np.load(os.path.join(GOLD, "psnr.npz"))["psnr"]
    for i, case in enumerate(CASES):
        res, gt = make_pair(case, 500 + i)
        got = M.psnr(res.numpy(), gt.numpy(), case["crop"])
        assert got == ref[i], (case, got, ref[i])
        sse, mx, n = M.psnr_sums(res.numpy(), gt.numpy(), case["crop"])
        if sse:
            assert float(20. * np.log10((1. if mx <= 1 else 255.) / np.sqrt(np.float64(sse) / np.float64(n)))) == ref[i]


def test_fullsize_fixtures_pin_the_oracle_at_baseline_sizes():
    """tests/golden/full_*.npz hold outputs of the UNMODIFIED reference modules at the BASELINE.json sizes; the generator
    (oracle/make_golden_fullsize.py) asserted the oracle reproduces each.  Re-check the recorded agreement for all five and
    re-run the two cheapest (Restormer 256x256, DINOv2 ViT-B/14 518x518) against the stored reference outputs."""
    import json
    import os

    import numpy as np

    from oracle import restormer as O, vit as OV, weights as W
    from oracle.make_golden_fullsize import FULL_CASES, fullsize_inputs
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in FULL_CASES:
        z = np.load(os.path.join(gold, name + ".npz"))
        meta = json.loads(str(z["meta"]))
        assert meta["oracle_vs_reference_max"] < 2e-4, (name, meta["oracle_vs_reference_max"])
        assert z["out"].dtype == np.float32 and np.isfinite(z["out"]).all()
    with torch.no_grad():
        for name in ("full_restormer_256", "full_dino_vitb_518"):
            z = np.load(os.path.join(gold, name + ".npz"))
            meta = json.loads(str(z["meta"]))
            x, _, _ = fullsize_inputs(meta)
            if meta["kind"] == "restormer":
                from textualdegremoval_b200.archs import define_network
                shapes = {k: v.shape for k, v in define_network(dict(type="Restormer", **meta["cfg"])).state_dict().items()}
                y = O.restormer_forward(W.seeded_state_dict(shapes, meta["seed"]), x, meta["cfg"]["heads"])
            else:
                from textualdegremoval_b200.archs import vit_b200 as VB
                shapes = {k: v.shape for k, v in VB.vit_base(**meta["cfg"]).state_dict().items()}
                y = OV.dino_vit_forward(W.seeded_state_dict(shapes, meta["seed"]), x)
            assert (y - torch.from_numpy(z["out"])).abs().max().item() < 2e-4, name
