"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every fixture holds: the ctor kwargs (json), the seed (weights and inputs are regenerated from
``oracle.weights``), and the reference module's fp32 outputs.  The reference modules are imported
through ``oracle.ref_loader`` (stubbed package inits, B1 index shim -- see its docstring).
"""
import json
import os

import numpy as np
import torch

from . import ref_loader as R
from . import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

RESTORMER_CASES = {
    "restormer_withbias": dict(cfg=dict(dim=16, num_blocks=[1, 2, 1, 1], num_refinement_blocks=1,
                                        heads=[1, 2, 4, 8], LayerNorm_type="WithBias", bias=False),
                               seed=11, shape=(2, 3, 64, 64)),
    "restormer_biasfree": dict(cfg=dict(dim=24, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                        heads=[1, 2, 4, 8], LayerNorm_type="BiasFree", bias=False),
                               seed=12, shape=(1, 3, 64, 96)),
    "restormer_gray_bias": dict(cfg=dict(inp_channels=1, out_channels=1, dim=16, num_blocks=[1, 1, 1, 1],
                                         num_refinement_blocks=1, heads=[1, 2, 4, 8], LayerNorm_type="BiasFree",
                                         bias=True),
                                seed=13, shape=(1, 1, 64, 64)),
}
GUIDED_CASES = {
    "guided_restormer_128": dict(cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                          heads=[1, 2, 4, 8], nf=16, ext_n_blocks=[1, 1, 1, 1],
                                          reffusion_n_blocks=[1, 2, 1, 1], LayerNorm_type="WithBias"),
                                 seed=21, lq=(2, 3, 128, 128), ref=(2, 3, 128, 128)),
    "guided_restormer_ragged": dict(cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
                                             heads=[1, 2, 4, 8], nf=16, ext_n_blocks=[2, 1, 1, 1],
                                             reffusion_n_blocks=[1, 1, 1, 1], LayerNorm_type="WithBias"),
                                    seed=22, lq=(1, 3, 120, 130), ref=(1, 3, 128, 192)),
    # dual-pixel defocus deblurring (options/train_restoration/004_0*.yml): 6 input channels, skip_conv, no image residual
    "guided_restormer_dual": dict(cfg=dict(inp_channels=6, out_channels=3, dim=16, num_blocks=[1, 1, 1, 1],
                                           num_refinement_blocks=1, heads=[1, 2, 4, 8], nf=16, ext_n_blocks=[1, 1, 1, 1],
                                           reffusion_n_blocks=[1, 1, 1, 1], LayerNorm_type="WithBias",
                                           dual_pixel_task=True),
                                  seed=23, lq=(1, 6, 128, 128), ref=(1, 6, 128, 128)),
}


NAFNET_CASES = {
    # BASELINE.json configs[0]: NAFNet-tiny Gaussian gray denoise sigma=15, one 64x64 tile
    "nafnet_tiny_gray64": dict(cfg=dict(img_channel=1, width=16, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 1],
                                        dec_blk_nums=[1, 1, 1, 1]), seed=31, shape=(1, 1, 64, 64), sigma=15),
    "nafnet_rgb_ragged": dict(cfg=dict(img_channel=3, width=16, middle_blk_num=2, enc_blk_nums=[1, 2],
                                       dec_blk_nums=[1, 1]), seed=32, shape=(2, 3, 50, 70), sigma=25),
}
NAF_GUIDED_CASES = {
    "guided_nafnet_256": dict(cfg=dict(img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[1, 1, 1, 2],
                                       dec_blk_nums=[1, 1, 1, 1], nf=16, ext_n_blocks=[1, 1, 1, 1],
                                       reffusion_n_blocks=[1, 1, 1, 1, 1]),
                              seed=41, lq=(1, 3, 256, 256), ref=(1, 3, 256, 256)),
}


# PromptIRRefFusion: the prompt blocks hard-code lin_dim 96 / 192 / 384, i.e. dim = 48 is the only width the reference
# class can run with; decoder=True is the only mode whose forward runs at all
PROMPTIR_CASES = {
    "guided_promptir_128": dict(cfg=dict(dim=48, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8],
                                         nf=48, ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1],
                                         LayerNorm_type="WithBias", decoder=True),
                                seed=51, lq=(1, 3, 128, 128), ref=(1, 3, 128, 128)),
}


DRS_CASES = {
    # option 007 (Rain200L / SPA variant, no MEFC) and options 008-010 (MEFC encoder_level0 + refinement), small widths
    "guided_drsformer_spa_128": dict(spa=True, cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], heads=[1, 2, 4, 8], nf=16,
                                                        ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1],
                                                        LayerNorm_type="WithBias"),
                                     seed=71, lq=(1, 3, 128, 128), ref=(1, 3, 128, 128)),
    # conv biases, BiasFree LayerNorm and a ragged input size (zero-padded to a multiple of 64 and cropped back)
    "guided_drsformer_spa_bias_ragged": dict(spa=True, cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], heads=[1, 2, 4, 8], nf=16,
                                                                ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1],
                                                                LayerNorm_type="BiasFree", bias=True),
                                             seed=73, lq=(1, 3, 120, 136), ref=(1, 3, 128, 192)),
    "guided_drsformer_128": dict(spa=False, cfg=dict(dim=16, num_blocks=[1, 1, 1, 1], heads=[1, 2, 4, 8], nf=16,
                                                     ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1],
                                                     LayerNorm_type="WithBias"),
                                 seed=72, lq=(1, 3, 128, 128), ref=(1, 3, 128, 128)),
}


def denoise_inputs(case):
    """The reference's deterministic test-noise recipe (data/restoration_dataset.py:479-480): np.random.seed(0) then
    N(0, (sigma/255)^2) added to a seeded clean tile."""
    gt = W.seeded_image("gt", case["shape"], case["seed"])
    np.random.seed(0)
    noise = np.random.normal(0, case["sigma"] / 255.0, tuple(case["shape"])).astype(np.float32)
    return gt + torch.from_numpy(noise), gt


def guided_inputs(case):
    """lq = seeded image; ref = a shifted/noised copy when shapes agree (so matching is non-degenerate)."""
    lq = W.seeded_image("lq", case["lq"], case["seed"])
    if tuple(case["ref"]) == tuple(case["lq"]):
        ref = torch.roll(lq, shifts=(5, -7), dims=(2, 3)) + 0.02 * (W.seeded_image("n", case["ref"], case["seed"]) - 0.5)
    else:
        ref = W.seeded_image("ref", case["ref"], case["seed"])
    return lq, ref


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    for name, case in RESTORMER_CASES.items():
        net = R.restormer(**case["cfg"])
        W.load_seeded(net, case["seed"])
        x = W.seeded_image("x", case["shape"], case["seed"])
        y = net(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))
    for name, case in GUIDED_CASES.items():
        net = R.restormer_ref_fusion(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))


def main_promptir():
    from . import promptir as OP
    torch.set_grad_enabled(False)
    for name, case in PROMPTIR_CASES.items():
        net = R.promptir_ref_fusion(**case["cfg"])
        sd = W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        yo = OP.promptir_ref_fusion_forward(sd, lq, ref, case["cfg"]["heads"])
        meta = dict(case, oracle_vs_reference_max=float((y - yo).abs().max()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()), "oracle vs reference", meta["oracle_vs_reference_max"])


def main_drsformer():
    from . import drsformer as OD
    torch.set_grad_enabled(False)
    for name, case in DRS_CASES.items():
        net = R.drsformer_ref_fusion(spa=case["spa"], **case["cfg"]).eval()
        sd = W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        yo = OD.drsformer_ref_fusion_forward(sd, lq, ref, case["cfg"]["heads"])
        meta = dict(case, oracle_vs_reference_max=float((y - yo).abs().max()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()), "oracle vs reference", meta["oracle_vs_reference_max"])


VIT_CASES = {
    "dino_vit_tiny": dict(cfg=dict(img_size=70, patch_size=14, embed_dim=64, depth=2, num_heads=4, mlp_ratio=4,
                                   init_values=1.0, ffn_layer="mlp", block_chunks=0), seed=51, shape=(2, 3, 56, 84)),
    "clip_vit_tiny": dict(cfg=dict(hidden_size=80, intermediate_size=160, num_hidden_layers=2, num_attention_heads=1,
                                   patch_size=14, image_size=56, hidden_act="gelu"), seed=52, shape=(2, 3, 56, 56)),
    "mappers_tiny": dict(cfg=dict(input_dim=80, mid_dim=48, num_words=3), seed=53, shape=(2, 17, 80)),
}


def load_mapper_classes():
    """Mapper / CleanMapper live in a training script that imports accelerate/diffusers (absent): extract the two
    ClassDef nodes with ast and exec them unchanged (SURVEY 8c recipe)."""
    import ast
    src = open(os.path.join(R.REF_ROOT, "scripts", "train", "main_train_tr_mapping.py")).read()
    ns = {"nn": torch.nn, "torch": torch}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name in ("Mapper", "CleanMapper"):
            exec(compile(ast.Module([node], []), "main_train_tr_mapping.py", "exec"), ns)
    return ns["Mapper"], ns["CleanMapper"]


def main_vit():
    from functools import partial
    torch.set_grad_enabled(False)
    R._stub_packages()
    from models.dino.attention import MemEffAttention
    from models.dino.block import Block
    from models.dino.vision_transformers import DinoVisionTransformer
    case = VIT_CASES["dino_vit_tiny"]
    net = DinoVisionTransformer(block_fn=partial(Block, attn_class=MemEffAttention), **case["cfg"]).eval()
    W.load_seeded(net, case["seed"])
    y = net(W.seeded_image("x", case["shape"], case["seed"]))
    np.savez_compressed(os.path.join(OUT, "dino_vit_tiny.npz"), meta=json.dumps(case), out=y.numpy())
    print("dino_vit_tiny", tuple(y.shape))
    # CLIP: arithmetic lives in third-party transformers (reference pins 4.31.0; installed here: see fixture meta)
    import transformers
    case = dict(VIT_CASES["clip_vit_tiny"], transformers=transformers.__version__)
    net = transformers.CLIPVisionModel(transformers.CLIPVisionConfig(**case["cfg"])).eval()
    W.load_seeded(net, case["seed"])
    y = net(W.seeded_image("x", case["shape"], case["seed"]), output_hidden_states=True)[0]
    np.savez_compressed(os.path.join(OUT, "clip_vit_tiny.npz"), meta=json.dumps(case), out=y.numpy())
    print("clip_vit_tiny", tuple(y.shape), "transformers", transformers.__version__)
    Mapper, CleanMapper = load_mapper_classes()
    case = VIT_CASES["mappers_tiny"]
    c = case["cfg"]
    m, cm = Mapper(c["input_dim"], c["mid_dim"], c["num_words"]).eval(), CleanMapper(c["mid_dim"], c["mid_dim"], c["num_words"]).eval()
    W.load_seeded(m, case["seed"]); W.load_seeded(cm, case["seed"] + 1)
    emb = W.seeded_image("emb", case["shape"], case["seed"]) * 2 - 1
    w1 = m([emb]); w2 = cm(w1)
    np.savez_compressed(os.path.join(OUT, "mappers_tiny.npz"), meta=json.dumps(case), out=w1.numpy(), out2=w2.numpy())
    print("mappers_tiny", tuple(w1.shape), tuple(w2.shape))


def main_nafnet():
    torch.set_grad_enabled(False)
    for name, case in NAFNET_CASES.items():
        net = R.nafnet(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, _ = denoise_inputs(case)
        y = net(lq)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))
    for name, case in NAF_GUIDED_CASES.items():
        net = R.nafnet_ref_fusion(**case["cfg"])
        W.load_seeded(net, case["seed"])
        lq, ref = guided_inputs(case)
        y = net(lq, ref)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(case), out=y.numpy())
        print(name, tuple(y.shape), float(y.abs().max()))



GRAD_CASES = {"restormer_withbias": "restormer", "guided_restormer_128": "guided", "nafnet_rgb_ragged": "nafnet",
              "guided_nafnet_256": "guided_nafnet"}


def grad_probe(name, g):
    """Small fingerprint of a gradient tensor: (L2 norm, dot with a key-seeded N(0,1) vector)."""
    r = torch.randn(g.shape, generator=W._gen("probe:" + name, 0))
    return float(g.double().norm()), float((g.double() * r.double()).sum())


def main_grads():
    """Parameter gradients of the UNMODIFIED reference modules (autograd, L1 loss against a seeded target): pins the
    oracle's autograd -- the checker of the explicit backward schedule -- to the reference.  Fixture = per-parameter
    (norm, probe) fingerprints, not the full tensors."""
    torch.set_grad_enabled(True)
    for name, kind in GRAD_CASES.items():
        case = dict(restormer=RESTORMER_CASES, guided=GUIDED_CASES, nafnet=NAFNET_CASES, guided_nafnet=NAF_GUIDED_CASES)[kind][name]
        if kind == "nafnet":                       # LayerNormFunction's hand-written backward (nafnet_arch_utils.py:277-289)
            net = R.nafnet(**case["cfg"])
            W.load_seeded(net, case["seed"])
            x, gt = denoise_inputs(case)
            y = net(x)
        elif kind == "guided_nafnet":
            net = R.nafnet_ref_fusion(**case["cfg"])
            W.load_seeded(net, case["seed"])
            lq, ref = guided_inputs(case)
            gt = W.seeded_image("gt", case["lq"], case["seed"])
            y = net(lq, ref)
        elif kind == "restormer":
            net = R.restormer(**case["cfg"])
            W.load_seeded(net, case["seed"])
            x = W.seeded_image("x", case["shape"], case["seed"])
            gt = W.seeded_image("gt", case["shape"], case["seed"])
            y = net(x)
        else:
            net = R.restormer_ref_fusion(**case["cfg"])
            W.load_seeded(net, case["seed"])
            lq, ref = guided_inputs(case)
            gt = W.seeded_image("gt", case["lq"], case["seed"])
            y = net(lq, ref)
        loss = (y - gt).abs().mean()
        loss.backward()
        names, norms, probes = [], [], []
        for n, p in net.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            a, b = grad_probe(n, g)
            names.append(n); norms.append(a); probes.append(b)
        np.savez_compressed(os.path.join(OUT, name + "_grad.npz"), meta=json.dumps(case), loss=float(loss),
                            names=np.array(names), norms=np.array(norms), probes=np.array(probes))
        print(name + "_grad", len(names), float(loss), float(np.sqrt((np.array(norms) ** 2).sum())))
    torch.set_grad_enabled(False)

if __name__ == "__main__":
    main()
    main_grads()
    main_nafnet()
    main_vit()
    main_promptir()
    main_drsformer()
