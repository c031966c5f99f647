"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the frozen ViT encoders and the mapper MLPs.

  * DINOv2 ViT (reference: /root/reference/models/dino/vision_transformers.py:179-266,320-326, block.py:43-114,
    attention.py:36-69, mlp.py, layer_scale.py, patch_embed.py) -- returns the normed PATCH tokens, as the reference's
    ``forward`` does (``self.head`` is Identity).
  * CLIP vision tower: the arithmetic lives in third-party ``transformers`` (pinned 4.31.0 by the reference's
    requirements.txt:2, not vendored).  Restated from the published architecture and pinned HERE against the installed
    transformers 5.5.0 ``CLIPVisionModel`` (make_golden.py); PARITY WITH 4.31.0 IS UNPINNED.
    Call sites: scripts/train/main_train_tr_mapping.py:609,780 (``image_encoder(image, output_hidden_states=True)[0]``).
  * Mapper / CleanMapper: scripts/train/main_train_tr_mapping.py:40-81, 84-122.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _mha(x, wqkv, bqkv, wo, bo, heads):
    B, N, D = x.shape
    hd = D // heads
    qkv = (x @ wqkv.t() + bqkv).view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    attn = torch.softmax((qkv[0] * hd ** -0.5) @ qkv[1].transpose(-1, -2), -1)
    return (attn @ qkv[2]).transpose(1, 2).reshape(B, N, D) @ wo.t() + bo


def dino_pos_embed(pos_embed, h, w, patch, offset=0.1):
    """interpolate_pos_encoding (:179-207)."""
    n = pos_embed.shape[1] - 1
    w0, h0 = w // patch, h // patch
    if w0 * h0 == n and w == h:
        return pos_embed
    dim = pos_embed.shape[-1]
    s = int(math.sqrt(n))
    sx, sy = float(w0 + offset) / math.sqrt(n), float(h0 + offset) / math.sqrt(n)
    pp = F.interpolate(pos_embed[:, 1:].float().reshape(1, s, s, dim).permute(0, 3, 1, 2), scale_factor=(sx, sy),
                       mode="bicubic")
    assert pp.shape[-2] == w0 and pp.shape[-1] == h0
    return torch.cat([pos_embed[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, dim)], 1)


def dino_vit_forward(sd, x, heads=12, patch=14):
    B, _, H, W = x.shape
    assert H % patch == 0 and W % patch == 0
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], 1)
    t = t + dino_pos_embed(sd["pos_embed"], W, H, patch)      # the reference passes (w, h) = x.shape[2:] names swapped
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p = f"blocks.{i}."
        y = _mha(_ln(t, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6), sd[p + "attn.qkv.weight"],
                 sd[p + "attn.qkv.bias"], sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"], heads)
        t = t + y * sd[p + "ls1.gamma"]
        y = _ln(t, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.gelu(y @ sd[p + "mlp.fc1.weight"].t() + sd[p + "mlp.fc1.bias"]) @ sd[p + "mlp.fc2.weight"].t() + sd[p + "mlp.fc2.bias"]
        t = t + y * sd[p + "ls2.gamma"]
        i += 1
    return _ln(t, sd["norm.weight"], sd["norm.bias"], 1e-6)[:, 1:]


def dino_select_crop(sd, lq, ref, heads=12, patch=14):
    """Reference-crop selection, models/image_restoration_ref_model.py:215-247: all h x w crops of ref at stride h//4,
    both sides bilinearly resized to ceil(h/14)*14, cosine similarity of the flattened patch tokens, top-1 crop."""
    B, C, h, w = lq.shape
    stride = h // 4
    unf = F.unfold(ref, kernel_size=(h, h), stride=(stride, stride))
    n = unf.shape[-1]
    crops = unf.transpose(-1, -2).contiguous().view(B * n, C, h, h)
    size = (int(math.ceil(h / patch) * patch), int(math.ceil(w / patch) * patch))
    f_l = dino_vit_forward(sd, F.interpolate(lq, size=size, mode="bilinear"), heads, patch).reshape(B, 1, -1)
    f_r = dino_vit_forward(sd, F.interpolate(crops, size=size, mode="bilinear"), heads, patch).reshape(B, n, -1)
    corr = F.normalize(f_l, dim=-1) @ F.normalize(f_r, dim=-1).transpose(-1, -2)
    idx = corr.argmax(-1)[:, 0]
    return crops.view(B, n, C, h, h)[torch.arange(B), idx], idx, corr


def clip_vision_forward(sd, pixel_values, heads, patch, eps=1e-5, act="gelu", prefix="vision_model."):
    """CLIPVisionTransformer up to (not including) post_layernorm -> last_hidden_state [B, 1 + N, D]."""
    g = lambda k: sd[prefix + k]
    B = pixel_values.shape[0]
    t = F.conv2d(pixel_values, g("embeddings.patch_embedding.weight"), None, stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat([g("embeddings.class_embedding").expand(B, 1, -1), t], 1) + g("embeddings.position_embedding.weight")
    t = _ln(t, g("pre_layrnorm.weight"), g("pre_layrnorm.bias"), eps)
    i = 0
    while prefix + f"encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"encoder.layers.{i}."
        wqkv = torch.cat([g(p + f"self_attn.{n}_proj.weight") for n in "qkv"], 0)
        bqkv = torch.cat([g(p + f"self_attn.{n}_proj.bias") for n in "qkv"], 0)
        t = t + _mha(_ln(t, g(p + "layer_norm1.weight"), g(p + "layer_norm1.bias"), eps), wqkv, bqkv,
                     g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"), heads)
        y = _ln(t, g(p + "layer_norm2.weight"), g(p + "layer_norm2.bias"), eps) @ g(p + "mlp.fc1.weight").t() + g(p + "mlp.fc1.bias")
        y = F.gelu(y) if act == "gelu" else y * torch.sigmoid(1.702 * y)
        t = t + y @ g(p + "mlp.fc2.weight").t() + g(p + "mlp.fc2.bias")
        i += 1
    return t


def _mlp4(sd, p, x):
    for j in (0, 3, 6):
        x = x @ sd[f"{p}.{j}.weight"].t() + sd[f"{p}.{j}.bias"]
        x = F.leaky_relu(_ln(x, sd[f"{p}.{j + 1}.weight"], sd[f"{p}.{j + 1}.bias"], 1e-5), 0.01)
    return x @ sd[f"{p}.9.weight"].t() + sd[f"{p}.9.bias"]


def mapper_forward(sd, emb, num_words):
    """Mapper.forward (:73-81): per word, MLP(CLS) + mean over patch tokens of MLP_patch(patches)."""
    out = [_mlp4(sd, f"mapping_{i}", emb[:, :1]) + _mlp4(sd, f"mapping_patch_{i}", emb[:, 1:]).mean(1, keepdim=True)
           for i in range(num_words)]
    return torch.cat(out, 1)


def clean_mapper_forward(sd, embs, num_words):
    """CleanMapper.forward (:106-122): word i goes through its own MLP."""
    return torch.cat([_mlp4(sd, f"mapping_{i}", embs[:, i:i + 1]) for i in range(num_words)], 1)
