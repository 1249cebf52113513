"""Pretrained-weight loading with the reference's semantics (semilearn/nets/utils.py:18-73, called by every ViT builder,
vit.py:352-354): `pretrained_path` is a local file or a URL (the shipped usb_cv YAMLs give https URLs; torch.hub caches
them), the weights sit under the checkpoint's 'model' key, a DataParallel 'module.' prefix is stripped, classifier tensors
(keys starting with fc / classifier / mlp / head) are dropped, and `pos_embed` is resampled bicubically to the model's token
grid.  Loading is non-strict, like the reference."""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

_SKIP_PREFIXES = ("fc", "classifier", "mlp", "head")


def resize_pos_embed_vit(posemb, posemb_new, num_tokens=1, gs_new=()):
    """[1, T_old, D] -> [1, T_new, D]: prefix tokens kept, the square patch grid resampled (bicubic, align_corners False)."""
    n_new = posemb_new.shape[1] - num_tokens
    tok, grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
    g_old = int(math.sqrt(grid.shape[0]))
    if not len(gs_new):
        gs_new = (int(math.sqrt(n_new)),) * 2
    grid = grid.reshape(1, g_old, g_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=tuple(gs_new), mode="bicubic", align_corners=False)
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_new[0] * gs_new[1], -1)
    return torch.cat([tok, grid], dim=1)


def load_checkpoint(model, checkpoint_path):
    if checkpoint_path and os.path.isfile(checkpoint_path):
        ck = torch.load(checkpoint_path, map_location="cpu")
    else:
        ck = torch.hub.load_state_dict_from_url(checkpoint_path, map_location="cpu")
    state = {}
    for key, val in ck["model"].items():
        if key.startswith("module"):
            key = key.split(".", 1)[1]
        if key.startswith(_SKIP_PREFIXES):
            continue
        if key == "pos_embed" and hasattr(model, "pos_embed") and val.shape != model.pos_embed.shape:
            val = resize_pos_embed_vit(val, model.pos_embed.data)
        state[key] = val
    print(model.load_state_dict(state, strict=False))
    if hasattr(model, "mark_weights_updated"):
        model.mark_weights_updated()
    return model
