"""Image helpers of the render modes: `tile_images` (mani_skill/utils/visualization/misc.py:54-115) and the sensor-observation ->
image conversion of the cameras (mani_skill/sensors/camera.py:256-294).  Torch tensors in, torch tensors out (on their device)."""
from __future__ import annotations

from typing import Dict, List

import torch


def tile_images(images: List[torch.Tensor], nrows: int = 1) -> torch.Tensor:
    """Tile (batched [B, H, W, 3] or single [H, W, 3]) images into one: sorted by height (tallest first, stable), stacked into columns
    of the first image's height x nrows while the width matches, columns placed left to right, the remainder black."""
    batched = images[0].dim() == 4
    b = 1 if batched else 0
    if nrows == 1:
        images = sorted(images, key=lambda x: x.shape[b], reverse=True)
    max_h = images[0].shape[b] * nrows
    cur_h, cur_w = 0, images[0].shape[b + 1]
    columns, column = [], []
    for im in images:
        if cur_h + im.shape[b] <= max_h and cur_w == im.shape[b + 1]:
            column.append(im)
            cur_h += im.shape[b]
        else:
            columns.append(column)
            column = [im]
            cur_h, cur_w = im.shape[b], im.shape[b + 1]
    columns.append(column)
    total_w = sum(c[0].shape[b + 1] for c in columns)
    shape = ((images[0].shape[0],) if batched else ()) + (max_h, total_w, 3)
    out = torch.zeros(shape, dtype=images[0].dtype, device=images[0].device)
    x = 0
    for c in columns:
        col = torch.cat(c, dim=b)
        out[..., :col.shape[b], x:x + col.shape[b + 1], :] = col
        x += col.shape[b + 1]
    return out


def normalize_depth(depth, min_depth=0, max_depth=None):
    """camera.py:256-263."""
    if min_depth is None:
        min_depth = depth.min()
    if max_depth is None:
        max_depth = depth.max()
    return ((depth - min_depth) / (max_depth - min_depth)).clip(0, 1)


def camera_observations_to_images(observations: Dict[str, torch.Tensor], max_depth=None) -> Dict[str, torch.Tensor]:
    """camera.py:266-294: rgb as is, depth / position normalised to grey levels, segmentation ids hashed to colours."""
    images = dict()
    for key, val in observations.items():
        if "rgb" in key or "Color" in key:
            rgb = val[..., :3]
            if rgb.dtype == torch.float:
                rgb = torch.clip(rgb * 255, 0, 255).to(torch.uint8)
            images[key] = rgb
        elif "depth" in key or "position" in key:
            depth = -val[..., 2:3] if "position" in key else val
            depth = (normalize_depth(depth, max_depth=max_depth) * 255).clip(0, 255).to(torch.uint8)
            images[key] = torch.repeat_interleave(depth, 3, dim=-1)
        elif "segmentation" in key:
            assert val.dim() == 4 and val.shape[-1] == 1, val.shape
            images[key] = (val * torch.tensor([11, 61, 127], device=val.device)).to(torch.uint8)
    return images
