"""Mask post-processing helpers of the reference's SAM2 utilities on the CUDA connected-components kernel
(model/segment_anything_2/sam2/utils/misc.py:47-63 and :216-227). The VideoGLaMM entry points configure `fill_hole_area=8`
(build_sam.py:63-64) but never reach this code (sam2_video_predictor.py:971-975 is `pass`), so it is an exported operator, not part of
`inference()`."""
from __future__ import annotations

import torch

from . import ops


def get_connected_components(mask: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """misc.py:47-63: binary masks [N,1,H,W] (any dtype; nonzero/True = foreground) -> (labels, counts), both int32 [N,1,H,W]."""
    return ops.connected_components(mask.to(torch.uint8).contiguous())


def fill_holes_in_mask_scores(mask: torch.Tensor, max_area: int) -> torch.Tensor:
    """misc.py:216-227: background components (scores <= 0) of area <= max_area become foreground with score 0.1."""
    assert max_area > 0, "max_area must be positive"
    labels, areas = get_connected_components(mask <= 0)
    is_hole = (labels > 0) & (areas <= max_area)
    return torch.where(is_hole, 0.1, mask)
