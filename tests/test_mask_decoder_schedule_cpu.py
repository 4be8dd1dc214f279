"""The row-block schedule planned for the fused mask-decoder kernel (oracle/mask_decoder_rowblock.py: image rows resident per block, THREE
grid-wide reductions, everything else block-local) reproduces the oracle's mask decoder exactly (fp32 reassociation only)."""
import pytest
import torch

from oracle import mask_decoder_rowblock as R, sam2 as O, weights as Wt


@pytest.mark.parametrize("n_blocks", [1, 4, 16])
def test_rowblock_schedule_equals_oracle(n_blocks):
    cfg = O.Sam2Cfg(**O.TINY)
    sd = Wt.make_sam2_state_dict(cfg, 1234)
    torch.manual_seed(0)
    B, C, fs = 2, cfg.d_model, cfg.feat_size
    emb = torch.randn(B, C, fs, fs)
    text = torch.randn(B, 1, C)
    sparse, dense = O.prompt_encode(sd, cfg, B, text)
    hi = [torch.randn(B, C // 8, 4 * fs, 4 * fs), torch.randn(B, C // 4, 2 * fs, 2 * fs)]
    with torch.no_grad():
        want = O.mask_decoder(sd, cfg, emb, O.dense_pe(sd, cfg), sparse, dense, hi, True)
        got = R.mask_decoder_rowblock(sd, cfg, emb, O.dense_pe(sd, cfg), sparse, dense, hi, n_blocks=n_blocks)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert (a - b).abs().max() <= 2e-5 * b.abs().max().clamp_min(1.0), (a - b).abs().max()
