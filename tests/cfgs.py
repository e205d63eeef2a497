"""UNet / GAE configurations shared by the tests (same literals as oracle/make_golden.py)."""
from hsi_dmgasr_b200.spec import GAEGeometry, UNetConfig

SMALL = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2),
                   attn_res=(8,), res_blocks=1, dropout=0.2, image_size=16)
FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
WIDE = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16),
                  attn_res=(), res_blocks=1, dropout=0.0, image_size=128)
UNET_CASES = {  # tag -> (cfg, weight seed, N, HW, noise levels)   [oracle/make_golden.py section 2]
    "small": (SMALL, 11, 3, 16, [0.9, 0.35, 0.01]),
    "full32": (FULL, 12, 2, 32, [0.71, 0.05]),
    "full128": (FULL, 12, 1, 128, [0.5]),
    "wide64": (WIDE, 13, 1, 64, [0.2]),
}
GAE_CASES = {  # tag -> (geometry, weight seed, HW)                  [oracle/make_golden.py section 4]
    "Cav": (GAEGeometry(31, 8, 2), 41, 16),
    "Chi": (GAEGeometry(128, 16, 4), 42, 16),
    "Pav": (GAEGeometry(102, 16, 4), 43, 16),
}
