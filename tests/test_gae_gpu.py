"""GAE encode/decode parity on the B200 (AE.py:283-324) against reference outputs in tests/golden/gae.npz."""
import pytest
import torch

from hsi_dmgasr_b200 import GAE, synth
from tests.cfgs import GAE_CASES
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def build(geom, seed, precision):
    gae = GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=geom.n_feats, precision=precision)
    gae.load_state_dict(synth.gae_state_dict(geom, seed), strict=True)
    return gae.cuda().eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("tag", list(GAE_CASES))
def test_gae_encode_decode(golden, tag, precision, tol):
    geom, seed, hw = GAE_CASES[tag]
    g = golden("gae.npz")
    gae = build(geom, seed, precision)
    x = synth.sr_cube(2, geom.n_colors, hw, seed=seed + 100).cuda()
    zs = gae.encode(x)
    assert len(zs) == geom.G and zs[0].shape == (2, 3, hw, hw)
    want_z = torch.from_numpy(g[f"{tag}.z"])
    err_z = rel_l2(torch.stack(zs), want_z)
    # decode the REFERENCE latents so encode and decode are gated independently
    y = gae.decode(x, [want_z[k].cuda() for k in range(geom.G)])
    err_y = rel_l2(y, torch.from_numpy(g[f"{tag}.dec"]))
    print(f"{tag} {precision}: encode {err_z:.2e} decode {err_y:.2e}")
    assert err_z < tol and err_y < tol
    y2, z2 = gae(x)
    assert y2.shape == x.shape and len(z2) == geom.G


def test_gae_batched_layout_and_clamp():
    geom, seed, hw = GAE_CASES["Cav"]
    gae = build(geom, seed, "fp32")
    x = synth.sr_cube(3, geom.n_colors, hw, seed=7).cuda()
    z = gae.encode_batched(x)
    zs = gae.encode(x)
    assert z.shape == (3 * geom.G, 3, hw, hw)
    for b in range(3):
        for k in range(geom.G):
            # (not bit-equal: the CALayer pooling uses float atomics, so two runs may differ in the last ulp)
            assert torch.allclose(z[b * geom.G + k], zs[k][b], rtol=1e-5, atol=1e-6)
    y = gae.decode_batched(z)
    yc = gae.decode_batched(z, clamp01=True)
    assert torch.equal(yc, y.clamp(0, 1))
    # each cube is independent of its batch neighbours
    y1 = gae.decode_batched(gae.encode_batched(x[1:2]))
    assert rel_l2(y1, y[1:2]) < 1e-6
