"""CPU-side checks (no GPU): host logic of the drop-in mirrors, the C-ABI library surface, sharding."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import _lib, synth
from hsi_dmgasr_b200.config import NoneDict, dict_to_nonedict, parse
from hsi_dmgasr_b200.pipeline import shard_bounds
from hsi_dmgasr_b200.schedule import BUFFER_NAMES, diffusion_buffers, make_beta_schedule
from hsi_dmgasr_b200.spec import GAE_PRESETS, GAEGeometry, UNetConfig, gae_param_shapes, unet_layers, unet_param_shapes
from tests.cfgs import FULL, SMALL, WIDE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


# ---- C ABI surface -------------------------------------------------------------------------------------------
def declared_symbols():
    names = []
    for header in ("hsidm.h", "hsidm_debug.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        names += re.findall(r"HSIDM_API\s+[\w\s\*]+?\b(hsidm_\w+)\s*\(", text)
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 28 and len(set(names)) == len(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and headers drifted apart"
    assert lib.hsidm_version() == 100
    assert lib.hsidm_launch_count() == 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.UNetCfg(6, 3, 32, 8, 2, (C.c_int32 * 8)(1, 2), 1, (C.c_int32 * 8)(8), 1, 0.0, 16, _lib.F32)
    out = C.c_void_p()
    rc = lib.hsidm_ctx_create(C.byref(cfg), 0, C.byref(out))
    assert rc == -4 and b"no CUDA device" in lib.hsidm_last_error() and not out.value
    from hsi_dmgasr_b200 import UNet
    net = UNet(inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=[8], res_blocks=1, image_size=16)
    with pytest.raises(_lib.HsidmError):
        net(torch.zeros(1, 6, 16, 16), torch.ones(1, 1))
    # the product package must not route through the oracle
    pkg = os.path.join(ROOT, "hsi_dmgasr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(pkg, fn)).read(), re.M), fn


# ---- topology / state_dict layout --------------------------------------------------------------------------
def test_unet_topology_matches_survey_appendix_b():
    downs, mid, ups = unet_layers(FULL)
    assert (len(downs), len(mid), len(ups)) == (15, 2, 19)
    assert [L.name for L in downs if L.attn] == ["downs.10", "downs.11"]
    assert [L.name for L in ups if L.attn] == ["ups.4", "ups.5", "ups.6"]
    assert [(L.cin, L.cout) for L in ups if L.kind == "res"][3:6] == [(1024, 512), (1024, 512), (768, 512)]
    shapes = unet_param_shapes(FULL)
    assert len(shapes) == 362 and sum(int(np.prod(s)) for s in shapes.values()) == 97807491
    # sr_sr3_64_512.json: image_size 128 and empty attn_res -> only the mid block attends
    d, m, u = unet_layers(WIDE)
    assert not any(L.attn for L in d + u) and m[0].attn
    assert sum(int(np.prod(s)) for s in unet_param_shapes(WIDE).values()) == 155334339


def test_python_mirror_exposes_reference_state_dict_keys():
    from hsi_dmgasr_b200 import GAE, GaussianDiffusion, UNet
    net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
               res_blocks=2, dropout=0.2, image_size=128)
    assert list(net.state_dict().keys()) == list(unet_param_shapes(FULL).keys())
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == unet_param_shapes(FULL)[k], k
    gd = GaussianDiffusion(net, image_size=128, channels=3, conditional=True)
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=50, linear_start=1e-6, linear_end=1e-2), "cpu")
    keys = list(gd.state_dict().keys())
    assert len(keys) == 374 and keys[:12] == list(BUFFER_NAMES) and keys[12] == "denoise_fn.noise_level_mlp.1.weight"
    for name, geom in GAE_PRESETS.items():
        gae = GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=64)
        assert list(gae.state_dict().keys()) == list(gae_param_shapes(geom).keys()), name
        assert gae.geometry() == geom and (gae.start_idx, gae.end_idx) == geom.groups()


def test_group_layouts_from_survey():
    assert GAE_PRESETS["Cav"].groups() == ([0, 6, 12, 18, 23], [8, 14, 20, 26, 31])
    s, e = GAE_PRESETS["Chi"].groups()
    assert (len(s), s[-2:], e[-2:]) == (11, [108, 112], [124, 128])
    s, e = GAE_PRESETS["Pav"].groups()
    assert (len(s), s[-2:], e[-2:]) == (9, [84, 86], [100, 102])
    assert set(GAE_PRESETS["Cav"].band_counts()) == {1, 2}


def test_gae_state_file_round_trip(tmp_path):
    """the tensors-only GAE file written by save_gae_state loads with weights_only=True semantics and
    reproduces geometry, key order and every tensor; a bare state_dict file still loads too."""
    import torch
    from hsi_dmgasr_b200 import load_gae, save_gae_state, synth
    from hsi_dmgasr_b200.gae import GAE
    geom = GAE_PRESETS["Pav"]
    gae = GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=geom.n_feats)
    gae.load_state_dict(synth.gae_state_dict(geom, 5))
    path = save_gae_state(gae, str(tmp_path / "gae_state.pth"))
    blob = torch.load(path, map_location="cpu", weights_only=True)          # no pickled classes inside
    assert blob["format"] == "hsidm-gae-state-v1" and blob["geometry"]["n_colors"] == 102
    back = load_gae(path)
    assert back.geometry() == geom and list(back.state_dict().keys()) == list(gae.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(back.state_dict().values(), gae.state_dict().values()))
    bare = str(tmp_path / "bare_sd.pth")
    torch.save(gae.state_dict(), bare)
    assert load_gae(bare).geometry() == geom


def test_gae_unpickler_refuses_code_execution(tmp_path):
    """The remapping unpickler resolves an exact allow-list only: a checkpoint whose __reduce__ names builtins.eval (or
    anything else off the list) is refused before anything runs, and allow_pickle=False never leaves weights_only."""
    import pickle
    import torch
    from hsi_dmgasr_b200 import load_gae

    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return (eval, (f"open({str(marker)!r}, 'w').close()",))

    bad = str(tmp_path / "evil.pth")
    torch.save({"x": Evil()}, bad)
    with pytest.raises(pickle.UnpicklingError):
        load_gae(bad)
    assert not marker.exists()

    class Evil2:
        def __reduce__(self):
            import os as _os
            return (_os.system, ("true",))

    bad2 = str(tmp_path / "evil2.pth")
    torch.save([Evil2()], bad2)
    with pytest.raises(pickle.UnpicklingError):
        load_gae(bad2)
    with pytest.raises(pickle.UnpicklingError):
        load_gae(bad2, allow_pickle=False)


def test_whole_module_pickle_round_trip_without_reference(tmp_path):
    """A whole-module pickle shaped like GAE_pretrained/GAE_4_*.pth (classes named __main__.* / common.*, AE.py:637) built
    from this package's own classes loads through the allow-listed unpickler - runs on boxes without /root/reference."""
    import torch
    from hsi_dmgasr_b200 import gae as G, load_gae, synth
    geom = GAE_PRESETS["Cav"]
    m = G.GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=geom.n_feats)
    m.load_state_dict(synth.gae_state_dict(geom, 9))
    path = str(tmp_path / "GAE_4_fake.pth")
    G.save_reference_style_pickle(m, path)
    import zipfile
    names = zipfile.ZipFile(path).read([n for n in zipfile.ZipFile(path).namelist() if n.endswith("data.pkl")][0])
    assert b"__main__" in names and b"common" in names and b"hsi_dmgasr_b200" not in names
    back = load_gae(path)
    assert isinstance(back, G.GAE) and back.geometry() == geom
    assert all(torch.equal(a, b) for a, b in zip(back.state_dict().values(), m.state_dict().values()))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_gae_pickles_load_through_the_shim():
    from hsi_dmgasr_b200 import load_gae
    from hsi_dmgasr_b200.gae import GAE
    want = {"Cav": 583004, "Har": 583004, "Chi": 648197, "Pav": 633195}
    for name, n_params in want.items():
        gae = load_gae(os.path.join(REF, "GAE_pretrained", f"GAE_4_{name}.pth"))
        assert isinstance(gae, GAE) and gae.geometry() == GAE_PRESETS[name]
        assert sum(p.numel() for p in gae.parameters()) == n_params
        assert list(gae.state_dict().keys()) == list(gae_param_shapes(GAE_PRESETS[name]).keys())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_config_schema_and_define_G():
    from hsi_dmgasr_b200.networks import define_G
    opt = parse(os.path.join(REF, "config", "sr_sr3_16_128ae.json"), phase="val")
    assert isinstance(opt, NoneDict) and opt["distributed"] is True and opt["missing"] is None
    assert opt["model"]["unet"]["norm_groups"] is None and opt["model"]["beta_schedule"]["val"]["n_timestep"] == 20
    opt["path"]["resume_state"] = None
    opt["gpu_ids"] = None                      # CPU box: skip the reference's `assert torch.cuda.is_available()`
    netG = define_G(opt)
    assert netG.denoise_fn.cfg == FULL and netG.conditional is True
    opt2 = parse(os.path.join(REF, "config", "sr_sr3_64_512.json"), phase="val", gpu_ids="0")
    assert opt2["distributed"] is False and define_G(opt2).denoise_fn.cfg == WIDE
    assert parse(os.path.join(REF, "config", "sr_sr3_16_128ae.json"), debug=True)["model"]["beta_schedule"]["val"]["n_timestep"] == 10


def test_ddpm_wrapper_checkpoint_round_trip(tmp_path):
    from hsi_dmgasr_b200.model import DDPM, _DROPPED_ON_LOAD
    unet = dict(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_multiplier=[1, 2], attn_res=[8],
                res_blocks=1, dropout=0.2)
    sched = dict(schedule="cosine", n_timestep=6, linear_start=1e-6, linear_end=1e-2)
    opt = dict_to_nonedict(dict(phase="val", gpu_ids=None, distributed=False,
                                path=dict(resume_state=None, checkpoint=str(tmp_path)),
                                model=dict(which_model_G="sr3", finetune_norm=False, unet=unet,
                                           beta_schedule=dict(train=sched, val=sched),
                                           diffusion=dict(image_size=16, channels=3, conditional=True))))
    m = DDPM(opt)
    path = m.save_network(epoch=3, iter_step=40)
    assert path.endswith("I40_E3_gen.pth")
    saved = torch.load(path)
    assert list(saved.keys()) == list(m.netG.state_dict().keys())
    opt["path"]["resume_state"] = path[:-len("_gen.pth")]
    m2 = DDPM(opt)
    for k, v in m2.netG.state_dict().items():
        same = torch.equal(v, saved[k])
        assert same != (k in _DROPPED_ON_LOAD), k        # model.py:189-192 drops exactly these three
    # the training step exists but, like everything on the path, only on the GPU: a val-phase wrapper has no
    # optimiser, and a CPU tensor is refused loudly rather than falling back
    from hsi_dmgasr_b200._lib import HsidmError
    with pytest.raises(HsidmError):
        m2.netG.p_losses({"HR": torch.zeros(1, 3, 16, 16), "SR": torch.zeros(1, 3, 16, 16)})


# ---- schedules -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,T", [("cosine", 20), ("cosine", 50), ("cosine", 2000), ("linear", 30), ("quad", 30),
                                    ("warmup10", 30), ("warmup50", 30), ("const", 10), ("jsd", 10)])
def test_product_schedules_bit_exact_vs_reference(golden, name, T):
    g = golden("schedules.npz")
    tabs = diffusion_buffers(make_beta_schedule(name, T, 1e-6, 1e-2))
    for k in list(BUFFER_NAMES) + ["sqrt_alphas_cumprod_prev"]:
        assert np.array_equal(tabs[k], g[f"{name}{T}.{k}"], equal_nan=True), k


# ---- sharding -----------------------------------------------------------------------------------------------------
def test_shard_bounds_partition():
    for n in (0, 1, 7, 16, 70):
        for world in (1, 2, 4, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(70, r, 8) for r in range(8)][:3] == [(0, 9), (9, 18), (18, 27)]   # C3: 70 tiles on 8 GPUs


def test_synthetic_inputs_are_reproducible():
    a, b = synth.sr_cube(2, 31, 16, seed=3), synth.sr_cube(2, 31, 16, seed=3)
    assert torch.equal(a, b) and float(a.min()) >= 0 and float(a.max()) <= 1
    x, tape = synth.noise_tape(3, 6, 3, 8, 8, seed=1)
    assert x.shape == (3, 3, 8, 8) and tape.shape == (3, 5, 3, 8, 8)
    sd = synth.unet_state_dict(SMALL, 1)
    assert list(sd.keys()) == list(unet_param_shapes(SMALL).keys())


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract's keys, `impl: reference`, a
    cpu_baseline describing this run, e2e = the line's own value with zero copy bytes, and - when oracle/_ref has been
    built - the unmodified reference modules as the thing that was timed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "patches/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and "workload" in line["config"]
    from oracle import build_ref
    assert cb["kind"] == ("reference" if build_ref.load() else "port")
    assert line["gpu_launches"] == 0


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
