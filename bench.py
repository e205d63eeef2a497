#!/usr/bin/env python
"""Headline benchmark: HSI SR patches/sec with full DDPM sampling on B200 (BASELINE.json metric).

Default workload = BASELINE configs[1] (C2), per GPU, weak scaling: a batch of 16 synthetic 128-band Chikusei-shaped
128x128 patches (bicubic x4 of a 32x32 cube), GAE geometry of GAE_4_Chi.pth (n_subs 16 / n_ovls 4 -> G = 11 groups), UNet of
config/sr_sr3_16_128ae.json, cosine schedule with T = 2000 steps: encode -> 2000 x (UNet forward on all 176 group latents
+ fused posterior step) -> decode -> clamp.  Random-init weights of that architecture, synthetic data (no UNet
checkpoint / dataset ships).

A "step" is one reverse-diffusion timestep over the whole latent batch (BASELINE's second metric, "UNet denoise step ms").
The timed region is EXACTLY K such steps (default K = T = 2000, i.e. one complete sampling pass) plus, when K == T, the GAE
encode before and decode after them; CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  `value` = patches of all ranks / that time.  With K < T the K timed steps are scaled to T and the separately timed
encode/decode are added ("full_sampling": false says so).

Other workloads (not what the driver runs; lines committed under profiles/):
  --workload c3   BASELINE configs[2]: one 102-band Pavia-Centre-shaped scene (1096x715) as 70 overlapping 128x128 tiles,
                  STRONG scaling: tiles sharded over the ranks, NCCL gather of the device tensors to rank 0 and the blend
                  on the GPU inside the timed region; value = tiles (patches) per second.
  --workload c4   BASELINE configs[3]: the 64_512 UNet at 512x512 on one Harvard-shaped cube (5 latents per step).
  --workload c5   BASELINE configs[4]: data-parallel TRAINING step (fp32 first version), value = cubes per second through training.
  --workload c1   BASELINE configs[0]: one 31-band CAVE-shaped cube, batch 1, T = 50 (5 group latents per step): the
                  reference's own call pattern; CPU arm run in full.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [...]                        # reference's CPU implementation (oracle/_ref, else the port) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...             # one rank per GPU, no per-step collective
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET = dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
            attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)       # config/sr_sr3_16_128ae.json
SCHED = dict(schedule="cosine", linear_start=1e-6, linear_end=1e-2)
UNET_GFLOP = 92.353        # per latent image per forward @128^2 (FlopCounterMode on the reference)
HW = 128
# name -> bands, GAE geometry, timesteps, cubes per GPU batch, GAE encode+decode GFLOP per cube
WORKLOADS = {
    "c2": dict(bands=128, geom=(128, 16, 4), T=2000, batch=16, gae_gflop=92.53 + 96.29,
               text="configs[1]: 128-band Chikusei-shaped 128x128 patches, GAE_4_Chi geometry (G=11), 4x SR, T=2000 cosine DDPM "
                    "sampling, batch 16 patches per GPU (176 group latents per step)"),
    "c3": dict(bands=102, geom=(102, 16, 4), T=2000, batch=16, gae_gflop=75.71 + 78.97, scene=(1096, 715), tile=128, overlap=16,
               text="configs[2]: 102-band Pavia-Centre-shaped scene 1096x715, GAE_4_Pav geometry (G=9), 70 overlapping 128x128 tiles "
                    "(overlap 16) sharded over the GPUs, T=2000 cosine DDPM sampling, NCCL gather + GPU blend in the timed region"),
    "c4": dict(bands=31, geom=(31, 8, 2), T=2000, batch=1, gae_gflop=16 * (41.30 + 43.23), hw=512, unet_gflop=1246.11,
               unet=dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16), attn_res=(),
                         res_blocks=1, dropout=0.2, image_size=128),
               text="configs[3]: config/sr_sr3_64_512.json UNet (155.3 M params, mid-block attention at 32x32 = 1024 tokens) on one 31-band "
                    "Harvard-shaped 512x512 cube, GAE_4_Har geometry (G=5): 5 latents of 512x512 per step, T=2000"),
    "c5": dict(bands=31, geom=(31, 8, 2), T=2000, batch=4, gae_gflop=41.30 + 43.23,
               text="configs[4]: training step (p_losses forward + hand-written backward + NCCL all-reduce of the gradient slab + Adam) of the "
                    "16_128ae UNet on 31-band Harvard-shaped 128x128 synthetic batches, 4 cubes per GPU, one optimiser step per band "
                    "group (G=5) like sr_gae.py:245-250; fp32 CUDA-core kernels (the bf16 tensor-core backward is not built yet)"),
    "c2v": dict(bands=128, geom=(128, 16, 4), T=20, batch=16, gae_gflop=92.53 + 96.29,
                text="configs[1] with the SHIPPED validation schedule (config/sr_sr3_16_128ae.json beta_schedule.val n_timestep = 20): "
                     "128-band Chikusei-shaped 128x128 patches, GAE_4_Chi geometry (G=11), batch 16 patches per GPU (176 group latents "
                     "per step); the GAE codec is a visible share of the pass here (use --gae-precision bf16 for its tensor-core mode)"),
    "c1": dict(bands=31, geom=(31, 8, 2), T=50, batch=1, gae_gflop=41.30 + 43.23,
               text="configs[0]: one 31-band CAVE-shaped 128x128 cube, GAE_4_Cav geometry (G=5), batch 1, T=50 cosine schedule "
                    "(5 group latents per step)"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_sustained=p["bf16_tflops_sustained"], tf_burst=p["bf16_tflops"], src="measured")
    return dict(hbm_gbs=6650.0, tf_sustained=1400.0, tf_burst=1590.0, src="fallback")


def traffic_table():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full) per kernel tag, from the committed
    capture summaries (profiles/r2_traffic.json, else the round-1 aggregate); {} if none is committed."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(wl: dict, steps: int, warmup: int, full: bool = False) -> dict:
    """The reference's CPU path, fp32, all host threads.  UNet forwards and the sampling loop run through the UNMODIFIED
    reference modules (model/sr3_modules/unet.py, diffusion.py byte-compiled into oracle/_ref by oracle/build_ref.py:
    kind "reference"); when oracle/_ref is absent, through the oracle port of the same files (kind "port").  The GAE codec
    (0.01 % of a patch; AE.py's loops hard-code 'cuda:0') always runs through the oracle port.

    Bounded sample (default): `steps` UNet forwards of ONE 128x128 group latent + one cube through GAE encode and decode,
    extrapolated linearly to a full patch (G*T forwards + codec); per-step cost is constant in t.
    full=True (workload c1): the whole patch - encode, T steps for each of the G groups at batch 1 like the reference's
    driver, decode - is executed and timed."""
    import torch
    from hsi_dmgasr_b200 import synth
    from hsi_dmgasr_b200.spec import GAEGeometry, UNetConfig
    from oracle import build_ref
    from oracle import hsidm_oracle as O      # the only place bench.py executes the oracle: as the measured CPU baseline

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, geom = UNetConfig(**UNET), GAEGeometry(*wl["geom"])
    usd, gsd = synth.unet_state_dict(cfg, 0), synth.gae_state_dict(geom, 1)
    T = wl["T"]
    ref = build_ref.load()
    kind = "reference" if ref else "port"
    if ref:
        unet_mod, diff_mod = ref
        net = unet_mod.UNet(in_channel=cfg.in_channel, out_channel=cfg.out_channel, norm_groups=cfg.norm_groups,
                            inner_channel=cfg.inner_channel, channel_mults=list(cfg.channel_mults), attn_res=list(cfg.attn_res),
                            res_blocks=cfg.res_blocks, dropout=cfg.dropout, image_size=cfg.image_size)
        net.load_state_dict(usd, strict=True)
        net.eval()
        forward = net
        what = "unmodified reference UNet.forward (oracle/_ref)"
    else:
        forward = lambda x, lv: O.unet_forward(usd, cfg.as_dict(), x, lv)   # noqa: E731
        what = "oracle port of UNet.forward"
    with torch.no_grad():
        if full:
            cube = synth.sr_cube(1, wl["bands"], HW, seed=2)
            forward(torch.randn(1, 6, HW, HW), torch.full((1, 1), 0.5))      # warm-up
            if ref:
                # sr_gae.py:444-467: encode, then per group GaussianDiffusion.super_resolution at batch 1, decode
                os.environ.setdefault("TQDM_DISABLE", "1")
                gd = diff_mod.GaussianDiffusion(net, image_size=HW, channels=3, conditional=True)
                gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), "cpu")
                t0 = time.perf_counter()
                zs = O.gae_encode(gsd, geom.as_dict(), cube)
                outs = [gd.super_resolution(z, False) for z in zs]
                O.gae_decode(gsd, geom.as_dict(), cube, outs)
                per_patch = time.perf_counter() - t0
            else:
                tab = O.schedule_tables(O.beta_schedule("cosine", T, 1e-6, 1e-2))
                x_T, tape = synth.noise_tape(geom.G, T, 3, HW, HW, seed=3)
                t0 = time.perf_counter()
                O.sr_cube(usd, cfg.as_dict(), tab, gsd, geom.as_dict(), cube, [x_T[g:g + 1] for g in range(geom.G)],
                          lambda g, i: tape[g:g + 1, T - 1 - i])
                per_patch = time.perf_counter() - t0
            return {"value": 1.0 / per_patch, "unit": "patches/s", "cores": cores, "kind": kind,
                    "sample": f"one complete patch, not extrapolated: GAE encode (port), {geom.G} groups x T={T} steps at batch 1 "
                              f"through {what}, GAE decode (port) ({per_patch:.1f} s)",
                    "ms_per_unet_step_per_latent": per_patch * 1e3 / (geom.G * T)}
        x = torch.randn(1, 6, HW, HW)
        lv = torch.full((1, 1), 0.5)
        for _ in range(max(1, min(warmup, 3))):
            forward(x, lv)
        t0 = time.perf_counter()
        for _ in range(steps):
            forward(x, lv)
        t_step = (time.perf_counter() - t0) / steps
        cube = synth.sr_cube(1, wl["bands"], HW, seed=2)
        t0 = time.perf_counter()
        zs = O.gae_encode(gsd, geom.as_dict(), cube)
        O.gae_decode(gsd, geom.as_dict(), cube, zs)
        t_codec = time.perf_counter() - t0
    per_patch = geom.G * T * t_step + t_codec
    return {"value": 1.0 / per_patch, "unit": "patches/s", "cores": cores, "kind": kind,
            "sample": f"{steps} forwards of one 6x{HW}x{HW} group latent through {what} ({t_step * 1e3:.1f} ms each) + 1 cube GAE "
                      f"encode+decode through the oracle port ({t_codec:.2f} s), extrapolated to G={geom.G} x T={T} forwards per patch",
            "ms_per_unet_step_per_latent": t_step * 1e3}


def gpu_library_leg(dev, n_lat: int) -> dict:
    """GPU LIBRARY baseline: the same network through stock PyTorch eager on this B200 - ATen + cuDNN +
    cuBLAS, i.e. what the reference's nn.Modules execute - via the oracle port of unet.py moved to the device.  Three
    precisions (fp32 with TF32 off, fp32 with TF32 on = the reference's defaults on Ampere+, bf16 autocast) and two call
    patterns (the reference's sequential batch-1 forwards, and all `n_lat` latents as one batch).  Reported as ms per
    denoise step of `n_lat` latents; a bounded sample (2 warm-up + 3 timed forwards each)."""
    import torch
    from hsi_dmgasr_b200 import synth
    from hsi_dmgasr_b200.spec import UNetConfig
    from oracle import hsidm_oracle as O      # baseline leg only: never on the product path

    cfg = UNetConfig(**UNET)
    usd = {k: v.to(dev) for k, v in synth.unet_state_dict(cfg, 0).items()}
    out = {"what": "oracle port of model/sr3_modules/unet.py under torch eager on cuda (ATen/cuDNN/cuBLAS), UNet forward only",
           "latents_per_step": n_lat, "ms_per_step": {}}
    keep = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True                       # sr_gae.py:149-150

    def timed(n, mode):
        x = torch.randn(n, 6, HW, HW, device=dev)
        lv = torch.full((n, 1), 0.5, device=dev)
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = False           # torch >= 1.12 default, as the reference runs
        def fwd():
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.unet_forward(usd, cfg.as_dict(), x, lv)
            return O.unet_forward(usd, cfg.as_dict(), x, lv)
        with torch.no_grad():
            for _ in range(2):
                fwd()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                fwd()
            b.record()
            torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / 3

    try:
        for mode in ("fp32", "tf32", "bf16_autocast"):
            try:
                t1 = timed(1, mode)
                tb = timed(n_lat, mode)
                out["ms_per_step"][mode] = {"sequential_batch1": t1 * n_lat, "batched": tb, "ms_per_forward_batch1": t1}
            except Exception as e:   # e.g. out of memory on a shared box: report, do not fail the bench
                out["ms_per_step"][mode] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = keep
    return out


def run_reference(args, wl) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    leg = cpu_reference_leg(wl, steps, args.warmup, full=args.workload == "c1")
    line = {"impl": "reference", "metric": "HSI SR patches/sec (full sampling)", "value": leg["value"], "unit": "patches/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": leg["ms_per_unet_step_per_latent"],
            "higher_is_better": True, "scaling": "strong" if args.workload == "c3" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, wl, wl["batch"]),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, wl, batch) -> dict:
    geom_g = -(-(wl["geom"][0] - wl["geom"][2]) // (wl["geom"][1] - wl["geom"][2]))
    return {"workload": wl["text"], "unet": "config/sr_sr3_64_512.json (155.3 M params)" if "unet" in wl else "config/sr_sr3_16_128ae.json (97.8 M params)",
            "timesteps": wl["T"],
            "patches_per_gpu": batch, "groups": geom_g,
            "parallelism": (f"tiles sharded over {args.gpus} GPU(s), one gather at the end" if args.workload == "c3"
                            else f"replica per GPU x{args.gpus}, patches sharded, no per-step collective"),
            "l2": "per-step working set (activations of all latents, >1 GB per layer at 176 latents) exceeds the 126 MB L2; no flush needed"
                  if args.workload != "c1" else "5-latent working set is L2 resident by nature of the workload (batch 1 is the reference's call pattern)"}


def roofline_leg(lib, gd, z, K, n_lat, ms_per_step, pk) -> dict:
    """Per-launch CUDA-event timing of ONE eager denoise step (UNet forward + posterior), repeated twice: every launch is
    bracketed by events on its own stream inside the library (hsidm_prof_*).  The dominant kernel family is the tcgen05
    implicit-GEMM convs; per-shape lines keep algorithmic and executed FLOPs apart (the sub-pixel upsample form executes
    16/36 of the reference's 3x3-over-upsampled FLOPs) and carry the ncu DRAM traffic of that shape when one is committed."""
    import torch
    from hsi_dmgasr_b200 import _lib
    _lib.check(lib.hsidm_prof_enable(1))
    x = torch.randn_like(z)
    reps = 2
    for _ in range(reps):
        gd.p_sample(x, K // 2, condition_x=z, noise=x)
    path = os.path.join(tempfile.gettempdir(), f"hsidm_prof_{os.getpid()}.csv")
    _lib.check(lib.hsidm_prof_dump(path.encode()))
    _lib.check(lib.hsidm_prof_enable(0))
    rows = list(csv.DictReader(open(path)))
    os.unlink(path)
    names = {0: "conv_tc", 1: "conv_simt", 2: "gn_finalize_or_stats", 3: "gn_apply", 4: "attn_gemm", 5: "posterior", 6: "other"}
    fam = {v: {"ms": 0.0, "work": 0.0, "launches": 0} for v in names.values()}
    shapes = {}
    for r in rows:
        f = fam[names[int(r["kind"])]]
        ms, work = float(r["ms"]) / reps, float(r["work"]) / reps
        f["ms"] += ms
        f["work"] += work
        f["launches"] += 1 / reps
        if int(r["kind"]) in (0, 4):
            s = shapes.setdefault(r["tag"], {"ms": 0.0, "flop": 0.0, "launches": 0})
            s["ms"] += ms
            s["flop"] += work
            s["launches"] += 1 / reps
    traffic = traffic_table()
    per_shape, exec_flop = [], 0.0
    for tag, s in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]):
        executed = s["flop"] * (16.0 / 36.0 if "up2x" in tag else 1.0)
        if tag.startswith(("halo", "pertap")):
            exec_flop += executed
        per_shape.append({"tag": tag, "launches_per_step": round(s["launches"], 1), "ms_per_step": round(s["ms"], 4),
                          "algorithmic_tflops": round(s["flop"] / (s["ms"] * 1e-3) / 1e12, 1) if s["ms"] else None,
                          "executed_tflops": round(executed / (s["ms"] * 1e-3) / 1e12, 1) if s["ms"] else None,
                          "frac_algorithmic": round(s["flop"] / (s["ms"] * 1e-3) / 1e12 / pk["tf_sustained"], 3) if s["ms"] else None,
                          "dram_bytes_per_launch_ncu": (traffic.get("per_tag") or {}).get(tag)})
    tc = fam["conv_tc"]
    achieved = tc["work"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    executed = exec_flop / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    post = fam["posterior"]
    hbm = {}
    for name in ("posterior", "gn_apply", "gn_finalize_or_stats"):
        f = fam[name]
        if f["ms"] > 0:
            gbs = f["work"] / (f["ms"] * 1e-3) / 1e9
            hbm[name] = {"ms_per_step": round(f["ms"], 4), "launches_per_step": round(f["launches"], 1), "algorithmic_gbs": round(gbs, 1),
                         "frac_of_hbm_peak": round(gbs / pk["hbm_gbs"], 3)}
    return {"kernel": "conv_halo_kernel<MT,BN,taps,pair> + conv_tc_kernel<BN> (tcgen05/TMEM/TMA implicit-GEMM convs incl. their fused "
                      "GroupNorm, all launches of one step)",
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
            "achieved_executed": executed, "frac_executed": executed / pk["tf_sustained"],
            "note": "achieved = the reference's algorithmic FLOPs / kernel time; *_executed counts the FLOPs the kernels really issue "
                    "(the sub-pixel upsample convs execute 16/36 of the reference's)",
            "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long step)",
            # DRAM bytes of ONE launch (ncu --set full) next to that launch's algorithmic bytes; the mean over the captured
            # launches of different layers is kept for continuity with round 1
            "traffic": traffic.get("dram_bytes_first_layer") or traffic.get("dram_bytes_per_launch"),
            "traffic_algorithmic": traffic.get("algorithmic_bytes_first_layer"), "traffic_layer": traffic.get("first_layer"),
            "traffic_layers": traffic.get("layers"), "traffic_mean_over_captures": traffic.get("dram_bytes_per_launch"),
            "traffic_note": traffic.get("source"),
            "launches_per_step": tc["launches"], "algorithmic_flop_per_step": tc["work"], "kernel_ms_per_step": tc["ms"],
            "step_breakdown_ms": {k: round(v["ms"], 3) for k, v in fam.items()},
            "per_shape": per_shape[:40], "hbm_kernels": hbm, "hbm_peak_gbs": pk["hbm_gbs"],
            "whole_step_tflops": UNET_GFLOP * n_lat / ms_per_step,
            "whole_step_frac_of_peak": UNET_GFLOP * n_lat / ms_per_step / pk["tf_sustained"]}


def run_training(args, wl, gd, gae, dev, rank, world, local, lib, barrier, max_over_ranks) -> None:
    """BASELINE configs[4]: K optimiser steps (default 10) of the reference's training loop body (sr_gae.py:236-250): encode HR
    and SR cubes once, then for each band group p_losses -> backward -> gradient all-reduce -> Adam."""
    import torch
    from hsi_dmgasr_b200 import synth
    from hsi_dmgasr_b200.diffusion import allreduce_gradients
    from hsi_dmgasr_b200.spec import GAEGeometry
    geom = GAEGeometry(*wl["geom"])
    B = args.batch or wl["batch"]
    K = 10 if args.steps >= wl["T"] else max(1, args.steps)
    W = max(args.warmup, 3)
    gd.train()
    gd.set_loss(dev)
    gd.set_new_noise_schedule(dict(SCHED, n_timestep=wl["T"]), dev)
    opt = torch.optim.Adam(list(gd.parameters()), lr=1e-5)          # config "train.optimizer.lr"
    hr = synth.sr_cube(B, wl["bands"], HW, seed=200 + rank).to(dev)
    sr = synth.sr_cube(B, wl["bands"], HW, seed=300 + rank).to(dev)
    z_hr, z_sr = gae.encode(hr), gae.encode(sr)

    def step(i):
        g = i % geom.G
        opt.zero_grad()
        l_pix = gd({"HR": z_hr[g], "SR": z_sr[g]})
        b, c, h, w = z_hr[g].shape
        l = l_pix.sum() / int(b * c * h * w)
        l.backward()
        allreduce_gradients(gd, world)
        opt.step()
        return l

    for i in range(W):
        step(i)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = lib.hsidm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(K):
        last = step(i)
    e1.record()
    barrier()
    t1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    clk = clocks.stop(t0, t1) if clocks else None
    if rank == 0:
        # fwd + bwd of a conv net = 3x the forward FLOPs (dgrad + wgrad), per latent image
        tflops = 3 * UNET_GFLOP * B / ms
        line = {"metric": "HSI SR patches/sec (training step, cubes through G optimiser steps)", "value": world * B / (geom.G * ms * 1e-3),
                "unit": "patches/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl, B),
                "latents_per_step": B, "loss": float(last), "clocks": clk, "gpu_launches": int(lib.hsidm_launch_count() - launches0),
                "achieved_tflops_fwd_bwd": tflops, "e2e": None, "roofline": None, "cpu_baseline": None,
                "note": "optimizer = torch.optim.Adam (as the reference); gradient all-reduce = one NCCL call on the contiguous slab"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed denoise steps K (default: the workload's full schedule T)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="patches (tiles) per sampling batch per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--gae-precision", default="fp32", choices=["bf16", "fp32"])
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end pass")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-eager GPU library baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    global UNET, HW, UNET_GFLOP
    UNET, HW, UNET_GFLOP = wl.get("unet", UNET), wl.get("hw", HW), wl.get("unet_gflop", UNET_GFLOP)
    T_FULL = wl["T"]
    if args.steps is None:
        args.steps = T_FULL
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from hsi_dmgasr_b200 import GAE, GaussianDiffusion, SRPipeline, UNet, _lib, synth
    from hsi_dmgasr_b200.pipeline import super_resolve_scene, tile_scene
    from hsi_dmgasr_b200.spec import GAEGeometry, UNetConfig

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hsidm hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun for N>1", file=sys.stderr)
    lib = _lib.load()
    pk = peaks()

    cfg, geom = UNetConfig(**UNET), GAEGeometry(*wl["geom"])
    W = max(args.warmup, 3)
    K = max(1, args.steps)
    full = K >= T_FULL
    K = T_FULL if full else K
    net = UNet(**{**UNET, "attn_res": list(UNET["attn_res"])}, precision=args.precision)
    net.load_state_dict(synth.unet_state_dict(cfg, 0))
    gd = GaussianDiffusion(net, image_size=128, channels=3, conditional=True).to(dev).eval()
    if args.workload == "c4":
        args.no_gpu_baseline = True     # the torch-eager leg is sized for the 128x128 workloads
    gae = GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=geom.n_feats, precision=args.gae_precision)
    gae.load_state_dict(synth.gae_state_dict(geom, 1))
    gae = gae.to(dev).eval()
    pipe = SRPipeline(gd, gae)
    B = args.batch or wl["batch"]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    extra = {}
    if args.workload == "c5":
        return run_training(args, wl, gd, gae, dev, rank, world, local, lib, barrier, max_over_ranks)
    if args.workload == "c3":
        # ---- strong scaling over the tiles of one scene ------------------------------------------------------------------
        hs, ws = wl["scene"]
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))   # torchrun pins OMP to 1 thread; the host-side tiling is timed
        scene = synth.sr_cube(1, wl["bands"], 4 * (-(-max(hs, ws) // 4)), seed=100)[0][:, :hs, :ws].contiguous()
        tiles_host, pos = tile_scene(scene, wl["tile"], wl["overlap"])
        n_tiles = tiles_host.shape[0]
        from hsi_dmgasr_b200.pipeline import shard_bounds
        lo, hi = shard_bounds(n_tiles, rank, world)
        n_lat = min(B, hi - lo) * geom.G
        # warm-up: W steps on this rank's first batch (builds workspaces, packs weights, captures the step graph)
        gd.set_new_noise_schedule(dict(SCHED, n_timestep=W), dev)
        pipe.super_resolve(tiles_host[lo:min(hi, lo + B)].to(dev), seed=1, first_cube=lo)
        if world > 1:   # ... and one small gather: NCCL sets up its point-to-point channels on first use
            from hsi_dmgasr_b200.pipeline import gather_rows
            gather_rows(torch.zeros((1, 8), device=dev), world, rank, world, dist)
        gd.set_new_noise_schedule(dict(SCHED, n_timestep=K), dev)
        # sampling-only time of this rank (for the K < T extrapolation): CUDA events around every sampling call
        samp_events = []
        orig_sr = gd.super_resolution

        def timed_sr(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_sr(*a, **k)
            e1.record()
            samp_events.append((e0, e1))
            return r
        gd.super_resolution = timed_sr
        host_scene = torch.empty(scene.shape, dtype=scene.dtype, pin_memory=True) if rank == 0 else None   # reusable result buffer
        barrier()
        clocks = ClockSampler(local) if rank == 0 else None
        launches0 = lib.hsidm_launch_count()
        barrier()
        t_wall0 = time.time()
        e_all = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e_all[0].record()
        out = super_resolve_scene(pipe, scene, dev, tile=wl["tile"], overlap=wl["overlap"], batch=B, rank=rank, world=world, seed=2)
        if rank == 0:
            host_scene.copy_(out, non_blocking=True)
        e_all[1].record()
        barrier()
        t_wall1 = time.time()
        gd.super_resolution = orig_sr
        launches = lib.hsidm_launch_count() - launches0
        clk = clocks.stop(t_wall0, t_wall1) if clocks else None
        total_meas = e_all[0].elapsed_time(e_all[1])
        loop_ms = sum(a.elapsed_time(b) for a, b in samp_events)
        if rank == 0:
            assert torch.isfinite(host_scene).all() and tuple(host_scene.shape) == tuple(scene.shape)
        steps_total = K * len(samp_events)
        ms_per_step = max_over_ranks(loop_ms) / max(1, steps_total)          # per step of one sampling batch on the slowest rank
        if full:
            total_ms = max_over_ranks(total_meas)
        else:   # scale the sampling share to T, keep the measured H2D + codec + gather + blend + D2H
            total_ms = max_over_ranks(total_meas - loop_ms + loop_ms * T_FULL / K)
        value = n_tiles / (total_ms * 1e-3)
        nbytes_in, nbytes_out = tiles_host.numel() * 4, scene.numel() * 4
        e2e = {"value": value, "unit": "patches/s", "h2d_bytes_per_step": nbytes_in / T_FULL, "d2h_bytes_per_step": nbytes_out / T_FULL,
               "h2d_bytes_per_pass": nbytes_in, "d2h_bytes_per_pass": nbytes_out,
               "api": "pipeline.super_resolve_scene (host scene -> tiles -> pinned H2D -> encode/sample/decode per batch -> NCCL gather "
                      "-> GPU blend -> pinned D2H of the scene): the timed region of `value` already is this call with host buffers",
               "timing": "CUDA events around the call on every rank, max over ranks"}
        extra = {"tiles": n_tiles, "tiles_this_rank": hi - lo, "sampling_calls_this_rank": len(samp_events),
                 "non_sampling_ms": total_meas - loop_ms, "ideal_speedup": n_tiles / max(
                     shard_bounds(n_tiles, r, world)[1] - shard_bounds(n_tiles, r, world)[0] for r in range(world))}
        enc_ms = dec_ms = None
        z = None
        patches_total = n_tiles
        scaling = "strong"
    else:
        n_lat = B * geom.G
        sr_host = synth.sr_cube(B, wl["bands"], HW, seed=100 + rank).pin_memory()
        sr = sr_host.to(dev)
        # ---- warm-up: W denoise steps (also builds workspaces, packs weights, captures the step graph) ---------------------
        gd.set_new_noise_schedule(dict(SCHED, n_timestep=W), dev)
        z = gae.encode_batched(sr)
        gd.super_resolution(z, return_all=True, seed=1)
        gae.decode_batched(z, clamp01=True)
        gd.set_new_noise_schedule(dict(SCHED, n_timestep=K), dev)
        if K <= 200:
            # A schedule change rebuilds the per-timestep tables and re-captures the step graph (T is baked into it); on
            # these boxes the first pass after it intermittently carries 0.1-0.2 s of driver-side one-off time
            # (scripts/c4_oneoff.py), which a 20-step timed region would book as 30-50 % more per step.  A short run
            # of at most 200 steps therefore takes one more untimed pass on the TIMED schedule; a T = 2000 pass (35 s) does
            # not need it.
            gd.super_resolution(z, return_all=True, seed=1)
        barrier()

        # ---- timed region: exactly K steps (+ encode/decode when K is the full schedule) ---------------------------------------
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        clocks = ClockSampler(local) if rank == 0 else None
        launches0 = lib.hsidm_launch_count()
        barrier()
        t_wall0 = time.time()
        ev[0].record()
        z = gae.encode_batched(sr)
        ev[1].record()
        lat = gd.super_resolution(z, return_all=True, seed=2)
        ev[2].record()
        out = gae.decode_batched(lat, clamp01=True)
        ev[3].record()
        barrier()
        t_wall1 = time.time()
        launches = lib.hsidm_launch_count() - launches0
        enc_ms, loop_ms, dec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
        clk = clocks.stop(t_wall0, t_wall1) if clocks else None
        assert torch.isfinite(out).all()
        _lib.check_health(dev)
        ms_per_step = max_over_ranks(loop_ms) / K
        if full:
            total_ms = max_over_ranks(enc_ms + loop_ms + dec_ms)
        else:
            total_ms = max_over_ranks(enc_ms + dec_ms) + ms_per_step * T_FULL
        value = world * B / (total_ms * 1e-3)
        patches_total = world * B
        scaling = "weak"

        # ---- end to end through the public API with HOST buffers (pinned H2D of the cubes, D2H of the SR cubes) -----------------
        e2e = None
        if not args.no_e2e:
            gd.set_new_noise_schedule(dict(SCHED, n_timestep=K), dev)
            barrier()
            t0 = time.perf_counter()
            res = pipe.super_resolve_host(sr_host, dev, seed=3)
            barrier()
            e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            assert res.shape == sr_host.shape
            if not full:   # K < T: scale the sampling part, keep the measured copies + codec
                e2e_ms = e2e_ms - max_over_ranks(loop_ms) + ms_per_step * T_FULL
            nbytes = sr_host.numel() * 4
            e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": nbytes / K,
                   "d2h_bytes_per_step": nbytes / K, "h2d_bytes_per_pass": nbytes, "d2h_bytes_per_pass": nbytes,
                   "extrapolated_from_steps": None if full else K,
                   "api": "SRPipeline.super_resolve_host (pinned host cubes -> GAE.encode -> GaussianDiffusion.super_resolution "
                          "-> GAE.decode -> host)", "timing": "host clock around the call, device synchronised on both sides"}

    # ---- roofline of the dominant kernel family + per-shape lines ------------------------------------------------------------
    roof = None
    if rank == 0:
        if z is None:
            z = torch.randn(n_lat, 3, HW, HW, device=dev)
        roof = roofline_leg(lib, gd, z, K, z.shape[0], ms_per_step, pk)

    cpu = gpu_base = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_reference_leg(wl, steps=4 if args.workload == "c4" else 20, warmup=1 if args.workload == "c4" else 2,
                                full=args.workload == "c1")
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        try:
            gpu_base = gpu_library_leg(dev, z.shape[0])
            best = min((v["batched"] for v in gpu_base["ms_per_step"].values() if "batched" in v), default=None)
            if best:
                gpu_base["speedup_of_this_library_over_best_eager"] = best / ms_per_step
                gpu_base["note"] = ("UNet forward only (no posterior step); this library's ms_per_step includes the posterior kernel")
        except Exception as e:
            gpu_base = {"error": str(e)[:200]}

    if rank == 0:
        line = {"metric": "HSI SR patches/sec (full sampling)", "value": value, "unit": "patches/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, wl, B),
                "full_sampling": full, "encode_ms": enc_ms, "decode_ms": dec_ms, "sampling_ms": loop_ms,
                "unet_denoise_step_ms": ms_per_step, "latents_per_step": n_lat, "patches": patches_total,
                "gae_precision": args.gae_precision, "clocks": clk, "e2e": e2e,
                "warmup_note": f"{W} untimed denoise steps on a {W}-step schedule (builds workspaces, packs weights)" +
                               ("" if K > 200 or args.workload == "c3" else f", then one untimed {K}-step pass on the timed schedule (table rebuild + graph re-capture happen there)"),
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "gpu_library_baseline": gpu_base, **extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
