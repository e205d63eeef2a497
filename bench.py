#!/usr/bin/env python
"""Headline benchmark: HSI SR patches/sec with full DDPM sampling on B200 (BASELINE.json metric, config[1]).

Workload (per GPU, weak scaling): a batch of 16 synthetic 128-band Chikusei-shaped 128x128 patches (bicubic x4 of a
32x32 cube), GAE geometry of GAE_4_Chi.pth (n_subs 16 / n_ovls 4 -> G = 11 groups), UNet of config/sr_sr3_16_128ae.json,
cosine schedule with T = 2000 steps: encode -> 2000 x (UNet forward on all 176 group latents + fused posterior step)
-> decode -> clamp.  Random-init weights of that architecture, synthetic data (no checkpoints / datasets ship).

A "step" is one reverse-diffusion timestep over the whole 176-latent batch (BASELINE's second metric, "UNet denoise
step ms").  The timed region is EXACTLY K such steps (default K = T = 2000, i.e. one complete sampling pass) plus, when
K == T, the GAE encode before and decode after them; CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks.  `value` = patches of all ranks / that time.  With K < T the K timed steps are scaled to T and
the separately timed encode/decode are added ("full_sampling": false says so).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [...]                        # reference's CPU implementation (oracle port) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...             # one rank per GPU, no data-path collective
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_FULL = 2000
BANDS, HW, GEOM = 128, 128, (128, 16, 4)          # Chikusei-shaped patches, GAE_4_Chi geometry
UNET = dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
            attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)       # config/sr_sr3_16_128ae.json
SCHED = dict(schedule="cosine", linear_start=1e-6, linear_end=1e-2)
UNET_GFLOP = 92.353        # per latent image per forward @128^2 (FlopCounterMode on the reference)
GAE_GFLOP = 92.53 + 96.29  # encode + decode per Chikusei cube


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_sustained=p["bf16_tflops_sustained"], tf_burst=p["bf16_tflops"], src="measured")
    return dict(hbm_gbs=6650.0, tf_sustained=1400.0, tf_burst=1590.0, src="fallback")


def traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/r1_traffic.json); None if no capture is committed."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(path))["dram_bytes_per_launch"]
    except Exception:
        return None


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
def cpu_reference_leg(steps: int, warmup: int) -> dict:
    """The reference's CPU path (oracle port of unet.py / AE.py, fp32, all host threads) on a bounded sample: `steps`
    UNet forwards of ONE 128x128 group latent + one Chikusei cube through GAE encode and decode, extrapolated linearly
    to a full patch (G*T forwards + codec); per-step cost is constant in t."""
    import torch
    from hsi_dmgasr_b200 import synth
    from hsi_dmgasr_b200.spec import GAEGeometry, UNetConfig
    from oracle import hsidm_oracle as O      # the only place bench.py executes the oracle: as the measured CPU baseline

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, geom = UNetConfig(**UNET), GAEGeometry(*GEOM)
    usd, gsd = synth.unet_state_dict(cfg, 0), synth.gae_state_dict(geom, 1)
    x = torch.randn(1, 6, HW, HW)
    lv = torch.full((1, 1), 0.5)
    with torch.no_grad():
        for _ in range(max(1, min(warmup, 3))):
            O.unet_forward(usd, cfg.as_dict(), x, lv)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.unet_forward(usd, cfg.as_dict(), x, lv)
        t_step = (time.perf_counter() - t0) / steps
        cube = synth.sr_cube(1, BANDS, HW, seed=2)
        t0 = time.perf_counter()
        zs = O.gae_encode(gsd, geom.as_dict(), cube)
        O.gae_decode(gsd, geom.as_dict(), cube, zs)
        t_codec = time.perf_counter() - t0
    per_patch = geom.G * T_FULL * t_step + t_codec
    return {"value": 1.0 / per_patch, "unit": "patches/s", "cores": cores, "kind": "port",
            "sample": f"{steps} UNet forwards of one 6x{HW}x{HW} group latent ({t_step * 1e3:.1f} ms each) + 1 cube GAE "
                      f"encode+decode ({t_codec:.2f} s), extrapolated to G={geom.G} x T={T_FULL} forwards per patch",
            "ms_per_unet_step_per_latent": t_step * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    leg = cpu_reference_leg(steps, args.warmup)
    line = {"impl": "reference", "metric": "HSI SR patches/sec (full sampling)", "value": leg["value"], "unit": "patches/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": leg["ms_per_unet_step_per_latent"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 16), "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, batch) -> dict:
    return {"workload": f"configs[1]: {BANDS}-band Chikusei-shaped {HW}x{HW} patches, GAE_4_Chi geometry (G=11), 4x SR, "
                        f"T={T_FULL} cosine DDPM sampling, batch {batch} patches per GPU (176 group latents per step)",
            "unet": "config/sr_sr3_16_128ae.json (97.8 M params)", "timesteps": T_FULL, "patches_per_gpu": batch,
            "parallelism": f"replica per GPU x{args.gpus}, patches sharded, no per-step collective",
            "l2": "per-step working set (activations of 176 latents, >1 GB per layer) exceeds the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=T_FULL)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="patches per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end pass")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hsi_dmgasr_b200 import GAE, GaussianDiffusion, SRPipeline, UNet, _lib, synth
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

    cfg, geom = UNetConfig(**UNET), GAEGeometry(*GEOM)
    W = max(args.warmup, 3)
    K = max(1, args.steps)
    full = K >= T_FULL
    K = T_FULL if full else K
    net = UNet(**{**UNET, "attn_res": list(UNET["attn_res"])}, precision=args.precision)
    net.load_state_dict(synth.unet_state_dict(cfg, 0))
    gd = GaussianDiffusion(net, image_size=128, channels=3, conditional=True).to(dev).eval()
    gae = GAE(n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors, n_feats=geom.n_feats)
    gae.load_state_dict(synth.gae_state_dict(geom, 1))
    gae = gae.to(dev).eval()
    pipe = SRPipeline(gd, gae)
    B = args.batch
    n_lat = B * geom.G
    sr_host = synth.sr_cube(B, BANDS, HW, seed=100 + rank).pin_memory()
    sr = sr_host.to(dev)

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

    # ---- warm-up: W denoise steps (also builds workspaces, packs weights, captures the step graph) ---------------------
    gd.set_new_noise_schedule(dict(SCHED, n_timestep=W), dev)
    z = gae.encode_batched(sr)
    gd.super_resolution(z, return_all=True, seed=1)
    gae.decode_batched(z, clamp01=True)
    gd.set_new_noise_schedule(dict(SCHED, n_timestep=K), dev)
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
    ms_per_step = max_over_ranks(loop_ms) / K
    if full:
        total_ms = max_over_ranks(enc_ms + loop_ms + dec_ms)
    else:
        total_ms = max_over_ranks(enc_ms + dec_ms) + ms_per_step * T_FULL
    value = world * B / (total_ms * 1e-3)

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
               "api": "SRPipeline.super_resolve_host (pinned host cubes -> GAE.encode -> GaussianDiffusion.super_resolution "
                      "-> GAE.decode -> host)", "timing": "host clock around the call, device synchronised on both sides"}

    # ---- roofline of the dominant kernel (tcgen05 conv family): CUDA events around every launch of one eager step -----------
    roof = None
    if rank == 0:
        _lib.check(lib.hsidm_prof_enable(1))
        x = torch.randn_like(z)
        gd.predict_noise(x, K // 2, z)
        gd.predict_noise(x, K // 2, z)
        ms, work, n = C.c_double(), C.c_double(), C.c_int64()
        shares = {}
        names = {0: "conv_tc", 1: "conv_simt", 2: "gn_stats", 3: "gn_apply", 4: "attn_gemm"}
        for kind, name in names.items():
            _lib.check(lib.hsidm_prof_read(kind, C.byref(ms), C.byref(work), C.byref(n)))
            shares[name] = {"ms_per_step": ms.value / 2, "work_per_step": work.value / 2, "launches_per_step": n.value // 2}
        _lib.check(lib.hsidm_prof_enable(0))
        tc = shares["conv_tc"]
        achieved = tc["work_per_step"] / (tc["ms_per_step"] * 1e-3) / 1e12 if tc["ms_per_step"] > 0 else 0.0
        roof = {"kernel": "conv_halo_kernel<MT,BN,taps,pair> + conv_tc_kernel<BN> (tcgen05/TMEM/TMA implicit-GEMM convs incl. their fused GroupNorm, all launches of one step)",
                "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"], "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long step)",
                "traffic": traffic_from_profiles(), "launches_per_step": tc["launches_per_step"],
                "algorithmic_flop_per_step": tc["work_per_step"], "kernel_ms_per_step": tc["ms_per_step"],
                "step_breakdown_ms": {k: round(v["ms_per_step"], 3) for k, v in shares.items()},
                "gn_apply_gbs": (shares["gn_apply"]["work_per_step"] / (shares["gn_apply"]["ms_per_step"] * 1e-3) / 1e9
                                 if shares["gn_apply"]["ms_per_step"] > 0 else None),
                "gn_stats_gbs": (shares["gn_stats"]["work_per_step"] / (shares["gn_stats"]["ms_per_step"] * 1e-3) / 1e9
                                 if shares["gn_stats"]["ms_per_step"] > 0 else None),
                "hbm_peak_gbs": pk["hbm_gbs"],
                "whole_step_tflops": UNET_GFLOP * n_lat / ms_per_step,
                "whole_step_frac_of_peak": UNET_GFLOP * n_lat / ms_per_step / pk["tf_sustained"]}

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_reference_leg(steps=20, warmup=2)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "HSI SR patches/sec (full sampling)", "value": value, "unit": "patches/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, B),
                "full_sampling": full, "encode_ms": enc_ms, "decode_ms": dec_ms, "sampling_ms": loop_ms,
                "unet_denoise_step_ms": ms_per_step, "latents_per_step": n_lat, "clocks": clk, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
