"""Turns the raw output of scripts/gpu_r2_evidence.sh (gpurun_out/<tag>/) into the tracked round-2 summaries in profiles/:
bench lines, ncu launch list per kernel, ncu --set full tables with per-shape DRAM traffic, sanitizer summary, latency
protocol, layer profiles.  Usage: python scripts/make_r2_profiles.py [tag]   (default tag r2f)"""
import collections, csv, io, json, os, re, shutil, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"
src = f"gpurun_out/{tag}"
os.makedirs("profiles", exist_ok=True)


def short(name):
    name = re.sub(r"\(CUtensorMap.*|\(const .*|\(hsidm.*|\(float.*|\(int\s*\*.*", "", name)
    name = name.replace("void ", "").replace("hsidm::<unnamed>::", "").replace("hsidm::", "").replace("(int)", "").replace("(bool)", "")
    return name.strip()


def last_json(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1])
    except Exception:
        return None


# ---- 1. bench lines -----------------------------------------------------------------------------------------------------
for name in ("bench_full", "bench_c1", "bench_c4"):
    d = last_json(f"{src}/{name}.json")
    if d:
        json.dump(d, open(f"profiles/r2_{name}.json", "w"), indent=1)
lat = []
if os.path.exists(f"{src}/step_latency.jsonl"):
    lat = [json.loads(l) for l in open(f"{src}/step_latency.jsonl") if l.startswith("{")]
    json.dump({"protocol": "scripts/step_latency.py: CUDA events around 64-step CUDA-graph sampling passes, median pass time / 64",
               "rows": lat}, open("profiles/r2_step_latency.json", "w"), indent=1)
for name in ("layer_prof.txt", "layer_prof_n5.txt"):
    if os.path.exists(f"{src}/{name}"):
        shutil.copy(f"{src}/{name}", f"profiles/r2_{name}")

# ---- 2. launch list of the bench command ----------------------------------------------------------------------------------
if os.path.exists(f"{src}/bench_launches.csv"):
    lines = [l for l in open(f"{src}/bench_launches.csv") if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        v = agg.setdefault(k, [0, 0.0])
        v[0] += 1
        v[1] += float(r["Metric Value"].replace(",", "")) / 1e6   # ns -> ms
    step_kernels = ("conv_halo_kernel", "conv_tc_kernel", "gemm_tc_kernel", "gn_apply_kernel<__nv_bfloat16>", "gn_apply_t_kernel",
                    "gn_finalize_kernel", "im2col", "softmax_bf16", "posterior", "upsample2x_kernel<__nv_bfloat16>", "dec_kernel",
                    "gn_stats_kernel<__nv_bfloat16>")
    tot = sum(v[1] for v in agg.values())
    step_tot = sum(v[1] for k, v in agg.items() if k.startswith(step_kernels))
    with open("profiles/r2_bench_launches.md", "w") as f:
        f.write("# ncu launch list of the bench command (round 2, final code)\n\n")
        f.write("Command (B200, under `gpurun`): `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 3 "
                "--no-e2e --no-cpu --no-gpu-baseline`\n(3 warm-up + 2 timed denoise steps replayed from the CUDA graph, the eager roofline pass, "
                "GAE encode/decode in fp32; cold-cache, serialised per-launch times - compare SHARES, not absolutes).\n")
        f.write(f"{sum(v[0] for v in agg.values())} launches captured. Raw CSV: `r2_bench_launches.csv`.\n\n")
        f.write("| kernel | launches | total ms | share of all | share of the denoise-step kernels |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / tot < 0.0005:
                continue
            in_step = k.startswith(step_kernels)
            share = f"{100 * v[1] / step_tot:.1f} %" if in_step else "- (GAE codec / setup)"
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % | {share} |\n")
        conv = sum(v[1] for k, v in agg.items() if k.startswith(("conv_halo_kernel", "conv_tc_kernel")))
        f.write(f"\nTensor-core conv family (`conv_halo_kernel` + `conv_tc_kernel`, including the fused GroupNorm+Swish of their inputs and the "
                f"statistics fold): {100 * conv / step_tot:.1f} % of the denoise-step kernel time here; `roofline.step_breakdown_ms` of the bench "
                f"line (warm, CUDA events) gives the same share.  No `gn_finalize_kernel` launches remain in the step.\n")
    shutil.copy(f"{src}/bench_launches.csv", "profiles/r2_bench_launches.csv")


# ---- 3. ncu --set full tables ---------------------------------------------------------------------------------------------
def raw_table(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    return {h: i for i, h in enumerate(rows[0])}, rows[1], rows[2:]


def val(row, col, name):
    try:
        return float(row[col[name]].replace(",", ""))
    except Exception:
        return float("nan")


def mbytes(row, col, units, name):
    v = val(row, col, name)
    return v * {"Mbyte": 1, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(units[col[name]], 1) if name in col else float("nan")


captures = [("h464", "conv_halo_kernel<4,64,9,single CTA> (64-channel 128x128 layers)"),
            ("h2128", "conv_halo_kernel<2,128,9,single CTA> (128-channel 64x64 layers)"),
            ("h1256p", "conv_halo_kernel<1,256,9,CTA pair> (256/512-channel layers)"),
            ("pertap", "conv_tc_kernel<256> (per-tap kernel: 8x8 stage, attention projections)"),
            ("flash", "attn_flash_kernel (attention scores + softmax + P.Xn in one kernel)"),
            ("i2c", "im2col_small_kernel (first conv's K = 64 rows)"),
            ("fin", "conv_halo_kernel<4,16,9,single CTA> (last conv, 64 -> 3)"),
            ("applyt", "gn_apply_t_kernel (attention GroupNorm + transposed copy)"),
            ("post", "posterior_kernel")]
per_tag = {}
with open("profiles/r2_ncu_full.md", "w") as f:
    f.write("# ncu --set full of the dominant kernels (round 2, final code)\n\n")
    f.write("Command per row group: `ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:<kernel> -s <skip> -c <count> "
            "python scripts/step_time.py --precision bf16 --batches 176 --iters 1` (B200, 176 latents @128x128; the posterior kernel from a short "
            "`bench.py` run).  ncu times are cold and serialised: read the pipe / traffic columns, take times from the bench line.\n\n")
    f.write("| capture | launch | us | tensor pipe active % | TC smem data pipe % | dram read MB | dram write MB | DRAM throughput % | L2->SM sectors | regs | dyn smem KB |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|\n")
    for key, what in captures:
        path = f"{src}/{key}.raw.csv"
        if not os.path.exists(path) or os.path.getsize(path) < 1000:
            f.write(f"| {what} | capture failed | | | | | | | | | |\n")
            continue
        col, units, data = raw_table(path)
        for i, r in enumerate(data):
            rd, wr = mbytes(r, col, units, "dram__bytes_read.sum"), mbytes(r, col, units, "dram__bytes_write.sum")
            per_tag.setdefault(key, []).append((rd + wr) * 1e6)
            f.write(f"| {what} | {i} | {val(r, col, 'gpu__time_duration.sum'):.1f} | "
                    f"{val(r, col, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                    f"{val(r, col, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | {rd:.1f} | {wr:.1f} | "
                    f"{val(r, col, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed' if 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed' in col else 'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                    f"{val(r, col, 'lts__t_sectors_srcunit_tex_op_read.sum'):.3g} | {val(r, col, 'launch__registers_per_thread'):.0f} | "
                    f"{val(r, col, 'launch__shared_mem_per_block_dynamic') * {'Kbyte/block': 1.0, 'byte/block': 1 / 1024, 'Mbyte/block': 1024.0}.get(units[col['launch__shared_mem_per_block_dynamic']], 1.0) if 'launch__shared_mem_per_block_dynamic' in col else float('nan'):.0f} |\n")
    f.write("\nAlgorithmic bytes for comparison (bf16 NHWC, 176 latents): a 64-channel 128x128 tensor is 369 MB, a 128-channel 64x64 tensor 185 MB, "
            "a 256-channel 32x32 tensor 92 MB, a 512-channel 16x16 tensor 46 MB; a conv reads its input(s) once and writes its output once, so e.g. "
            "a 64 -> 64 conv at 128x128 has 738 MB of algorithmic traffic (+ 369 MB residual where it has one).  DRAM traffic close to or BELOW that "
            "figure means no wasted re-reads (the deeper layers stay far below: their operands still sit in the 126 MB L2).\n")
# The h464 capture skips four launches of conv_halo_kernel<4,64,9> (the two down blocks' conv1 / conv2) and takes the next
# three: conv1 and conv2 of the first 128x128 up block and conv1 of the second (unet.cu res_block order).  Algorithmic bytes
# = inputs read once + output written once, bf16 NHWC, 176 latents: a 64-channel 128x128 tensor is 369.1 MB.
T64 = 176 * 128 * 128 * 64 * 2
H464_LAYERS = [("halo MT4 BN64 cin128+64 cout64 128x128 n176 +nb +st +gn", "up block 1 conv1: reads 128 + 64 channels, writes 64", 3 * T64 + T64),
               ("halo MT4 BN64 cin64+0 cout64 128x128 n176 +st +gn", "up block 1 conv2 with the res_conv folded in: reads 64 channels + the 192-channel block input, writes 64", T64 + 3 * T64 + T64),
               ("halo MT4 BN64 cin64+64 cout64 128x128 n176 +nb +st +gn", "up block 2 conv1: reads 64 + 64 channels, writes 64", 2 * T64 + T64)]
layers, per_shape_tag = [], {}
if len(per_tag.get("h464", [])) == len(H464_LAYERS):
    for (ltag, what, alg), meas in zip(H464_LAYERS, per_tag["h464"]):
        layers.append({"tag": ltag, "layer": what, "dram_bytes": meas, "algorithmic_bytes": alg, "ratio": round(meas / alg, 3)})
        per_shape_tag[ltag] = meas
json.dump({"source": f"ncu --set full, {tag}: dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "layers": layers, "per_tag": per_shape_tag,
           "dram_bytes_first_layer": layers[0]["dram_bytes"] if layers else None,
           "algorithmic_bytes_first_layer": layers[0]["algorithmic_bytes"] if layers else None,
           "first_layer": (layers[0]["tag"] + " - " + layers[0]["layer"]) if layers else None,
           "per_capture_bytes_per_launch": {k: sum(v) / len(v) for k, v in per_tag.items()},
           "dram_bytes_per_launch": (sum(per_tag["h464"]) / len(per_tag["h464"])) if "h464" in per_tag else None,
           "kernel": "conv_halo_kernel<4,64,9> (the 64-channel 128x128 layers: largest share of the step)"},
          open("profiles/r2_traffic.json", "w"), indent=1)

# ---- 4. sanitizers ----------------------------------------------------------------------------------------------------------
with open("profiles/r2_sanitizer.md", "w") as f:
    f.write("# compute-sanitizer on the round-2 code\n\n")
    for tool in ("memcheck", "racecheck"):
        out = open(f"{src}/{tool}.out").read().strip().splitlines()[-2:] if os.path.exists(f"{src}/{tool}.out") else ["(not run)"]
        log = open(f"{src}/{tool}.log").read() if os.path.exists(f"{src}/{tool}.log") else ""
        summ = re.findall(r"(ERROR SUMMARY.*|RACECHECK SUMMARY.*)", log)
        f.write(f"## {tool}\n\npytest tail: `{' / '.join(out)}`\n\nsummary: `{'; '.join(summ) if summ else 'no summary line (see notes)'}`\n\n")
        haz = re.findall(r"(Race reported between .*?)\n", log)
        if haz:
            f.write("reports:\n\n" + "\n".join(f"* `{h[:260]}`" for h in sorted(set(haz))[:8]) + "\n\n")
print("profiles/r2_* written from", src)
