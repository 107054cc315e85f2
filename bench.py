#!/usr/bin/env python
"""Benchmark of the sketch-guided SD1.5 sampling path (BASELINE.json metric: 512x512 50-step images/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle) on the host cores

A bench "step" is ONE IMAGE: the whole 50-step DDIM sampling loop (26 guided steps) of one 512x512 sample with
CFG 7.5 and LGP sketch guidance, batch 1 per GPU (BASELINE.json configs[1]); with N GPUs every rank samples its own
images (weak scaling, no per-step collective; the only collective is the start-up weight broadcast).
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_INFERENCE_STEPS = 50
GUIDANCE = 7.5
METRIC = "512x512 50-step sketch-guided SD1.5 images/sec"


def guided_count(n):
    return sum(1 for i in range(n) if i <= 0.5 * n)       # modules/pipeline.py:90-92


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_dram_bytes_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/r1_gemm_tma_ncu_v2.txt: mean over the captured launches); None if the summary is not there."""
    path = os.path.join(ROOT, "profiles", "r1_gemm_tma_ncu_v2.txt")
    if not os.path.exists(path):
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for line in open(path):
        if line.startswith("dram__bytes_read.sum [") or line.startswith("dram__bytes_write.sum ["):
            unit = line.split("[")[1].split("]")[0]
            vals = [float(x) for x in line.split(":", 1)[1].split("|")]
            tot += sum(vals) * mult.get(unit, 1.0)
            n = len(vals)
    return tot / n if n else None


# ------------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_sample(threads, repeats=1, warmup=0):
    """The reference's CPU path (oracle/port.py: the restated loop body of modules/pipeline.py over the diffusers
    shim, torch CPU ops; fp32 UNet + fp16 LGP) on a bounded sample of the workload: one guided and one unguided
    denoising step of the SD1.5 512x512 job.  Returns per-repeat (t_guided, t_unguided) seconds."""
    import torch
    from oracle import port
    torch.set_num_threads(threads)
    unet = port.make_unet("sd15")
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    sch = port.make_scheduler()
    sch.set_timesteps(NUM_INFERENCE_STEPS)
    taps, _ = port.register_taps(unet)
    noise = lat.clone()
    out = []
    for r in range(warmup + repeats):
        times = []
        for guided in (True, False):
            t = sch.timesteps[0] if guided else sch.timesteps[-1]
            t0 = time.perf_counter()
            x_in = torch.cat([lat] * 2).requires_grad_(True)
            with torch.enable_grad() if guided else torch.no_grad():
                eps = unet(x_in, t, encoder_hidden_states=emb).sample
            eu, ec = eps.detach().chunk(2)
            e = eu + GUIDANCE * (ec - eu)
            new = sch.step(e, t, lat, eta=0.0).prev_sample
            if guided:
                with torch.enable_grad():
                    new = port.anti_gradient(lgp, sch, taps, x_in, new, noise, t, tgt, 1.6)
            times.append(time.perf_counter() - t0)
        if r >= warmup:
            out.append(tuple(times))
    return out


def images_per_sec_from_steps(tg, tu):
    g = guided_count(NUM_INFERENCE_STEPS)
    return 1.0 / (g * tg + (NUM_INFERENCE_STEPS - g) * tu)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    res = cpu_reference_sample(threads, repeats=args.steps, warmup=min(args.warmup, 1))
    tg = sum(r[0] for r in res) / len(res)
    tu = sum(r[1] for r in res) / len(res)
    ips = images_per_sec_from_steps(tg, tu)
    sample = ("each bench step = 1 guided + 1 unguided denoising step of the 50-step job (timesteps 981 and 1); "
              "images/sec = 1 / (26 t_guided + 24 t_unguided)")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * (tg + tu), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SD1.5 512x512 50-step DDIM CFG=7.5 + LGP sketch guidance, batch 1 (configs[1])",
                   "cpu_ms_guided_step": 1e3 * tg, "cpu_ms_unguided_step": 1e3 * tu},
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sketch2img_b200 import _lib, synthetic
    from sketch2img_b200 import distributed as D
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    from sketch2img_b200.unet import SD15_CONFIG, UNet2DConditionModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sketch-guided path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    _lib.lib()      # fail loudly if libs2i.so is missing
    cfg = dict(SD15_CONFIG)

    # ---- weights: rank 0 generates, one NCCL broadcast over NVLink distributes (the only collective of the job)
    t0 = time.time()
    if rank == 0:
        sd = synthetic.unet_state_dict(cfg, seed=1138)
    else:
        sd = {k: torch.empty(shape) for k, shape in synthetic.unet_param_shapes(cfg).items()}
    if world > 1:
        sd = D.broadcast_state_dict(sd, src=0, device=dev)
    unet = UNet2DConditionModel(cfg, sd, device=dev)
    del sd
    torch.manual_seed(1139)
    lgp = LatentEdgePredictor(synthetic.lgp_input_dim(cfg), 4, 9)
    with torch.no_grad():
        for m in lgp.layers:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.add_(0.1 * torch.randn_like(m.weight))
                m.bias.add_(0.1 * torch.randn_like(m.bias))
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    setup_s = time.time() - t0

    n_img = args.warmup + args.steps
    lat_h, emb_h, tgt_h = synthetic.sample_inputs(cfg, n_img, seed=1139 + 1000 * rank, pin=True)

    def emb_of(k):
        return torch.stack([emb_h[k], emb_h[n_img + k]])        # [uncond, cond]

    def sample_resident(lat, emb, tgt):
        return pipe("synthetic", num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=GUIDANCE, latents=lat,
                    sketch_image=tgt, prompt_embeds=emb, output_type="latent")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident inputs for the kernel-side number
    lat_d = lat_h.to(dev)
    emb_d = [emb_of(k).to(dev) for k in range(n_img)]
    tgt_d = tgt_h.to(dev)
    for k in range(args.warmup):
        sample_resident(lat_d[k:k + 1], emb_d[k], tgt_d[k:k + 1])
    barrier()

    clocks = ClockSampler(local)
    clocks.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(args.warmup, n_img):
        sample_resident(lat_d[k:k + 1], emb_d[k], tgt_d[k:k + 1])
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_total = D.max_over_ranks(ev0.elapsed_time(ev1), dev)

    # ---- end to end: host (pinned) inputs in, final latent back on the host, every image
    out_h = torch.empty(1, 4, lat_h.shape[2], lat_h.shape[3]).pin_memory()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    h2d = d2h = 0
    for k in range(args.warmup, n_img):
        lat = lat_h[k:k + 1].to(dev, non_blocking=True)
        emb = emb_of(k).pin_memory().to(dev, non_blocking=True)
        tgt = tgt_h[k:k + 1].to(dev, non_blocking=True)
        res = sample_resident(lat, emb, tgt)
        out_h.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller holds the image before asking for the next
        h2d = lat.numel() * 4 + emb.numel() * 4 + tgt.numel() * 4
        d2h = res.numel() * 4
    ev3.record()
    barrier()
    clk = clocks.stop()
    ms_e2e = D.max_over_ranks(ev2.elapsed_time(ev3), dev)

    ips = world * args.steps / (ms_total * 1e-3)
    ips_e2e = world * args.steps / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        # ---- per-kernel-class device time of one image (CUDA events around every libs2i launch, same stream)
        tf_peak, hbm_peak, peak_src = measured_peaks()
        _lib.profile_begin(tf_peak, hbm_peak)
        sample_resident(lat_d[:1], emb_d[0], tgt_d[:1])
        prof = _lib.profile_end()
        fl = synthetic.flops_per_image(cfg, NUM_INFERENCE_STEPS, guided_count(NUM_INFERENCE_STEPS))
        # dominant kernel: the tcgen05 + TMA implicit GEMM (gemm_tma_kernel / gemm_tc_kernel; profiler classes gemm_*).
        # achieved = sum of the algorithmic FLOPs of its launches (2 M N K per launch, recorded at each launch site)
        #            / sum of their CUDA-event durations on the launch stream (graph replay off for this one image)
        gemm = {k: v for k, v in prof.items() if k.startswith("gemm")}
        gemm_ms = sum(v["ms"] for v in gemm.values())
        gemm_n = sum(v["launches"] for v in gemm.values())
        gemm_fl = sum(v["flops"] for v in gemm.values())
        all_ms = sum(v["ms"] for v in prof.values())
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        gemm_roof_ms = sum(v["roof_ms"] for v in gemm.values())
        gemm_bytes = sum(v["bytes"] for v in gemm.values())

        def tensor_class(tag):
            v = prof.get(tag)
            if not v or v["ms"] <= 0:
                return None
            a = v["flops"] / (v["ms"] * 1e-3) / 1e12
            return {"launches": v["launches"], "ms_per_image": round(v["ms"], 3), "achieved_tflops": round(a, 1),
                    "frac_of_peak": round(a / tf_peak, 4)}
        roofline = {"bound": "tensor", "kernel": "gemm_tma_kernel + gemm_tc_kernel (tcgen05 + TMA implicit GEMM: conv3x3 / conv1x1 / "
                                                 "Linear / LGP MLP, forward and input-gradient)",
                    "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                    "traffic": ncu_dram_bytes_per_launch(), "peak_source": peak_src,
                    "algorithmic_flops_per_launch_mean": gemm_fl / gemm_n if gemm_n else None,
                    "algorithmic_bytes_per_launch_mean": gemm_bytes / gemm_n if gemm_n else None,
                    "per_launch_roofline": {
                        "frac": gemm_roof_ms / gemm_ms if gemm_ms > 0 else None, "roofline_ms_per_image": round(gemm_roof_ms, 3),
                        "hbm_peak_gbs": hbm_peak,
                        "how": "sum over the launches of max(algorithmic FLOPs / tensor peak, algorithmic bytes / HBM peak) / sum "
                               "of their measured durations: at B = 2 the projections with the fp32 residual stream sit left of "
                               "the ridge (64 FLOP/B for a 320 -> 320 projection), so the tensor-only fraction above understates "
                               "them; algorithmic bytes = activation operand + weights once (fp16), fp32 residual, outputs"},
                    "launches_per_image": gemm_n, "kernel_ms_per_image": gemm_ms,
                    "share_of_device_time": gemm_ms / all_ms if all_ms else None,
                    # the per-launch events above force plain launches; the timed region replays the step from a CUDA graph
                    # (no host gaps, programmatic dependent launch): the class's share of the event-timed image applied to
                    # the graph-replayed image time gives its throughput inside the timed region
                    "achieved_in_timed_region_estimate": (gemm_fl / (gemm_ms / all_ms * ms_total / args.steps * 1e-3) / 1e12
                                                          if gemm_ms > 0 and all_ms > 0 else None),
                    "algorithmic_flops_per_image_all_kernels": fl["image"],
                    "other_tensor_kernels": {"attn_fwd_kernel": tensor_class("attn_fwd"), "attn_bwd_kernel": tensor_class("attn_bwd")},
                    "note": "B = 2 (one CFG pair): every operand is L2-resident (ncu: ~0 DRAM bytes per launch, profiles/"
                            "r1_gemm_tma_ncu_v2.txt); the binding resource is the chip-wide L2 -> SM operand rate and per-launch "
                            "latency, not HBM or the tensor pipe (DESIGN.md section 4)",
                    "how": "sum of algorithmic FLOPs of one image's GEMM launches / sum of their CUDA-event durations "
                           "(s2i_profile_begin/end on the launch stream)"}
        breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in
                     sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        line = {
            "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate (tcgen05 kind::f16), fp32 residual stream",
            "data": "synthetic",
            "config": {"workload": "SD1.5 512x512 50-step DDIM CFG=7.5 + LGP sketch guidance, batch 1 per GPU (configs[1])",
                       "num_inference_steps": NUM_INFERENCE_STEPS, "guided_steps": guided_count(NUM_INFERENCE_STEPS),
                       "images_per_gpu_per_bench_step": 1, "ms_per_denoise_step": ms_total / args.steps / NUM_INFERENCE_STEPS,
                       "l2": "per-step working set (1.7 GB fp16 weights + >1 GB activations) exceeds the 126 MB L2; no flush needed",
                       "weights": "random init (seed 1138), broadcast once from rank 0", "setup_s": round(setup_s, 1)},
            "e2e": {"value": ips_e2e, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "kernel_breakdown_ms_per_image": breakdown,
            "kernel_breakdown_note": "one extra image run eagerly with a CUDA event per launch (graph replay off): it includes host "
                                     "launch gaps and the per-GEMM split-K zero-fill launches (gemm_split_zero) that the captured "
                                     "step replaces by one launch, so the classes sum to more than ms_per_step",
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            (tg, tu), = cpu_reference_sample(threads, repeats=1, warmup=0)
            line["cpu_baseline"] = {
                "value": images_per_sec_from_steps(tg, tu), "unit": "images/sec", "cores": threads, "kind": "port",
                "sample": "1 guided + 1 unguided denoising step of the same job (%.2f s + %.2f s), extrapolated to "
                          "26 guided + 24 unguided steps" % (tg, tu)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: libraries that write to fd 1 on their own (NCCL prints its version banner
    there on the first collective) are pointed at stderr for the duration of the run."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
