#!/usr/bin/env python
"""Benchmark of the Make-An-Audio-3 sampling path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus 1 --steps K --warmup W                 # this framework
    torchrun ... bench.py --gpus N --steps K --warmup W           # one rank per GPU, prompts sharded (weak scaling)
    python bench.py --impl reference --steps K --warmup W         # the reference arithmetic on the host CPU

Workload = BASELINE.json configs[1]: txt2audio-cfm-cfg-XL (Next-DiT D=1152, 16 heads, depth 28), 64 prompts over
8 GPUs = 8 prompts per GPU, 10 s clips (T=312 latent frames -> 159 744 samples), 25 CFM points = 24 Euler steps,
CFG 3.0 (DiT batch 16), then VAE decode and BigVGAN (large-256x layout) for the 8 clips.  One "step" = one such batch
per GPU.  metric = generated audio seconds per wall second, whole job.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "generated audio-sec/sec (10 s clips, 25 CFM steps, CFG=3)"
UNIT = "audio-s/s"
SR, HOP = 16000, 256
T_LATENT, L_CTX, CD = 312, 154, 1024
N_POINTS, GUIDANCE = 25, 3.0


def clip_seconds(T=T_LATENT):
    return 2 * T * HOP / SR  # 9.984 s


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.mask, self.stop_flag, self.max_mhz = [], 0, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=2)
        reasons = [n for b, n in self.REASONS.items() if self.mask & b]
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(model="XL", n_dit_steps=2, mel_frames=156, threads=None):
    """Bounded sample of the workload on the host CPU through the oracle port of the reference arithmetic
    (oracle/restated.py, fp32, torch CPU): n_dit_steps Euler steps of one prompt (CFG batch 2) scaled to 24,
    the full VAE decode of one clip, BigVGAN on mel_frames of the 624 mel frames scaled linearly.
    Returns (audio-s/s for one clip, description, threads)."""
    from ma3_b200.pipeline import MODEL_CONFIGS, VAE_DDCONFIG
    from oracle import restated as O, weights as W
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(MODEL_CONFIGS[model])
    cfg.pop("max_len")
    video = "num_experts" in cfg
    dsd = W.dit_state_dict(**cfg, video=video, seed=0)
    vsd = W.vae_decoder_state_dict(VAE_DDCONFIG, 20)
    h = W.BIGVGAN_LARGE_256X
    bsd = W.bigvgan_state_dict(h)
    c, uc, x0 = W.synthetic_inputs(prompts=1, latent_ch=20, T=T_LATENT, L=L_CTX, Cd=cfg["context_dim"])
    ints, dts = O.timestep_ints(N_POINTS)
    with torch.no_grad():
        x = x0
        t0 = time.perf_counter()
        for ti, dt in list(zip(ints, dts))[:n_dit_steps]:
            t = torch.full((2,), ti, dtype=torch.long)
            v = O.dit_forward(dsd, torch.cat([x, x]), t, torch.cat([uc, c]), heads=cfg["num_heads"], video=video,
                              num_experts=cfg.get("num_experts", 0))
            x = x + dt * (v[:1] + GUIDANCE * (v[1:] - v[:1]))
        t_dit = (time.perf_counter() - t0) * (len(ints) / n_dit_steps)
        t0 = time.perf_counter()
        mel = O.vae_decode(vsd, x, VAE_DDCONFIG)
        t_vae = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.bigvgan_forward(bsd, mel[..., :mel_frames], h)
        t_voc = (time.perf_counter() - t0) * (mel.shape[-1] / mel_frames)
    total = t_dit + t_vae + t_voc
    desc = (f"oracle port (torch fp32) of the reference arithmetic, 1 prompt of {model}: {n_dit_steps}/24 Euler steps "
            f"(CFG batch 2) x{len(ints) // n_dit_steps}, full VAE decode, BigVGAN on {mel_frames}/624 mel frames "
            f"x{624 / mel_frames:.0f}; est. {t_dit:.1f}+{t_vae:.1f}+{t_voc:.1f} s per 9.984 s clip")
    return clip_seconds() / total, desc, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    desc, threads = "", 0
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, desc, threads = cpu_sample(args.model)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * clip_seconds() / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"txt2audio-cfm-cfg-{args.model}: {args.prompts} prompts/GPU x {world} GPU, 10 s clips (T=312, "
                        f"L=154), 25 CFM points = 24 Euler steps, CFG 3.0 (DiT batch {2 * args.prompts}/GPU), "
                        "VAE decode + BigVGAN large-256x (assumed h, SURVEY 8(d))",
            "prompts_per_gpu": args.prompts, "global_prompts": args.prompts * world, "parallelism": f"dp{world} (prompts)",
            "precision": "DiT+VAE bf16 operands / fp32 accumulate+residual; vocoder fp16 / fp32 accumulate",
            "l2": "no explicit flush: per-step working set (1.5 GB bf16 DiT weights + >2 GB activations) >> 126 MB L2",
            "weights": "random init"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="XL", choices=["M", "XL", "XXL", "MOE"])
    ap.add_argument("--prompts", type=int, default=8, help="prompts per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from ma3_b200 import lib, ops, sampler as S
    from ma3_b200.pipeline import MODEL_CONFIGS, build_random_pipeline, gather_waveforms

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.require_device()

    cfg = MODEL_CONFIGS[args.model]
    Cd = cfg["context_dim"]
    L = 40 if args.model == "MOE" else L_CTX
    T = 256 if args.model == "MOE" else T_LATENT
    B = args.prompts
    pipe = build_random_pipeline(args.model, vocoder_h=dict(BIGVGAN_H), seed=rank, device=dev,
                                 use_graph=not args.no_graph)
    g = torch.Generator().manual_seed(1234 + rank)
    cond_h = torch.randn(B, L, Cd, generator=g).pin_memory()
    unc_h = torch.randn(1, L, Cd, generator=torch.Generator().manual_seed(4321)).expand(B, L, Cd).contiguous().pin_memory()
    x0_h = torch.randn(B, 20, T, generator=torch.Generator().manual_seed(2024 + rank)).pin_memory()
    cond, unc, x0 = cond_h.to(dev), unc_h.to(dev), x0_h.to(dev)
    samples = 2 * T * HOP
    wav_h = torch.empty(B * world, samples).pin_memory()

    def step_device():
        wav = pipe.generate(cond, unc, x0, scale=GUIDANCE, timesteps=N_POINTS)
        return gather_waveforms(wav)

    def step_e2e():
        c = cond_h.to(dev, non_blocking=True)
        u = unc_h.to(dev, non_blocking=True)
        x = x0_h.to(dev, non_blocking=True)
        wav = gather_waveforms(pipe.generate(c, u, x, scale=GUIDANCE, timesteps=N_POINTS))
        wav_h.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller needs the audio before the next request
        return wav_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0, g0 = lib.launch_count(), S.GRAPH_REPLAY_LAUNCHES
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = (lib.launch_count() - n0) + (S.GRAPH_REPLAY_LAUNCHES - g0)
        return float(ms) / k, launches // k

    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local)
    clocks.start()
    ms_step, launches = timed(step_device, args.steps)
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop()
    audio_s = B * world * clip_seconds(T)
    value = audio_s / (ms_step / 1e3)
    e2e_value = audio_s / (ms_e2e / 1e3)
    h2d = cond_h.numel() * 4 + unc_h.numel() * 4 + x0_h.numel() * 4
    d2h = B * world * samples * 4   # rank 0 copies the gathered waveforms of all ranks to the host

    # per-stage device time (sampler graph / VAE decode / vocoder), same inputs, CUDA events
    stage_ms = None
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = [0.0, 0.0, 0.0]
        reps = 3
        for _ in range(reps):
            torch.cuda.synchronize()
            ev[0].record()
            z, _ = pipe.sample_cfg(cond, GUIDANCE, unc, B, timesteps=N_POINTS, x_latent=x0)
            ev[1].record()
            mel = pipe.decode_first_stage(z)
            ev[2].record()
            pipe.vocoder.vocode_tensor(mel)
            ev[3].record()
            torch.cuda.synchronize()
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) / reps
        stage_ms = {"sample_cfg_24_steps": round(acc[0], 3), "vae_decode": round(acc[1], 3), "bigvgan": round(acc[2], 3)}

    # live roofline of the dominant kernel: CUDA events around every launch on the launching stream.  With graphs the
    # plans are re-captured with the events as graph nodes and replayed once, so each interval is device time of one
    # kernel with nothing from the host in it; --no-graph brackets the eager launches instead.
    roofline, breakdown, gemm_shapes = None, None, None
    if rank == 0:
        ops.PROFILE = []
        if args.no_graph:
            torch.cuda.synchronize()
            pipe.generate(cond, unc, x0, scale=GUIDANCE, timesteps=N_POINTS)
        else:
            ops.PROFILE_CAPTURED_ONLY = True
            pipe.sampler._graphs, pipe._tail = {}, {}
            pipe.generate(cond, unc, x0, scale=GUIDANCE, timesteps=N_POINTS)   # eager pass + capture (events -> nodes)
            torch.cuda.synchronize()
            pipe.generate(cond, unc, x0, scale=GUIDANCE, timesteps=N_POINTS)   # replay: the events get their times
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        ops.PROFILE_CAPTURED_ONLY = False
        agg = {}
        for tag, a, b, work in prof:
            fam = tag.split("/")[0]
            d = agg.setdefault(fam, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += work
            d[2] += 1
        tot_ms = sum(d[0] for d in agg.values())
        # the GEMM family by problem shape (in-step device time, launches, achieved TFLOP/s), largest first
        shapes = {}
        for tag, a, b, work in prof:
            if tag.startswith("tap_gemm/"):
                d = shapes.setdefault(tag[len("tap_gemm/"):], [0.0, 0.0, 0])
                d[0] += a.elapsed_time(b)
                d[1] += work
                d[2] += 1
        gemm_shapes = {k: {"ms": round(v[0], 3), "launches": v[2], "us_per_launch": round(1e3 * v[0] / v[2], 1),
                           "TFLOPs": round(v[1] / (v[0] * 1e-3) / 1e12, 1)}
                       for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:24]}
        breakdown = {k: {"ms": round(v[0], 3), "launches": v[2], "share": round(v[0] / tot_ms, 4)} for k, v in
                     sorted(agg.items(), key=lambda kv: -kv[1][0])}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        top = max(agg.items(), key=lambda kv: kv[1][0])
        name, (ms, work, n) = top[0], top[1]
        if name in ("tap_gemm", "attention"):
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            ach = work / (ms * 1e-3) / 1e12
            roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (sustained)",
                        "launches": n, "avg_launch_us": 1e3 * ms / n, "flops_per_launch": work / n}
        else:
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = work / (ms * 1e-3) / 1e9
            roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                        "launches": n, "avg_launch_us": 1e3 * ms / n, "bytes_per_launch": work / n}
        # DRAM traffic per launch of that family from the committed ncu --set full capture, when there is one
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
            if tr:
                roofline["traffic"] = tr["bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
        except Exception:
            pass
        # secondary: the HBM-bound vocoder activation kernel the north_star singles out
        if "act1d" in agg:
            ms, work, n = agg["act1d"]
            breakdown["act1d"]["achieved_GBps"] = round(work / (ms * 1e-3) / 1e9, 1)
            breakdown["act1d"]["frac_of_hbm_peak"] = round(work / (ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), 4)
        for fam in ("tap_gemm", "attention"):
            if fam in agg:
                ms, work, n = agg[fam]
                breakdown[fam]["achieved_TFLOPs"] = round(work / (ms * 1e-3) / 1e12, 1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, desc, threads = cpu_sample(args.model)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "stage_ms": stage_ms,
                "kernel_breakdown": breakdown, "gemm_shapes": gemm_shapes}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# assumed BigVGAN hyper-parameters of the benchmark (the reference repo does not ship them)
BIGVGAN_H = dict(resblock="1", num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
                 upsample_initial_channel=1536, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True,
                 sampling_rate=16000, hop_size=256)

if __name__ == "__main__":
    main()
