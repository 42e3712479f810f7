#!/usr/bin/env python
"""Benchmark of the Make-An-Audio-3 sampling path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus 1 --steps K --warmup W                 # this framework
    torchrun ... bench.py --gpus N --steps K --warmup W           # one rank per GPU, prompts sharded (weak scaling)
    python bench.py --impl reference --steps K --warmup W         # the reference arithmetic on the host CPU

    python bench.py --config N                                    # BASELINE.json configs[N-1] at its stated shape

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


# ------------------------------------------------------------------------------------------------ reference arms
# BASELINE.json configs[0..4] at their stated shapes: model, prompts per GPU, T, L
CONFIG_PRESETS = {1: ("M", 1, 312, 154), 2: ("XL", 8, 312, 154), 3: ("XXL", 1, 936, 154), 4: ("M", 16, 312, 77),
                  5: ("MOE", 1, 256, 40)}


def _reference_stack(model, device):
    """The reference's OWN modules (CFM -> DiffusionWrapper -> Next-DiT, AutoencoderKL, BigVGAN) built through its
    instantiate_from_config from oracle/_ref (or /root/reference), loaded with the same seeded weights the B200 path
    is tested with.  Returns (cfm, bigvgan) or None when no copy of the reference is present."""
    from oracle import ref_loader as R, weights as W
    from ma3_b200.pipeline import MODEL_CONFIGS, VAE_DDCONFIG
    if not R.available():
        return None
    cfg = dict(MODEL_CONFIGS[model])
    ne = cfg.get("num_experts", 0)
    wcfg = {k: v for k, v in cfg.items() if k not in ("max_len", "num_experts")}
    dsd = W.dit_state_dict(**wcfg, video=ne > 0, num_experts=ne, seed=0)
    vsd = W.vae_decoder_state_dict(VAE_DDCONFIG, 20)
    cfm = R.build_cfm(cfg, VAE_DDCONFIG, 20, dsd, vsd, video=ne > 0).to(device)
    voc = R.build_bigvgan(W.bigvgan_state_dict(W.BIGVGAN_LARGE_256X), W.BIGVGAN_LARGE_256X).to(device)
    return cfm, voc


def _reference_clip(stack, c, uc, x0):
    """cfm1_audio.py:89-111 -> ddpm_audio.py:358-371 -> vocoder/bigvgan/models.py:183-205, exactly the calls of
    scripts/txt2audio_for_2cap_flow.py:170-188 with the conditioner's output replaced by synthetic embeddings."""
    cfm, voc = stack
    z, _ = cfm.sample_cfg(c, GUIDANCE, uc, x0.shape[0], timesteps=N_POINTS, x_latent=x0)
    mel = cfm.decode_first_stage(z)
    return voc(mel)


def _port_clip(model, c, uc, x0):
    """Fallback when no copy of the reference is present: the oracle port (oracle/restated.py), whole clip."""
    from ma3_b200.pipeline import MODEL_CONFIGS, VAE_DDCONFIG
    from oracle import restated as O, weights as W
    cfg = dict(MODEL_CONFIGS[model])
    ne = cfg.get("num_experts", 0)
    wcfg = {k: v for k, v in cfg.items() if k not in ("max_len", "num_experts")}
    dsd = W.dit_state_dict(**wcfg, video=ne > 0, num_experts=ne, seed=0)
    vsd = W.vae_decoder_state_dict(VAE_DDCONFIG, 20)
    bsd = W.bigvgan_state_dict(W.BIGVGAN_LARGE_256X)
    vel = lambda x, t, ctx: O.dit_forward(dsd, x, t, ctx, heads=cfg["num_heads"], video=ne > 0, num_experts=ne)

    def run():
        z, _, _ = O.sample_cfg(vel, x0, c, uc, GUIDANCE, n_points=N_POINTS)
        return O.bigvgan_forward(bsd, O.vae_decode(vsd, z, VAE_DDCONFIG), W.BIGVGAN_LARGE_256X)
    return run


class CpuReference:
    """The reference's CPU path on the box's host cores: ONE WHOLE clip per step (1 prompt of the workload: 24 Euler
    steps at CFG batch 2, VAE decode, BigVGAN on all mel frames) -- nothing is extrapolated inside a clip; audio-s/s on
    the CPU is independent of the prompt count, which is the only scaling to the 8-prompt workload."""

    def __init__(self, model, T, L, threads=None):
        from ma3_b200.pipeline import MODEL_CONFIGS
        from oracle import weights as W
        os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.model, self.T, self.L = model, T, L
        Cd = MODEL_CONFIGS[model]["context_dim"]
        self.c, self.uc, self.x0 = W.synthetic_inputs(prompts=1, latent_ch=20, T=T, L=L, Cd=Cd)
        stack = _reference_stack(model, "cpu")
        if stack is not None:
            self.kind = "reference"
            self.fn = lambda: _reference_clip(stack, self.c, self.uc, self.x0)
        else:
            self.kind = "port"
            self.fn = _port_clip(model, self.c, self.uc, self.x0)

    def step(self):
        with torch.no_grad():
            t0 = time.perf_counter()
            wav = self.fn()
            dt = time.perf_counter() - t0
        assert wav.shape[-1] == 2 * self.T * HOP
        return dt

    def describe(self, secs):
        what = ("the reference's own CFM.sample_cfg -> decode_first_stage -> BigVGAN (unmodified sources under "
                "oracle/_ref, fp32, torch CPU)"
                if self.kind == "reference" else "oracle port (oracle/restated.py, fp32, torch CPU)")
        return (f"{what}: 1 prompt of {self.model} (T={self.T}, L={self.L}), one WHOLE clip per step = 24 Euler steps at "
                f"CFG batch 2 + full VAE decode + BigVGAN large-256x on all {2 * self.T} mel frames; "
                f"{secs:.1f} s per {clip_seconds(self.T):.3f} s clip")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    ref = CpuReference(args.model, args.T, args.L)
    warm = min(args.warmup, 1)        # a CPU clip takes ~10-30 s: one untimed clip warms caches and allocators
    for _ in range(warm):
        ref.step()
    secs = [ref.step() for _ in range(max(1, args.steps))]
    per_clip = statistics.mean(secs)
    value = clip_seconds(args.T) / per_clip
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": warm, "ms_per_step": 1e3 * per_clip,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.threads, "kind": ref.kind,
                             "sample": ref.describe(per_clip)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def library_baseline(model, B, T, L, dev, reps=2):
    """SURVEY.md section 8(d) / BASELINE.md section 3: the UNMODIFIED reference modules on the same B200 (PyTorch's
    library kernels: cuBLAS, cuDNN, SDPA), same weights / inputs / batch as the benchmarked step -- (a) fp32 as shipped,
    (b) bf16 autocast with flag_large_dit_moe.is_flash_attn = False (the only bf16 mode of the reference that runs,
    flag_large_dit_moe.py:364,382-388).  This, not the CPU arm, is what the hand-written kernels have to beat."""
    from ma3_b200.pipeline import MODEL_CONFIGS
    from oracle import ref_loader as R, weights as W
    if not R.available():
        return {"unavailable": "no copy of the reference on this box (oracle/_ref is written by build())"}
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    out = {"what": "reference CFM.sample_cfg -> decode_first_stage -> BigVGAN on this GPU through PyTorch library "
                   f"kernels, {B} prompts, same weights and inputs"}
    try:
        stack = _reference_stack(model, dev)
        Cd = MODEL_CONFIGS[model]["context_dim"]
        c, uc, x0 = W.synthetic_inputs(prompts=B, latent_ch=20, T=T, L=L, Cd=Cd)
        c, uc, x0 = c.to(dev), uc.to(dev), x0.to(dev)
        import ldm.modules.diffusionmodules.flag_large_dit_moe as moe_mod

        def timed(ctx):
            ms = []
            for i in range(reps + 1):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                with torch.no_grad(), ctx():
                    _reference_clip(stack, c, uc, x0)
                e1.record()
                torch.cuda.synchronize()
                if i:
                    ms.append(e0.elapsed_time(e1))
            return statistics.mean(ms)

        import contextlib
        ms32 = timed(contextlib.nullcontext)
        out["fp32_as_shipped"] = {"ms_per_step": round(ms32, 2), "value": round(B * clip_seconds(T) / (ms32 / 1e3), 2),
                                  "unit": UNIT}
        flag = moe_mod.is_flash_attn
        moe_mod.is_flash_attn = False
        try:
            ms16 = timed(lambda: torch.autocast("cuda", dtype=torch.bfloat16))
            out["bf16_autocast"] = {"ms_per_step": round(ms16, 2), "value": round(B * clip_seconds(T) / (ms16 / 1e3), 2),
                                    "unit": UNIT}
        finally:
            moe_mod.is_flash_attn = flag
        del stack
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001  (the repo's own numbers stand without it)
        out["error"] = repr(e)[:300]
    return out


def workload_config(args, world):
    T, L = args.T, args.L
    return {"workload": f"{CONFIG_NAMES.get(args.model, args.model)} ({args.model}): {args.prompts} prompts/GPU x {world} GPU, "
                        f"{clip_seconds(T):.3f} s clips (T={T}, L={L}), 25 CFM points = 24 Euler steps, CFG 3.0 "
                        f"(DiT batch {2 * args.prompts}/GPU), VAE decode + BigVGAN large-256x (assumed h, SURVEY 8(d))",
            "prompts_per_gpu": args.prompts, "global_prompts": args.prompts * world, "parallelism": f"dp{world} (prompts)",
            "T_latent": T, "L_context": L,
            "precision": "DiT+VAE bf16 operands / fp32 accumulate+residual; vocoder fp16 / fp32 accumulate",
            "l2": "no explicit flush: per-step working set (1.5 GB bf16 DiT weights + >2 GB activations) >> 126 MB L2",
            "weights": "random init"}


CONFIG_NAMES = {"M": "txt2audio-cfm-cfg", "XL": "txt2audio-cfm-cfg-XL", "XXL": "txt2audio-cfm-cfg-XXL",
                "MOE": "video2audio-cfm-cfg-moe"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help="BASELINE.json configs[N-1] at its stated shape (sets --model/--prompts/--T/--L); 0 = use the flags")
    ap.add_argument("--model", default="XL", choices=["M", "XL", "XXL", "MOE"])
    ap.add_argument("--prompts", type=int, default=8, help="prompts per GPU")
    ap.add_argument("--T", type=int, default=0, help="latent frames (default 312 = 10 s; 936 = 30 s; MOE: 256)")
    ap.add_argument("--L", type=int, default=0, help="context tokens (default 154; music 77; MOE: 40)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lib-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.config:
        args.model, args.prompts, args.T, args.L = CONFIG_PRESETS[args.config]
    args.T = args.T or (256 if args.model == "MOE" else T_LATENT)
    args.L = args.L or (40 if args.model == "MOE" else L_CTX)
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
    L, T = args.L, args.T
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
        # DRAM traffic per launch of that family from this round's ncu --set full captures (tools/profile_r02.sh ->
        # tools/ncu_summarise.py --traffic): for the tap-GEMM family the launch-weighted mean over its DiT members
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            fams = {"tap_gemm": ["tap_gemm", "rownorm"]}.get(name, [name])
            caps = [c for f in fams if f in tj for c in tj[f]["captures"]]
            if caps:
                roofline["traffic"] = sum(c["dram_bytes"] for c in caps) / len(caps)
                roofline["traffic_by_kernel"] = {c["kernel"][:60] + f" ({c['gpu_time_us']:.0f} us)": c["dram_bytes"] for c in caps}
                roofline["traffic_source"] = "; ".join(tj[f]["source"] for f in fams if f in tj)
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

    lib_base = None
    if rank == 0 and world == 1 and not args.no_lib_baseline:
        pipe.sampler.close()
        lib_base = library_baseline(args.model, B, T, L, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = CpuReference(args.model, T, L)
            secs = ref.step()          # one whole clip (about 10-30 s of CPU work)
            cpu = {"value": clip_seconds(T) / secs, "unit": UNIT, "cores": ref.threads, "kind": ref.kind,
                   "sample": ref.describe(secs)}
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "torch_b200_baseline": lib_base, "stage_ms": stage_ms,
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
