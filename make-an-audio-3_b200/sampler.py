"""CFM ODE sampling with the reference's operator surface.

`CFMSampler(model, num_timesteps)` mirrors ldm/models/diffusion/cfm1_audio_sampler.py:26-104 and the
`CFM.sample` / `CFM.sample_cfg` methods of ldm/models/diffusion/cfm1_audio.py:60-111: same arguments, same returns
`(x_final [B,C,T], traj [n_points,B,C,T])`.  `model` is either a B200 DiT from ma3_b200.dit or any wrapper exposing it
as `.model.diffusion_model` (the reference's CFM / LatentDiffusion_audio object with its `unet_config.target` swapped).

The fixed-step Euler integrator of torchdyn (un-vendored dependency of the reference) is restated here:
x_{k+1} = x_k + dt_k f(t_k, x_k), t advanced as t + dt in fp32, integer timestep = trunc(1000 t)
(cfm1_audio.py:103,156).  What changes on B200: the cond/uncond pair is one batched pass, context K/V and all adaLN
modulations are computed once, the guidance combine and the Euler update are fused into the final-layer kernel, and
the whole step loop is replayed from a CUDA graph (no host synchronisation inside the loop).
"""
import ctypes
import os
import torch

from . import lib as L
from . import ops
from .dit import TxtFlagLargeDiT


# kernel launches replayed from captured CUDA graphs (ma3_launch_count only sees launches enqueued through the C ABI)
GRAPH_REPLAY_LAUNCHES = 0


def euler_schedule(n_points=None, t_start=None):
    """Integer timesteps and step sizes exactly as the reference's loop produces them (fp32 arithmetic on t)."""
    ts = torch.linspace(0, 1, 25 if n_points is None else n_points)
    if t_start is not None:
        ts = ts[t_start:]
    if len(ts) < 2:
        raise ValueError("need at least two time points")
    t = ts[0]
    dt = ts[1] - t
    ints, dts = [], []
    for k in range(1, len(ts)):
        ints.append(int((t * 1000).long()))
        dts.append(float(dt))
        t = t + dt
        if k < len(ts) - 1:
            dt = ts[k + 1] - t
    return ints, dts


def _find_dit(model):
    if isinstance(model, TxtFlagLargeDiT):
        return model
    inner = getattr(getattr(model, "model", None), "diffusion_model", None)
    if isinstance(inner, TxtFlagLargeDiT):
        return inner
    raise L.Ma3Error("CFMSampler needs a ma3_b200 DiT (directly or as model.model.diffusion_model)")


def _as_context(cond):
    """Mirror DiffusionWrapper.forward's crossattn handling (ddpm.py:1413-1420): lists are concatenated on dim 1."""
    if isinstance(cond, dict):
        cond = cond.get("c_crossattn", next(iter(cond.values())))
    if isinstance(cond, (list, tuple)):
        cond = torch.cat(list(cond), 1)
    return cond


class CFMSampler:
    def __init__(self, model, num_timesteps=1000, schedule="linear", use_graph=True):
        self.model = model
        self.dit = _find_dit(model)
        self.num_timesteps = num_timesteps
        self.use_graph = use_graph
        self._graphs = {}

    # ---------------------------------------------------------------- reference surface
    def _default_shape(self, batch_size):
        m = self.model
        mel_dim = getattr(m, "mel_dim", self.dit.in_channels)
        mel_length = getattr(m, "mel_length", 256)
        if getattr(m, "channels", 0) > 0:
            raise L.Ma3Error("4-D latents (channels > 0) are not produced by the 1-D Next-DiT path")
        return (batch_size, mel_dim, mel_length)

    @staticmethod
    def _slice(cond, batch_size):
        if cond is None:
            return None
        if isinstance(cond, dict):
            return {k: (v[:batch_size] if not isinstance(v, list) else [x[:batch_size] for x in v]) for k, v in cond.items()}
        return [c[:batch_size] for c in cond] if isinstance(cond, list) else cond[:batch_size]

    @torch.no_grad()
    def sample(self, cond, batch_size=16, timesteps=None, shape=None, x_latent=None, t_start=None, **kwargs):
        return self._run(cond, None, None, batch_size, timesteps, shape, x_latent, t_start)

    @torch.no_grad()
    def sample_cfg(self, cond, unconditional_guidance_scale, unconditional_conditioning, batch_size=16, timesteps=None,
                   shape=None, x_latent=None, t_start=None, **kwargs):
        return self._run(cond, float(unconditional_guidance_scale), unconditional_conditioning, batch_size, timesteps,
                         shape, x_latent, t_start)

    # ---------------------------------------------------------------- engine
    def _run(self, cond, scale, uncond, batch_size, timesteps, shape, x_latent, t_start):
        dit = self.dit
        dev = dit.proj_in.weight.device
        if shape is None:
            shape = self._default_shape(batch_size)
        c = _as_context(self._slice(cond, batch_size)).to(dev).float()
        cfg = scale is not None
        if cfg:
            uc = _as_context(uncond).to(dev).float()
            ctx = torch.cat([uc, c])          # batch order [uncond, cond] (cfm1_audio.py:158)
        else:
            ctx = c
        x0 = torch.randn(shape, device=dev) if x_latent is None else x_latent.to(dev).float()
        B, C, T = x0.shape
        N = ctx.shape[0]
        if N != (2 * B if cfg else B):
            raise ValueError(f"conditioning batch {ctx.shape[0]} does not match latent batch {B}")
        ints, dts = euler_schedule(timesteps, t_start)
        S = len(ints)

        L.auto_pdl(N * T)   # kernel attribute, baked into the captured nodes: chosen per workload before anything is launched
        cbuf = dit.prepare_context(ctx)
        work = dit._workspace(N, T)
        # The plan (buffers + captured graph) is valid for these shapes and this schedule, and only as long as none of
        # the device buffers the captured launches point at has been re-allocated: dit.generation counts those
        # (load_state_dict / .to() / a freqs_cis table of another shape / a new context or workspace shape), and the
        # plan holds a reference to every captured buffer so that nothing is freed while the graph is alive.
        key = (cfg, B, C, T, tuple(ctx.shape), tuple(ints), scale, dit.generation)
        st = self._graphs.get(key)
        if st is None:
            st = {"traj": torch.empty(S + 1, B, C, T, device=dev, dtype=torch.float32),
                  "cond": dit.cond_buffers(S, N, T),
                  "v": torch.empty(N, C, T, device=dev, dtype=torch.float32), "graph": None,
                  "keep": (dit._packed, cbuf, work)}
            self._graphs = {key: st}          # keep one plan: buffers are large
        dit.prepare_timesteps(torch.tensor(ints, dtype=torch.int64), out=st["cond"])   # step-invariant, outside the loop
        st["traj"][0].copy_(x0)

        def loop():
            for k in range(S):
                self._step(st, k, ints[k], dts[k], scale, cfg, N, T)

        if not self.use_graph:
            loop()
        elif st["graph"] is None:
            loop()                             # eager pass: allocates workspaces, configures kernels
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            cap = torch.cuda.Stream()
            # The fp32 residual stream (read twice and reduced into twice per block) is given an L2 access-policy
            # window on the capture stream, so every captured launch keeps its lines resident while ~190 MB of
            # operands and weights per block stream through the 126 MB cache: +0.9 .. 1.3 % on the XL step in
            # same-box A/B runs (windows on u / mid / att measured neutral).  MA3_L2_PERSIST=0 switches it off,
            # =<workspace buffer name> moves it.
            which = os.environ.get("MA3_L2_PERSIST", "h")
            if which not in ("", "0"):
                self._l2_window(work, "h" if which == "1" else which, cap)
            with torch.cuda.graph(g, stream=cap):
                loop()
            self._l2_window(None, None, cap)   # the window is baked into the captured kernel nodes; clear the stream's
            st["graph_launches"] = L.launch_count() - n0
            st["graph"] = g                    # the eager pass above already produced this call's trajectory
        else:
            global GRAPH_REPLAY_LAUNCHES
            st["graph"].replay()
            GRAPH_REPLAY_LAUNCHES += st["graph_launches"]
        traj = st["traj"].clone()
        return traj[-1], traj

    def close(self):
        """Drop the captured plan and give the persisting-L2 set-aside back to the device."""
        self._graphs = {}
        try:
            L.load().ma3_l2_persist_release()
        except Exception:  # noqa: BLE001
            pass

    @staticmethod
    def _l2_window(work, name, stream):
        """Best effort (a ~1 % optimisation must never abort sampling): put an L2 access-policy window on `stream` over
        workspace buffer `name`, or clear it when work is None.  The persisting set-aside it needs is a device-wide
        limit and stays sized to the window while the plan lives (include/ma3_b200.h, ma3_l2_persist)."""
        import warnings
        try:
            buf = None
            if work is not None:
                buf = getattr(work, name, None)
                if not torch.is_tensor(buf):
                    warnings.warn(f"MA3_L2_PERSIST={name!r} is not a workspace buffer; L2 window not set")
                    return
            rc = L.load().ma3_l2_persist(ctypes.c_void_p(buf.data_ptr() if buf is not None else 0),
                                         buf.numel() * buf.element_size() if buf is not None else 0,
                                         ctypes.c_void_p(stream.cuda_stream))
            if rc != 0:
                warnings.warn("ma3_l2_persist failed (continuing without the L2 window): "
                              + L.load().ma3_last_error().decode("utf-8", "replace"))
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"L2 window not set: {e!r}")

    def _step(self, st, k, t_int, dt, scale, cfg, N, T):
        dit = self.dit
        p = dit._packed
        D = dit.hidden_size
        mod = st["cond"]["mod"][k]
        w = dit.run_blocks(st["traj"][k], st["cond"], k, t_ints=[t_int] * N if dit.num_experts else None)
        off = p["final_off"]
        if cfg:
            ops.final_layer_cfg_euler(w.h, mod, off, off + D, p["final_w"], p["final_b"], N, T, scale, dt,
                                      st["traj"][k], st["traj"][k + 1])
        else:
            ops.final_layer(w.h, mod, off, off + D, p["final_w"], p["final_b"], N, T, st["v"])
            ops.cfg_euler_update(st["v"], st["traj"][k], st["traj"][k + 1], dt, 0.0, cfg=False)
