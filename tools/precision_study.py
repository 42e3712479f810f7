"""CPU study: which bf16 roundings of the B200 DiT path dominate the per-step guided-velocity error (north_star gate
1e-2) at full XL depth.  The oracle's forward is re-run with rounding applied at named points; everything else fp32.
Run: python tools/precision_study.py [model] [t_int]"""
import math, os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as Fn
from oracle import restated as O, weights as W

MODEL = sys.argv[1] if len(sys.argv) > 1 else "XL"
T_INT = int(sys.argv[2]) if len(sys.argv) > 2 else 83
CFGS = {"M": dict(in_channels=20, context_dim=1024, hidden_size=768, num_heads=32, depth=16),
        "XL": dict(in_channels=20, context_dim=1024, hidden_size=1152, num_heads=16, depth=28)}
cfg = CFGS[MODEL]
torch.set_num_threads(os.cpu_count())
sd = W.dit_state_dict(**cfg, seed=5)
c, uc, x0 = W.synthetic_inputs(prompts=1, latent_ch=20, T=312, L=154, Cd=1024)
x = torch.cat([x0, x0]); ctx = torch.cat([uc, c]); t = torch.full((2,), T_INT, dtype=torch.long)
heads = cfg["num_heads"]


def fwd(on, f16=()):
    """on: set of tags rounded to bf16 (or fp16 when the tag is also in f16)."""
    def r(v, tag):
        if tag not in on:
            return v
        return v.to(torch.float16 if tag in f16 else torch.bfloat16).float()
    D = cfg["hidden_size"]; hd = D // heads
    cos, sin = O.rope_table(hd, 1000)
    w = lambda k, tag="w": r(sd[k], tag)
    h = x.transpose(1, 2) @ sd["proj_in.weight"].t() + sd["proj_in.bias"]
    te = r(O.timestep_embedding(t), "cond_act")
    te = Fn.silu(te @ w("t_embedder.mlp.0.weight", "cond_w").t() + sd["t_embedder.mlp.0.bias"])
    te = r(te, "cond_act") @ w("t_embedder.mlp.2.weight", "cond_w").t() + sd["t_embedder.mlp.2.bias"]
    y = ctx
    pool = y.mean(1)
    cap = Fn.layer_norm(pool, (pool.shape[-1],), sd["cap_embedder.0.weight"], sd["cap_embedder.0.bias"], 1e-5)
    cap = r(cap, "cond_act") @ w("cap_embedder.1.weight", "cond_w").t() + sd["cap_embedder.1.bias"]
    sa = r(Fn.silu(te + cap), "cond_act")
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        mod = sa @ w(p + "adaLN_modulation.1.weight", "cond_w").t() + sd[p + "adaLN_modulation.1.bias"]
        sh1, sc1, g1, sh2, sc2, g2 = [m.unsqueeze(1) for m in mod.chunk(6, dim=1)]
        u = r(O.rmsnorm(h, sd[p + "attention_norm.weight"]) * (1 + sc1) + sh1, "u")
        yn = r(O.rmsnorm(y, torch.ones(())), "ctx")
        N, T, _ = u.shape
        a = p + "attention."
        q = (u @ w(a + "wq.weight").t()).view(N, T, heads, hd)
        k = (u @ w(a + "wk.weight").t()).view(N, T, heads, hd)
        v = r((u @ w(a + "wv.weight").t()).view(N, T, heads, hd), "v")
        q, k = r(O.apply_rope(q, cos, sin) * (math.log2(math.e) / math.sqrt(hd)), "qk"), r(O.apply_rope(k, cos, sin), "qk")
        qh = q.permute(0, 2, 1, 3)

        def attn(qh, kh, vh):
            s = qh @ kh.transpose(-1, -2)                     # log2 units
            pm = torch.exp2(s - s.amax(-1, keepdim=True))
            l = r(pm, "p").sum(-1, keepdim=True) if "p_sum_rounded" in on else pm.sum(-1, keepdim=True)
            return (r(pm, "p") @ vh) / l
        o = attn(qh, k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))
        yw = sd[p + "attention_y_norm.weight"][None, :]
        L = y.shape[1]
        yk = r((yn @ r(sd[a + "wk_y.weight"] * yw, "ctx").t()), "ctx").view(N, L, heads, hd).permute(0, 2, 1, 3)
        yv = r((yn @ r(sd[a + "wv_y.weight"] * yw, "ctx").t()), "ctx").view(N, L, heads, hd).permute(0, 2, 1, 3)
        oy = attn(qh, yk, yv) * torch.tanh(sd[a + "gate"]).view(1, heads, 1, 1)
        o = r((o + oy).permute(0, 2, 1, 3).reshape(N, T, D), "att")
        h = h + g1 * (o @ w(a + "wo.weight").t())
        z = r(O.rmsnorm(h, sd[p + "ffn_norm.weight"]) * (1 + sc2) + sh2, "u")
        f = p + "feed_forward."
        mid = r(Fn.silu(z @ w(f + "w1.weight").t()) * (z @ w(f + "w3.weight").t()), "mid")
        h = h + g2 * (mid @ w(f + "w2.weight").t())
    mod = sa @ w("final_layer.adaLN_modulation.1.weight", "cond_w").t() + sd["final_layer.adaLN_modulation.1.bias"]
    shift, scale = [m.unsqueeze(1) for m in mod.chunk(2, dim=1)]
    h = Fn.layer_norm(h, (D,), None, None, 1e-6) * (1 + scale) + shift
    out = (h @ sd["final_layer.linear.weight"].t() + sd["final_layer.linear.bias"]).transpose(1, 2)
    return out[:1] + 3.0 * (out[1:] - out[:1])


with torch.no_grad():
    ref = fwd(set())
    ALL = ["w", "cond_w", "cond_act", "u", "qk", "v", "p", "att", "mid", "ctx"]
    print(f"{MODEL} depth {cfg['depth']} t={T_INT}: guided velocity max-rel-err vs fp32")
    print(f"  everything bf16            : {O.max_rel_err(fwd(set(ALL)), ref):.5f}")
    for tag in ALL:
        print(f"  only {tag:9s} bf16        : {O.max_rel_err(fwd({tag}), ref):.5f}")
    for tag in ALL:
        print(f"  all but {tag:9s}          : {O.max_rel_err(fwd(set(ALL) - {tag}), ref):.5f}")
    print(f"  all, cond path fp32        : {O.max_rel_err(fwd(set(ALL) - {'cond_w', 'cond_act'}), ref):.5f}")
    print(f"  all, activations fp16      : {O.max_rel_err(fwd(set(ALL), f16={'u', 'qk', 'v', 'p', 'att', 'mid', 'ctx'}), ref):.5f}")
    print(f"  all, act fp16 + cond fp32  : {O.max_rel_err(fwd(set(ALL) - {'cond_w', 'cond_act'}, f16={'u', 'qk', 'v', 'p', 'att', 'mid', 'ctx'}), ref):.5f}")
