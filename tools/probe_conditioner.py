"""GPU probe: FrozenCLAPFLANEmbedder.encode_tokens (BERT-base + CLAP projection, T5 v1.1-large encoder) on this package's
kernels against the transformers modules the reference calls, same random weights: error and time per batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_conditioners import _hf_models, _state_dict, _projection
from ma3_b200 import conditioners as Cn, lib as L

B, T = int(os.environ.get("B", 8)), 77
bert, t5, proj = _hf_models(12, 24, seed=3)
emb = Cn.FrozenCLAPFLANEmbedder(state_dict=_state_dict(bert, t5, proj))
g = torch.Generator().manual_seed(1)
ori = torch.randint(0, 30522, (B, T), generator=g).cuda(); struct = torch.randint(0, 32128, (B, T), generator=g).cuda()
bert, t5 = bert.cuda(), t5.cuda(); pj = {k: v.cuda() for k, v in proj.items()}

def ref():
    with torch.no_grad():
        return torch.cat([_projection(bert(input_ids=ori).last_hidden_state, pj), t5(input_ids=struct).last_hidden_state], 1)

def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

n0 = L.launch_count(); out = emb.encode_tokens(ori, struct); nl = L.launch_count() - n0
r = ref()
print(f"B={B}: max err / max |ref| = {float((out - r).abs().max() / r.abs().max()):.4f}, cosine = "
      f"{float(torch.nn.functional.cosine_similarity(out.flatten(), r.flatten(), dim=0)):.6f}, {nl} launches")
print(f"ours {timed(lambda: emb.encode_tokens(ori, struct)):.2f} ms | transformers fp32 {timed(ref):.2f} ms | "
      f"transformers bf16 autocast {timed(lambda: torch.autocast('cuda', torch.bfloat16)(ref)()):.2f} ms")
