"""Restatement of the one torchdyn feature the reference uses: NeuralODE(solver='euler') on a fixed t_span.

torchdyn is an un-vendored, un-pinned dependency of the reference (call sites: ldm/models/diffusion/cfm1_audio.py:19,
74,80,102,109).  Semantics restated from torchdyn 1.0.x `odeint` -> `_fixed_odeint` with `Euler.step`:
t_span moved to x.device; x_{k+1} = x_k + dt * f(t_k, x_k); t advanced by dt; dt = t_span[k+1] - t; all states
stacked.  PARITY UNPINNED at this boundary: no copy of torchdyn exists on this box and the reference has no test
for it.
"""
import torch


class NeuralODE(torch.nn.Module):
    def __init__(self, vector_field, solver="euler", sensitivity="adjoint", atol=1e-4, rtol=1e-4, **kw):
        super().__init__()
        if solver != "euler":
            raise NotImplementedError("only the fixed-step Euler solver is restated")
        self.vf = vector_field

    def forward(self, x, t_span):
        t_span = t_span.to(x.device)
        t = t_span[0]
        dt = t_span[1] - t
        sol = [x]
        steps = 1
        while steps <= len(t_span) - 1:
            dx = self.vf(t, x, {})
            x = x + dt * dx
            t = t + dt
            sol.append(x)
            if steps < len(t_span) - 1:
                dt = t_span[steps + 1] - t
            steps += 1
        return t_span, torch.stack(sol)
