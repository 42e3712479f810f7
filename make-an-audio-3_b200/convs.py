"""Conv1d / ConvTranspose1d on channels-last 16-bit activations, expressed as tap-GEMMs (see include/ma3_b200.h).

Weights are repacked once: Conv1d [C_out, C_in, k] -> [k * C_out, C_in_pad] (one K-major matrix per tap);
ConvTranspose1d [C_in, C_out, k] -> per output phase r the taps j == (r + pad) mod stride.
"""
import torch

from . import ops


def cpad(c):
    """Channel counts are padded to a multiple of 16 (UMMA K granularity); pad channels are kept exactly zero."""
    return (c + 15) // 16 * 16


class PackedConv:
    def __init__(self, weight, bias, *, dilation=1, padding=None, dtype=torch.bfloat16, device="cuda"):
        cout, cin, k = weight.shape
        self.cout, self.cin, self.k, self.dil = cout, cin, k, dilation
        self.pad = (k * dilation - dilation) // 2 if padding is None else padding
        self.cin_pad = cpad(cin)
        w = torch.zeros(k, cout, self.cin_pad, dtype=torch.float32)
        w[:, :, :cin] = weight.detach().float().cpu().permute(2, 0, 1)
        self.w = w.view(k * cout, self.cin_pad).to(device=device, dtype=dtype).contiguous()
        self.bias = bias.detach().float().to(device).contiguous() if bias is not None else None
        self.taps = [(j * dilation - self.pad, j * cout) for j in range(k)]

    def __call__(self, x, out, *, res=None, alpha=1.0, accumulate=False, act=0, out_ld=None):
        """x [B, T, cin_pad] -> out [B, T, >= cout] (same T: 'same' padding)."""
        B, T, C = x.shape
        assert C == self.cin_pad and x.is_contiguous(), (C, self.cin_pad)
        ld = out_ld if out_ld is not None else out.shape[-1]
        ops.gemm(x, self.w, M=T, N=self.cout, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=self.w.shape[0],
                 taps=self.taps, out=out, out_ld=ld, out_batch_stride=T * ld, bias=self.bias, res=res,
                 res_ld=res.shape[-1] if res is not None else None,
                 res_batch_stride=T * res.shape[-1] if res is not None else 0, alpha=alpha, accumulate=accumulate,
                 act=act)
        return out


class PackedConvTranspose:
    """ConvTranspose1d(stride s, kernel k, padding p): out[s*q + r] = sum_{j == (r+p) mod s} W_j^T x[q + (r+p-j)/s]."""

    def __init__(self, weight, bias, *, stride, padding, dtype=torch.bfloat16, device="cuda"):
        cin, cout, k = weight.shape
        self.cin, self.cout, self.k, self.s, self.p = cin, cout, k, stride, padding
        self.cin_pad = cpad(cin)
        w = torch.zeros(k, cout, self.cin_pad, dtype=torch.float32)
        w[:, :, :cin] = weight.detach().float().cpu().permute(2, 1, 0)
        self.w = w.view(k * cout, self.cin_pad).to(device=device, dtype=dtype).contiguous()
        self.bias = bias.detach().float().to(device).contiguous() if bias is not None else None
        self.phases = []
        for r in range(stride):
            taps = [((r + padding - j) // stride, j * cout) for j in range(k) if (r + padding - j) % stride == 0]
            self.phases.append(taps)

    def out_len(self, T):
        return (T - 1) * self.s - 2 * self.p + self.k

    def __call__(self, x, out):
        """x [B, T, cin_pad] -> out [B, T_out, >= cout]; every phase writes rows r, r+s, r+2s, ..."""
        B, T, C = x.shape
        To, ld = out.shape[1], out.shape[2]
        assert C == self.cin_pad and To == self.out_len(T)
        for r, taps in enumerate(self.phases):
            rows = (To - r + self.s - 1) // self.s
            if rows <= 0:
                continue
            assert taps, "phase without taps (k < stride) is not supported"
            ops.gemm(x, self.w, M=rows, N=self.cout, K=C, batch=B, a_rows=T, a_batch_stride=T * C,
                     b_rows=self.w.shape[0], taps=taps, out=out, out_ld=ld, out_batch_stride=To * ld,
                     out_row_mul=self.s, out_row_off=r, bias=self.bias)
        return out
