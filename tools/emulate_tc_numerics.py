"""CPU emulation of the tensor-core path's operand rounding, against the fp64 oracle, at cfg3's own horizon.

Design aid (not product code, not a test): reproduces WHERE the dense tcgen05 path rounds (state / chain slabs / tap weights /
operator as bf16 operands, fp32 accumulate) with torch CPU tensors, so that candidate precision modes can be compared on
N=1024, F=64, K=5, T=64 under the reference init before any kernel is written.

    python tools/emulate_tc_numerics.py [T] [B]

Modes:  fp32   everything in float32 (calibration: the exact path's own drift against fp64)
        bf16   operands rounded to bf16 (8-bit mantissa)                      -> GCRNN_PREC_BF16_TC
        bf16x2 operands kept as hi + lo bf16 pairs (16-bit mantissa)          -> GCRNN_PREC_BF16X2_TC
        tf32   operands rounded to tf32 (10-bit mantissa)
"""
import sys
import os
import math

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gcrnn_oracle as orc          # noqa: E402
import gated_gcrnns_b200.graphs as graphs       # noqa: E402


def rnd_bits(x, bits):
    """round-to-nearest-even to `bits` explicit mantissa bits (bf16: 7, tf32: 10)."""
    if bits >= 52:
        return x
    if bits == 7:
        return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)
    m, e = torch.frexp(x)
    s = 2.0 ** (bits + 1)
    return torch.ldexp(torch.round(m * s) / s, e)


class Round(torch.autograd.Function):
    """y = round(x) forward; the incoming gradient is rounded the same way (the adjoint chain's operands are rounded too)."""
    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return do_round(x, mode)

    @staticmethod
    def backward(ctx, g):
        return do_round(g, ctx.mode), None


def do_round(x, mode):
    if mode == 'bf16':
        return rnd_bits(x, 7)
    if mode == 'bf16x2':
        hi = rnd_bits(x, 7)
        return hi + rnd_bits(x - hi, 7)
    if mode == 'tf32':
        return rnd_bits(x, 10)
    return x


def lsigf_r(h, S, x, b, mode):
    """LSIGF with operand rounding: chain input, every chain slab and the taps are rounded; accumulation is fp64."""
    F, E, K, G = h.shape
    hr = Round.apply(h, mode)
    z = Round.apply(x, mode)
    y = torch.einsum('fg,bgn->bfn', hr[:, 0, 0, :], z)
    for k in range(1, K):
        z = Round.apply(z @ S[0], mode)
        y = y + torch.einsum('fg,bgn->bfn', hr[:, 0, k, :], z)
    return y + (b.reshape(1, F, 1) if b is not None else 0)


def forward_emul(p, S, X, h0, mode, round_gates=False):
    B, T, G, N = X.shape
    F = p['weight_A'].shape[0]
    Sm = do_round(S, mode) if mode != 'fp32' else S
    h = h0
    out = []
    gm = mode if round_gates else 'none'

    def sub(prefix, x):
        A, Bw, b = p[prefix + 'weight_A'], p[prefix + 'weight_B'], p[prefix + 'bias']
        # gate sub-cells: x chain from rounded operands (shared with the main input filter), h0 chain rounded
        return torch.tanh(lsigf_r(A, Sm, x, b, mode) + lsigf_r(Bw, Sm, h0, b, mode))
    for t in range(T):
        x = X[:, t]
        ui = sub('GFL_in.', x).reshape(B, F * N)
        gi = torch.sigmoid(ui @ p['MLP_in.0.weight'].t() + p['MLP_in.0.bias']).reshape(B, 1, 1)
        uf = sub('GFL_forget.', x).reshape(B, F * N)
        gf = torch.sigmoid(uf @ p['MLP_forget.0.weight'].t() + p['MLP_forget.0.bias']).reshape(B, 1, 1)
        a = lsigf_r(p['weight_A'], Sm, x, p['bias'], mode)
        r = lsigf_r(p['weight_B'], Sm, h, p['bias'], mode)
        h = torch.tanh(gi * a + gf * r)
        out.append(h)
    return torch.stack(out, 1)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    weighted = len(sys.argv) > 3 and sys.argv[3] == 'weighted'
    N, F, K, G = 1024, 64, 5, 1
    torch.set_num_threads(os.cpu_count())
    S = graphs.dense_random(N, 0.3, seed=0).double()
    if weighted:
        g = torch.Generator().manual_seed(5)
        W = torch.rand(N, N, generator=g, dtype=torch.float64)
        W = (W + W.t()) / 2
        S = S * W
        S = S / torch.linalg.eigvalsh(S[0]).abs().max()
    torch.manual_seed(0)
    p = orc.init_cell_params(G, F, K, K, N, True, None, 1, True)
    p = {k: v.double() for k, v in p.items()}
    torch.manual_seed(5)
    X = torch.randn(B, T, G, N, dtype=torch.float64)
    dH = torch.randn(B, T, F, N, dtype=torch.float64)
    for h0name, h0 in (('h0=0', torch.zeros(B, F, N, dtype=torch.float64)), ('h0~0.3N', 0.3 * torch.randn(B, F, N, dtype=torch.float64))):
        res = {}
        for mode in ('fp64', 'fp32', 'bf16', 'tf32', 'bf16x2'):
            q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
            if mode == 'fp32':
                q32 = {k: v.detach().float().requires_grad_(True) for k, v in p.items()}
                H = orc.cell_forward(q32, S.float(), X.float(), h0.float(), True, None)
                (H * dH.float()).sum().backward()
                res[mode] = (H.detach().double(), {k: v.grad.double() for k, v in q32.items() if v.grad is not None})
                continue
            H = forward_emul(q, S, X, h0, 'none' if mode == 'fp64' else mode)
            (H * dH).sum().backward()
            res[mode] = (H.detach(), {k: v.grad for k, v in q.items() if v.grad is not None})
        Href, gref = res['fp64']
        print(f'== {h0name}  T={T} B={B} weighted={weighted}  max|H|={Href.abs().max():.3f}')
        for mode in ('fp32', 'bf16', 'tf32', 'bf16x2'):
            H, g = res[mode]
            curve = [((H[:, t] - Href[:, t]).abs().max() / Href.abs().max()).item() for t in range(T)]
            pick = [0, 1, 3, 7, 15, 31, 47, T - 1]
            ge = {k: ((g[k] - gref[k]).abs().max() / gref[k].abs().max().clamp_min(1e-300)).item() for k in gref}
            worst = max(ge.items(), key=lambda kv: kv[1])
            print(f'{mode:7s} H err @t{[t for t in pick if t < T]}: ' + ' '.join(f'{curve[t]:.1e}' for t in pick if t < T)
                  + f' | max {max(curve):.1e} | worst grad {worst[0]} {worst[1]:.1e} | weight_B {ge["weight_B"]:.1e} weight_A {ge["weight_A"]:.1e}')


if __name__ == '__main__':
    main()
