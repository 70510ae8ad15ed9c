"""Precision of the split-operand GEMMs against fp64, per shape (diagnostic; run on a B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlgym_ppo_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def split_rows(x, parts, ps):
    out = torch.zeros((x.shape[0], parts * ps), dtype=torch.bfloat16, device=dev)
    ops.rows_split(x.contiguous(), out, parts, ps)
    return out


def join(t, parts, ps, cols):
    return sum(t[:, q * ps:q * ps + cols].float() for q in range(parts))


for M, N, K in [(5000, 256, 96), (5000, 256, 256), (5000, 96, 256), (777, 64, 21), (5000, 128, 89)]:
    Kp, Np = ops.pad64(K), ops.pad64(N)
    x = torch.randn(M, K, device=dev) * (torch.rand(M, K, device=dev) > 0.3)
    w = (torch.rand(N, K, device=dev) - 0.5) / 8
    b = torch.randn(N, device=dev) * 0.1
    xs = split_rows(x, 3, Kp)
    assert rel(join(xs, 3, Kp, K), x) < 1e-7
    n8 = ops.pad8(N)
    wq = torch.zeros((n8, 3 * Kp), dtype=torch.bfloat16, device=dev)
    wt = torch.zeros((ops.pad8(K), 2 * Np), dtype=torch.bfloat16, device=dev)
    ops.weight_split(w.contiguous(), wq, 3, Kp, wt, 2, Np)
    print(f"M{M} N{N} K{K}: wq parts err {rel(join(wq[:N], 3, Kp, K), w):.1e}, wt parts err "
          f"{rel(join(wt[:K], 2, Np, N), w.t()):.1e}")
    if N % 8 == 0:
        y = torch.zeros((M, 3 * Np), dtype=torch.bfloat16, device=dev)
        ops.linear_fwd(xs, wq, b, y, N, K, True, M=M, split=ops.make_split(3, 3, 3, 3, Kp, Kp, Np))
        want = torch.relu(x.double() @ w.double().t() + b.double())
        print(f"   fwd 3x3 order3: {rel(join(y, 3, Np, N), want):.2e}")
        y1 = torch.zeros((M, Np), dtype=torch.bfloat16, device=dev)
        ops.linear_fwd(xs, wq, b, y1, N, K, True, M=M, split=ops.make_split(1, 1, 1, 1, 0, 0, 0))
        print(f"   fwd plain bf16 through the split entry: {rel(y1[:, :N].float(), want):.2e}")
    # dgrad: dX[M,K] = dY[M,N] W[N,K]  (mask from hprev = x here, so mask = x > 0)
    if K % 8 == 0:
        dy = torch.randn(M, N, device=dev) * 1e-6
        dys = split_rows(dy, 2, Np)
        dx = torch.zeros((M, 2 * Kp), dtype=torch.bfloat16, device=dev)
        ops.linear_dgrad(dys, wt, xs, dx, N, K, M=M, split=ops.make_split(2, 2, 2, 2, Np, Np, Kp))
        want = (dy.double() @ w.double()) * (x > 0)
        print(f"   dgrad 2x2 order2: {rel(join(dx, 2, Kp, K), want):.2e}")
        # wgrad: dW[N,K] = dY^T X
        dw = torch.zeros((N, K), device=dev)
        db = torch.zeros(N, device=dev)
        ops.linear_wgrad(dys, xs, dw, db, N, K, M=M, split=ops.make_split(2, 2, 2, 1, Np, Kp))
        print(f"   wgrad 2x2 order2: {rel(dw, dy.double().t() @ x.double()):.2e}  db {rel(db, dy.double().sum(0)):.2e}")
torch.cuda.synchronize()
