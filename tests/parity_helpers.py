"""Helpers shared by the GPU parity tests."""
import numpy as np


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def close(a, b, tol):
    """scale-aware tolerance of SURVEY.md 8(c): |a - b| <= tol * max(|b|, 1)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0)))


def capture_steps(lr):
    """Post-clip gradients seen by Adam at every optimiser step (the golden scripts wrap optimizer.step the same way)."""
    from rlgym_ppo_b200 import ops
    captured = []
    orig = lr._optimizer_step

    def step():
        ops.grad_sqnorm(lr._grads, lr._seg, lr._sqnorm)
        gn = lr._sqnorm.sqrt().cpu().numpy()
        coef = np.minimum(1.0, 0.5 / (gn + 1e-6))
        gr = lr._grads.cpu().numpy().copy()
        n_p = int(lr._seg[1])
        gr[:n_p] *= coef[0]
        gr[n_p:] *= coef[1]
        captured.append(gr)
        orig()

    lr._optimizer_step = step
    lr.use_cuda_graph = False        # the hook reads tensors back, which a graph capture forbids
    return captured


def unflatten(lr, flat):
    shapes = [tuple(p.shape) for p in lr.policy.parameters()] + [tuple(p.shape) for p in lr.value_net.parameters()]
    out, off = [], 0
    for shp in shapes:
        n = int(np.prod(shp))
        out.append(flat[off:off + n].reshape(shp))
        off += n
    return out
