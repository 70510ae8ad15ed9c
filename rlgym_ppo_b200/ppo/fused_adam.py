"""torch.optim.Adam-shaped handle over the fused clip+Adam kernel (ppo_learner.py:56-59, 187-193).

The arithmetic runs in rlppo_clip_adam over the learner's flat arenas; this object only carries what the
reference touches from Python: `.param_groups[i]['lr']` (rewritten by Learner.update_learning_rate,
learner.py:205-216) and a `state_dict()` / `load_state_dict()` in stock torch.optim.Adam format so the four
checkpoint files stay interchangeable with the reference's (ppo_learner.py:240-271).
"""
import torch


class FusedAdam(object):
    def __init__(self, stack, lr, m_flat, v_flat, step_view, betas=(0.9, 0.999), eps=1e-8):
        self._stack = stack
        self._m, self._v = m_flat, v_flat          # views into the learner's moment arenas
        self._step = step_view                     # 1-element int64 device view
        n_tensors = 2 * len(stack.linears)
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                             foreach=None, capturable=False, differentiable=False, fused=None,
                             decoupled_weight_decay=False)
        self.param_groups = [dict(self.defaults, params=list(range(n_tensors)))]

    @property
    def lr(self):
        return float(self.param_groups[0]["lr"])

    def zero_grad(self, set_to_none=False):
        self._stack.grads.zero_()

    def _param_shapes(self):
        shapes = []
        for l in self._stack.linears:
            shapes += [tuple(l.weight.shape), tuple(l.bias.shape)]
        return shapes

    def state_dict(self):
        step = float(self._step.item())
        state = {}
        if step > 0:
            off = 0
            for i, shp in enumerate(self._param_shapes()):
                n = 1
                for s in shp:
                    n *= s
                state[i] = {"step": torch.tensor(step),
                            "exp_avg": self._m[off:off + n].view(shp).clone(),
                            "exp_avg_sq": self._v[off:off + n].view(shp).clone()}
                off += n
        groups = [{k: v for k, v in g.items()} for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        state = sd["state"]
        shapes = self._param_shapes()
        off = 0
        step = 0.0
        for i, shp in enumerate(shapes):
            n = 1
            for s in shp:
                n *= s
            if i in state:
                self._m[off:off + n].copy_(state[i]["exp_avg"].reshape(-1))
                self._v[off:off + n].copy_(state[i]["exp_avg_sq"].reshape(-1))
                step = float(state[i]["step"])
            else:
                self._m[off:off + n].zero_()
                self._v[off:off + n].zero_()
            off += n
        self._step.fill_(int(step))
        for g, src in zip(self.param_groups, sd["param_groups"]):
            for k, v in src.items():
                if k != "params":
                    g[k] = v
