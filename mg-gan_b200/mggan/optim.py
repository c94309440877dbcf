"""AdamW with fused global-norm clipping (reference set-up: mggan/abstract_train.py:45-50,
steps mggan/model/train.py:131-135, :209-213, :656-658).

A `torch.optim.Optimizer` subclass, so `state_dict()` / `load_state_dict()` keep the
`{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}` layout that reference
checkpoints store under "gen_opt" / "disc_opt".  `step()` launches two kernels per <=64 tensors
(`mggan_grad_sqnorm`, `mggan_clip_adamw`) instead of a Python loop over ~100 tensors; tensors
whose `.grad` is None are skipped and keep their step count, like torch.optim.AdamW.
"""
import torch

from . import kernels as K


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.last_grad_sqnorm = None
        self.graph = None        # a mggan.graph.GraphedIteration while an iteration is being captured

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]):
            # checkpoints written by the reference's pinned torch 1.6 store `step` as a Python int
            st["step"] = torch.tensor(float(st["step"]))
        return st

    @torch.no_grad()
    def step(self, closure=None, max_norm=0.0, grad_scale=1.0, reduce_fn=None):
        """max_norm > 0 clips the global L2 norm of all gradients first (clip_grad_norm_ semantics:
        coefficient min(1, max_norm / (norm + 1e-6))).  `reduce_fn(list_of_grads)` is called before
        the norm (data-parallel all-reduce hook)."""
        assert closure is None
        work = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    work.append((group, p))
        if not work:
            return
        grads = [p.grad.contiguous() for _, p in work]
        sq = None
        if reduce_fn is not None:
            red = reduce_fn(grads)
            if isinstance(red, tuple):          # the peer-memory reducer returns the squared norm of the reduced gradient too
                red, sq = red
            grads = red or grads
        if not (max_norm and max_norm > 0):
            sq = None
        elif sq is None:
            sq = K.grad_sqnorm(grads)
        self.last_grad_sqnorm = sq
        start = 0
        while start < len(work):             # one launch set per param group
            group = work[start][0]
            end = start
            while end < len(work) and work[end][0] is group:
                end += 1
            ps, gs, ms, vs, steps = [], [], [], [], []
            for (_, p), g in zip(work[start:end], grads[start:end]):
                st = self._init_state(p)
                if self.graph is None:
                    st["step"] += 1
                ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
                steps.append(max(int(st["step"].item()), 1))
            b1, b2 = group["betas"]
            dyn = None
            if self.graph is not None:
                # captured iteration: lr and the bias corrections come from device memory that the replay loop
                # refreshes (it also advances the step counters, so nothing is counted during the capture pass)
                dyn = [K.ptr(self.graph.adam_table(self, group, ps[c:c + K.X.TABLE_MAX]))
                       for c in range(0, len(ps), K.X.TABLE_MAX)]
            K.clip_adamw(ps, gs, ms, vs, steps, sq, max_norm or 0.0, group["lr"], b1, b2, group["eps"],
                         group["weight_decay"], grad_scale, dyn=dyn)
            start = end
