# -*- coding: utf-8 -*-
"""``torch.optim.Adam`` for the training loop (reference bin/train.py:457-460, ``optimizer.step()`` :539) whose step is
ONE kernel launch (``wnb_adam_flat``, csrc/adam.cu).

The backward of ``nets.WaveNet`` leaves every ``p.grad`` as a slice of one flat buffer (``module._wnb_flat_grad``).  On
its first step this optimizer lays the parameters (``p.data`` becomes a view, values unchanged) and both moment estimates
out the same way; after that a step is a single streaming pass over four flat arrays instead of torch's multi-tensor
kernels over 184 reference-shaped tensors (11 launches, 170 us of an 8.5 ms step at the BASELINE architecture).

It IS a ``torch.optim.Adam``: same constructor arguments, same arithmetic (fused-Adam formulas in fp32), ``state_dict()``
/ ``load_state_dict()`` in torch's format (``exp_avg`` / ``exp_avg_sq`` / ``step`` per parameter -- checkpoints written
by ``bin/train.py`` keep loading into the reference's ``torch.optim.Adam`` and the other way round).  Whenever the flat
layout is not available -- gradients that are not slices of one buffer (the per-block and composed paths), several
parameter groups, amsgrad / maximize, CPU parameters -- it falls back to ``torch.optim.Adam.step`` (fused)."""
import torch

from ._lib import check, load, ptr, stream


class Adam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, module=None, **kw):
        params = list(params)
        kw.setdefault("fused", bool(params) and all(torch.is_tensor(p) and p.is_cuda for p in params))
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        self._module = module       # the nets.WaveNet whose backward fills module._wnb_flat_grad
        self._flat = None           # dict(p=, m=, v=, plist=, offs=, k=, step_t=) once laid out
        self.flat_steps = 0         # steps taken through the one-kernel path (tests / bench)

    # ---- layout -------------------------------------------------------------------------------------------------
    def _flat_possible(self, g):
        if g is None or len(self.param_groups) != 1:
            return False
        grp = self.param_groups[0]
        if grp.get("amsgrad") or grp.get("maximize") or grp.get("capturable") or grp.get("differentiable") or \
                grp.get("decoupled_weight_decay") or torch.is_tensor(grp["lr"]):
            return False
        return g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.data_ptr() % 16 == 0

    def _offsets(self, g):
        """offset (floats) of every parameter's gradient inside the flat buffer g, or None when they do not tile it"""
        base, n = g.data_ptr(), g.numel()
        plist, offs, tot = [], [], 0
        for p in self.param_groups[0]["params"]:
            if p.grad is None:
                continue                                  # torch.optim.Adam skips these as well
            q = p.grad
            d = q.data_ptr() - base
            if q.dtype != torch.float32 or not q.is_contiguous() or d < 0 or d % 4 or d // 4 + q.numel() > n or \
                    q.numel() != p.numel() or p.dtype != torch.float32 or p.device != g.device:
                return None
            plist.append(p)
            offs.append(d // 4)
            tot += q.numel()
        if tot != n or not plist:
            return None
        order = sorted(range(len(plist)), key=lambda i: offs[i])
        for a, b in zip(order[:-1], order[1:]):
            if offs[a] + plist[a].numel() != offs[b]:
                return None
        return plist, offs

    def _setup(self, g):
        lay = self._offsets(g)
        if lay is None:
            return False
        plist, offs = lay
        pf, mf, vf = torch.empty_like(g), torch.zeros_like(g), torch.zeros_like(g)
        k = 0
        for p, o in zip(plist, offs):
            n = p.numel()
            pf[o:o + n].copy_(p.detach().reshape(-1))
            st = self.state.get(p, {})
            if "exp_avg" in st:                           # resumed / previously stepped through torch's path
                mf[o:o + n].copy_(st["exp_avg"].reshape(-1))
                vf[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                k = max(k, int(float(st["step"])))
        step_t = torch.full((), float(k), dtype=torch.float32, device=g.device)
        for p, o in zip(plist, offs):
            n = p.numel()
            p.data = pf[o:o + n].view(p.shape)            # same values, now a slice of the flat buffer
            self.state[p] = {"step": step_t, "exp_avg": mf[o:o + n].view(p.shape), "exp_avg_sq": vf[o:o + n].view(p.shape)}
        self._flat = dict(p=pf, m=mf, v=vf, plist=plist, offs=offs, k=k, step_t=step_t, calls=0)
        return True

    def _still_flat(self, g):
        f = self._flat
        f["calls"] += 1
        plist, offs, base = f["plist"], f["offs"], g.data_ptr()
        if g.numel() != f["p"].numel():
            return False
        idx = range(len(plist)) if f["calls"] % 32 == 1 else (0, len(plist) // 2, len(plist) - 1)
        pb = f["p"].data_ptr()
        for i in idx:
            p = plist[i]
            if p.grad is None or p.grad.data_ptr() != base + 4 * offs[i] or p.data_ptr() != pb + 4 * offs[i]:
                return False
        if f["calls"] % 32 == 1:
            with_grad = sum(1 for p in self.param_groups[0]["params"] if p.grad is not None)
            return with_grad == len(plist)
        return True

    def _leave_flat(self):
        """before torch's own step runs again: every parameter gets its own step counter back (torch adds 1 per entry)"""
        if self._flat is not None:
            for p in self._flat["plist"]:
                self.state[p]["step"] = self._flat["step_t"].clone()
            self._flat = None

    # ---- torch.optim.Optimizer surface --------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        g = getattr(self._module, "_wnb_flat_grad", None) if self._module is not None else None
        ok = self._flat_possible(g)
        if ok and self._flat is not None and not self._still_flat(g):
            self._leave_flat()
        if ok and self._flat is None:
            ok = self._setup(g)
        if not ok:
            self._leave_flat()
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        f, grp = self._flat, self.param_groups[0]
        f["k"] += 1
        f["step_t"].add_(1.0)
        b1, b2 = grp["betas"]
        with torch.cuda.device(g.device):
            check(load().wnb_adam_flat(ptr(f["p"]), ptr(g), ptr(f["m"]), ptr(f["v"]), g.numel(), float(grp["lr"]), float(b1),
                                       float(b2), float(grp["eps"]), float(grp["weight_decay"]), 1.0 - b1 ** f["k"],
                                       1.0 - b2 ** f["k"], stream()), "adam_flat")
        self.flat_steps += 1
        return loss

    def load_state_dict(self, state_dict):
        self._flat = None                                  # (state tensors are replaced; laid out again at the next step)
        return super().load_state_dict(state_dict)

    def add_param_group(self, param_group):
        if getattr(self, "_flat", None) is not None:
            self._leave_flat()
        return super().add_param_group(param_group)
