"""Fused AdamW for the B200 path.

Drop-in for the reference's ``torch.optim.AdamW(model.parameters(), lr=..., weight_decay=..., eps=...)``
(``train_multi_modal.py:197-202``): same constructor, ``param_groups`` (so ``OneCycleLR`` of ``:204-210`` keeps
driving ``lr`` / ``betas``), ``step()`` / ``zero_grad()`` / ``state_dict()``.  The engine keeps every parameter and
gradient of the model in two flat fp32 buffers (engine.ParamStore), so one ``mmfm_adamw_step`` launch updates the
whole model: 28 bytes of HBM traffic per parameter instead of one multi-tensor launch chain per group.  The bf16
shadow weights are refreshed by the forward schedule's multi-tensor cast, which reads the updated masters.

There is no fallback: parameters that do not live in an engine's flat buffer raise.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import ops
from ._lib import MmfmError


class AdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, *, maximize: bool = False, **unused):
        if amsgrad or maximize:
            raise NotImplementedError("the fused AdamW implements the reference's configuration: amsgrad=False, "
                                      "maximize=False")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("one parameter group (the reference passes model.parameters())")
        self._store = None
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._steps = 0

    # -----------------------------------------------------------------------------------------------------
    def _bind(self):
        """Locate the engine store that owns the parameters (created lazily by the first forward pass)."""
        from .engine import live_stores
        ps = self.param_groups[0]["params"]
        for st in live_stores():
            if not st.adopted():
                st.adopt()
            lo, hi = st.flat.data_ptr(), st.flat.data_ptr() + st.flat.numel() * 4
            if all(lo <= p.data_ptr() < hi for p in ps):
                owned = {id(q) for q in st.params.values()}
                if {id(p) for p in ps} != owned:
                    raise MmfmError("the fused AdamW updates the whole flat buffer: pass model.parameters() of the "
                                    "complete model (frozen or foreign parameters are not supported)")
                self._store = st
                self._m = torch.zeros_like(st.flat)
                self._v = torch.zeros_like(st.flat)
                for n, p in st.params.items():      # torch-style per-parameter state as views of the flat moments
                    self.state[p] = {"step": torch.tensor(float(self._steps)), "exp_avg": st.view(self._m, n),
                                     "exp_avg_sq": st.view(self._v, n)}
                return
        raise MmfmError("parameters are not owned by a B200 engine: move the model to CUDA and run one forward pass "
                        "(model(mod_dict)) before optimizer.step(); there is no fallback optimizer")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._store is None:
            self._bind()
        st = self._store
        if not st.adopted():
            st.adopt()
        g = self.param_groups[0]
        # gradients must be the flat buffer's views (engine.backward installs them); anything else is copied in
        skipped = []     # parameters without a gradient: torch.optim.AdamW leaves them (weights, moments, step) untouched
        for n, p in st.params.items():
            if p.grad is None:
                skipped.append((n, st.p(n).clone(), st.view(self._m, n).clone(), st.view(self._v, n).clone()))
                st.g(n).zero_()
            elif p.grad.data_ptr() != st.g(n).data_ptr():
                st.g(n).copy_(p.grad)
        self._steps += 1
        b1, b2 = g["betas"]
        ops.adamw_step(st.flat, st.grad, self._m, self._v, lr=g["lr"], beta1=b1, beta2=b2, eps=g["eps"],
                       weight_decay=g["weight_decay"], step=self._steps)
        for n, w, m, v in skipped:      # the one launch covers the whole flat buffer: put the skipped ranges back
            st.p(n).copy_(w)
            st.view(self._m, n).copy_(m)
            st.view(self._v, n).copy_(v)
        skip = {id(st.params[n]) for n, *_ in skipped}
        for q, s in self.state.items():
            if id(q) not in skip:
                s["step"] += 1
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if self._store is None:
            return
        st = self._store
        for n, p in st.params.items():          # re-home the loaded moments into the flat buffers
            s = self.state.get(p)
            if s:
                st.view(self._m, n).copy_(s["exp_avg"])
                st.view(self._v, n).copy_(s["exp_avg_sq"])
                s["exp_avg"], s["exp_avg_sq"] = st.view(self._m, n), st.view(self._v, n)
                self._steps = int(float(s["step"]))
