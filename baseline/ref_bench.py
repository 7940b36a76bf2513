"""Timing of the UNMODIFIED reference (``baseline/_ref``) for ``bench.py``: the CPU arm (``--impl reference`` and the
``cpu_baseline`` key) and the "library bar" (the same reference model on the B200 through stock PyTorch: fp32 with
TF32 off, and ``torch.autocast(bfloat16)``).  Benchmark infrastructure; nothing in the product package imports it.

The model is built exactly as ``train_multi_modal.py:160-189`` does, from the reference's own ``mm.yaml`` /
``trainer_mm.yaml`` (only the workload's neuron count is a constructor argument, as in the script), and driven with the
``mod_dict`` the trainer builds (``trainer/base.py:51-103``; ``synthetic.make_mod_dict`` mirrors it key for key).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

from . import ref_loader

MODES = ("encoding", "decoding", "token_masking")


def _model(neurons: int, n_beh: int, train: bool, device="cpu"):
    import torch
    cfg = ref_loader.load_config()
    torch.manual_seed(42)
    m = ref_loader.build_reference_model(cfg, neurons, n_beh).to(device)
    m.train(train)
    return m


def cpu_rate(neurons: int, n_beh: int, T: int, trials: int, *, steps: int = 0, warmup: int = 1, seconds: float = 0.0,
             train: bool = True, threads: Optional[int] = None, modes=MODES) -> Dict[str, object]:
    """fwd+bwd trials/s of the unmodified reference ``MultiModal`` on the host CPU, fp32, ``trials`` trials per step.
    Either exactly ``steps`` timed steps, or as many as fit in ``seconds`` (at least 2)."""
    import torch
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    cores = os.cpu_count() or 1
    threads = threads or cores
    torch.set_num_threads(threads)
    model = _model(neurons, n_beh, train)
    times: List[float] = []

    def one(i):
        batch = make_batch(trials, neurons, n_beh, T, step=i)
        md = make_mod_dict(batch, ["ap", "behavior"], modes[i % len(modes)])
        t0 = time.perf_counter()
        out = model(md)
        out.loss.backward()
        model.zero_grad(set_to_none=True)
        return time.perf_counter() - t0

    for i in range(warmup):
        one(i)
    t_start = time.perf_counter()
    i = warmup
    while True:
        times.append(one(i))
        i += 1
        if steps and len(times) >= steps:
            break
        if not steps and len(times) >= 2 and time.perf_counter() - t_start >= seconds:
            break
    per = sum(times) / len(times)
    torch.set_num_threads(cores)
    return {"value": trials / per, "s_per_step": per, "steps": len(times), "warmup": warmup, "threads": threads,
            "cores": cores, "trials_per_step": trials, "mode": "train()" if train else "eval()"}


def library_bar(neurons: int, n_beh: int, T: int, B: int, device, *, steps: int = 6, warmup: int = 3,
                train: bool = True) -> Dict[str, object]:
    """The same unmodified reference model on the GPU under stock PyTorch (cuBLAS / SDPA / ATen kernels), B trials per
    step, CUDA-event timed.  Legs: fp32 (TF32 off) and bf16 autocast, each with the trainer's three modes cycled (what
    ``--mixed_training`` runs; one step in three pays the reference Masker's host-side draws) and with the two
    masker-free modes only (pure library-kernel time)."""
    import torch
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = _model(neurons, n_beh, train, device)
    batches = [{k: (v.to(device) if torch.is_tensor(v) else v) for k, v in make_batch(B, neurons, n_beh, T, step=i).items()}
               for i in range(3)]

    def run(modes, autocast):
        def one(i):
            md = make_mod_dict(batches[i % 3], ["ap", "behavior"], modes[i % len(modes)], device=device)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out = model(md)
            out.loss.backward()
            model.zero_grad(set_to_none=True)
        for i in range(warmup):
            one(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / (ms * 1e-3), "ms_per_step": ms}

    out = {"unit": "trials/s", "batch": B, "steps": steps, "warmup": warmup,
           "what": "unmodified reference MultiModal (baseline/_ref) .cuda(), stock PyTorch " + torch.__version__
                   + " kernels (cuBLAS / SDPA / ATen), fwd + loss.backward(), " + ("train()" if train else "eval()")}
    out["fp32_cycled"] = run(MODES, False)
    out["bf16_autocast_cycled"] = run(MODES, True)
    out["fp32_no_masker"] = run(MODES[:2], False)
    out["bf16_autocast_no_masker"] = run(MODES[:2], True)
    del model
    torch.cuda.empty_cache()
    return out
