"""GPU (>= 2 devices): data-parallel gradients == arithmetic mean of the per-rank single-GPU gradients
(DDP semantics), with the bucketed all-reduce overlapped with backward."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from _util import small_config
    from multi_modal_foundation_model_b200.model import build_model
    from multi_modal_foundation_model_b200.parallel import DataParallel
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.manual_seed(100 + rank)            # different init per rank: the wrapper must broadcast rank 0's
        model = build_model(40, 2, small_config()).cuda().eval()
        ddp = DataParallel(model, bucket_mb=0.25)   # small buckets -> several all-reduces inside backward
        for step in (0, 2, 1):                   # eager, graph capture (backward + all-reduces), graph replay
            batch = make_batch(4, 40, 2, 100, step=10 * step + rank, pad_bins=5)
            md = make_mod_dict(batch, ["ap", "behavior"], "encoding" if step == 0 else "decoding", device="cuda")
            model.zero_grad(set_to_none=True)
            out = model(md)
            out.loss.backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}
        weights = {n: p.detach().cpu().clone() for n, p in model.named_parameters()}
        torch.save({"grads": grads, "weights": weights}, os.path.join(out_dir, f"rank{rank}.pt"))
        ddp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_gradients_are_rank_mean(tmp_path):
    import torch.multiprocessing as mp
    from _util import small_config
    from multi_modal_foundation_model_b200.model import build_model
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(world)]
    for n in r[0]["grads"]:
        assert torch.equal(r[0]["grads"][n], r[1]["grads"][n]), f"ranks disagree on {n}"
        assert torch.equal(r[0]["weights"][n], r[1]["weights"][n]), f"weights not broadcast: {n}"
    # single-GPU reference: same weights, each rank's shard in turn, mean of the two gradients
    torch.manual_seed(100)
    model = build_model(40, 2, small_config()).cuda().eval()
    acc = None
    for rank in range(world):
        batch = make_batch(4, 40, 2, 100, step=10 + rank, pad_bins=5)
        md = make_mod_dict(batch, ["ap", "behavior"], "decoding", device="cuda")
        model.zero_grad(set_to_none=True)
        model(md).loss.backward()
        g = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}
        acc = g if acc is None else {n: acc[n] + g[n] for n in g}
    for n, gref in acc.items():
        gref = gref / world
        got = r[0]["grads"][n]
        scale = gref.abs().max().item() + 1e-8
        assert (got - gref).abs().max().item() <= 2e-3 * scale + 1e-7, n
