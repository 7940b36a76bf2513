"""Golden outputs of the UNMODIFIED reference ``heldout_mask`` (src/utils/eval_utils.py:988-1045; extracted from the
source text because the module imports packages that are absent here) for every evaluation mode.

    python tests/golden/make_heldout_golden.py      (build container; writes heldout.npz)
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/utils/eval_utils.py"
REGIONS = np.array(["CA1", "DG", "LP", "PO"])

CASES = [
    ("manual", dict(heldout_idxs=np.array([0, 3, 5]))),
    ("manual", dict(heldout_idxs=np.array([2, 4]))),
    ("most", dict(n_active=3)),
    ("inter_region", dict(heldout_idxs=np.array([0, 1]), target_regions=["DG", "CA1"])),
    ("intra_region", dict(heldout_idxs=np.array([1]), target_regions=["LP"])),
    ("intra_region", dict(heldout_idxs=np.array([]), target_regions=["CA1"])),
    ("forward_pred", dict(heldout_idxs=np.arange(12, 20))),
    ("modal_spike", dict(heldout_idxs=np.arange(0, 20))),
    ("modal_behavior", dict(heldout_idxs=np.array([0, 7, 19]))),
]


def inputs():
    g = torch.Generator().manual_seed(3)
    spikes = torch.poisson(torch.rand(4, 20, 12, generator=g) * 2.0, generator=g)
    regions = REGIONS[np.arange(12) % 4]
    return spikes, regions


def main():
    tree = ast.parse(open(SRC).read())
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "heldout_mask":
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    spikes, regions = inputs()
    z = {}
    for i, (mode, kw) in enumerate(CASES):
        r = ns["heldout_mask"](spikes.clone(), mode=mode, neuron_regions=regions, **kw)
        z[f"case{i}/spikes"] = r["spikes"].numpy()
        z[f"case{i}/eval_mask"] = r["eval_mask"].numpy().astype(np.int8)
        z[f"case{i}/hd"] = np.asarray(r["heldout_idxs"]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "heldout.npz"), **z)
    print("written", len(CASES), "cases")


if __name__ == "__main__":
    main()
