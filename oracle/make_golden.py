"""Generate ``tests/golden/*.npz`` by RUNNING THE REFERENCE'S OWN CODE from ``/root/reference``.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference`` does not exist on the GPU
box): ``python oracle/make_golden.py``.  The reference modules are loaded by file path because the
package ``__init__`` files import absent third-party packages.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    dh = load("ref_data_handling", "deadtrees/utils/data_handling.py")
    losses = load("ref_losses_mod", "deadtrees/loss/losses.py")
    gdl = load("ref_gdl_mod", "deadtrees/loss/gdl.py")
    gwdl = load("ref_gwdl_mod", "deadtrees/loss/gwdl.py")

    # ---- tiler: make/unmake blocks on seeded arrays (incl. the reference's own test vector) -------
    rng = np.random.default_rng(1234)
    cases = {}
    for i, (p, m, n, d) in enumerate([(3, 4, 4, 2), (4, 32, 48, 8), (1, 64, 16, 16), (4, 96, 64, 32)]):
        x = rng.integers(0, 256, size=(p, m, n), dtype=np.uint8) if i else np.array(
            [np.arange(16).reshape(4, 4)] * 3, dtype=np.uint8)
        blocks = dh.make_blocks_vectorized(x, d)
        pred = rng.integers(0, 3, size=(blocks.shape[0], d, d), dtype=np.int64)
        merged = dh.unmake_blocks_vectorized(pred, d, m, n)
        cases[f"x{i}"], cases[f"d{i}"] = x, np.int64(d)
        cases[f"blocks{i}"], cases[f"pred{i}"], cases[f"merged{i}"] = blocks, pred, merged
    cases["ncases"] = np.int64(4)
    np.savez_compressed(OUT / "tiler_blocks.npz", **cases)

    # ---- losses on seeded probabilities ---------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    out = {}
    for i, (B, K, H, W) in enumerate([(2, 3, 16, 16), (3, 2, 8, 24), (2, 3, 32, 32)]):
        logits = torch.randn(B, K, H, W, generator=g) * 2.0
        mask = torch.randint(0, K, (B, H, W), generator=g)
        if i == 2:
            mask[mask == 2] = 1  # class 2 absent -> GDL weight 1e9 path
        probs = logits.softmax(dim=1)
        onehot = losses.class2one_hot(mask, K)
        dist = torch.randn(B, K, H, W, generator=g)
        fg = list(range(1, K))
        out[f"logits{i}"], out[f"mask{i}"] = logits.numpy(), mask.numpy()
        out[f"onehot{i}"] = onehot.numpy()
        out[f"dice{i}"] = losses.DiceLoss(idc=fg)(probs, onehot).numpy()
        out[f"focal{i}"] = losses.FocalLoss(idc=list(range(K)), gamma=2)(probs, onehot).numpy()
        out[f"gdl{i}"] = gdl.GeneralizedDiceLoss()(probs, onehot).numpy()
        out[f"dist{i}"] = dist.numpy()
        out[f"surface{i}"] = losses.BoundaryLoss(idc=fg)(probs, dist).numpy()
        # gradients w.r.t. logits through the reference code (autograd), for the backward kernels
        for name, fn in (("dice", lambda p: losses.DiceLoss(idc=fg)(p, onehot)),
                         ("focal", lambda p: losses.FocalLoss(idc=list(range(K)), gamma=2)(p, onehot)),
                         ("gdl", lambda p: gdl.GeneralizedDiceLoss()(p, onehot))):
            z = logits.clone().requires_grad_(True)
            fn(z.softmax(dim=1)).backward()
            out[f"grad_{name}{i}"] = z.grad.numpy()
        # Generalized Wasserstein Dice loss as SemSegment builds and calls it (segmodel.py:118-124, 176-178): on the
        # PROBABILITIES (soft-maxed again inside) and directly on scores; gradients w.r.t. the logits through autograd
        dist_mat = np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]])
        if K == 2:
            dist_mat = dist_mat[0:2, 0:2]
        gw = gwdl.GeneralizedWassersteinDiceLoss(dist_matrix=dist_mat)
        for name, pre in (("gwdl_probs", lambda z: z.softmax(dim=1)), ("gwdl_scores", lambda z: z)):
            z = logits.clone().requires_grad_(True)
            val = gw(pre(z), torch.argmax(onehot, dim=1))
            val.backward()
            out[f"{name}{i}"], out[f"grad_{name}{i}"] = val.detach().numpy(), z.grad.numpy()
    # a 3-class matrix whose maximum is not 1 (normalised by the constructor, gwdl.py:73-78)
    gw = gwdl.GeneralizedWassersteinDiceLoss(dist_matrix=np.array([[0.0, 2.0, 4.0], [2.0, 0.0, 1.0], [4.0, 1.0, 0.0]]))
    out["gwdl_unnorm0"] = gw(torch.from_numpy(out["logits0"]), torch.from_numpy(out["mask0"])).numpy()
    out["ncases"] = np.int64(3)
    np.savez_compressed(OUT / "losses.npz", **out)

    # ---- distance maps of the boundary loss: the reference's one_hot2dist (np.bool alias restored for numpy >= 1.24) ----
    if not hasattr(np, "bool"):
        np.bool = bool
    rng = np.random.default_rng(77)
    dm = {}
    shapes = [(3, 24, 20), (3, 32, 32), (2, 16, 40), (3, 12, 12), (3, 9, 7)]
    for i, (K, H, W) in enumerate(shapes):
        if i == 0:      # blobs
            yy, xx = np.mgrid[0:H, 0:W]
            lab = ((np.hypot(yy - 8, xx - 6) < 5).astype(np.int64) + 2 * (np.hypot(yy - 17, xx - 14) < 4)).clip(0, K - 1)
        elif i == 3:    # a single class everywhere (no background pixel for that class) and two absent classes
            lab = np.zeros((H, W), dtype=np.int64)
        else:
            lab = rng.integers(0, K, size=(H, W)) if i != 1 else (rng.random((H, W)) < 0.03).astype(np.int64) * 2
        onehot = losses.class2one_hot(torch.from_numpy(lab)[None], K)[0].numpy()          # int32, as in the dataloader
        dm[f"labels{i}"], dm[f"K{i}"] = lab, np.int64(K)
        dm[f"dist_int{i}"] = losses.one_hot2dist(onehot, resolution=[1, 1])                # dtype of seg: truncating
        dm[f"dist_f32{i}"] = losses.one_hot2dist(onehot, resolution=[1, 1], dtype=np.float32)
    dm["ncases"] = np.int64(len(shapes))
    np.savez_compressed(OUT / "dist.npz", **dm)
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
