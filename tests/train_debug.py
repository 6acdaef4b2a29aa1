"""GPU diagnostic: per-parameter gradient error of the training step against CPU autograd, in backward order."""
import copy
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))  # run as: python tests/train_debug.py
import test_gpu_train as T  # noqa: E402
from deadtrees_b200.parallel import backward_param_order  # noqa: E402
from gpu_util import oracle_model  # noqa: E402
from oracle import ref_train  # noqa: E402


def main():
    cin = 4
    for precision, n, tile in (("fp32", 2, 64), ("bf16", 2, 64), ("bf16", 2, 128), ("bf16", 2, 256)):
        oracle = oracle_model(cin, 3)
        ref_model = copy.deepcopy(oracle)
        img, mask = T._batch(n, cin, tile, 3)
        if precision == "bf16":
            ref = ref_train.train_step_bf16(ref_model, img, mask)
        else:
            ref = ref_train.train_step(ref_model.double(), img.double(), mask, lr=0, clip=0)
        seg = T._semsegment(oracle, cin, precision)
        batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": "t"}] * n)}
        logits = seg.model(img.cuda())
        print(f"== {precision} n={n} T={tile}: logits max err {(logits.detach().cpu() - ref['logits']).abs().max():.3e} "
              f"(ref max {ref['logits'].abs().max():.3f})")
        seg2 = T._semsegment(oracle, cin, precision)
        loss = seg2.training_step(batch, 0)
        loss.backward()
        torch.cuda.synchronize()
        print(f"   loss {float(loss.detach()):.6f} vs {ref['loss']:.6f}")
        params = dict(seg2.model.named_parameters())
        for name in backward_param_order(list(params)):
            r = ref["grads"][name].double().flatten()
            g = params[name].grad.cpu().double().flatten()
            rel = float((g - r).norm() / (r.norm() + 1e-30))
            cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
            if name.endswith("weight") and ("conv" in name or "downsample.0" in name or "head" in name):
                print(f"   {name:44s} rel={rel:.3e} cos={cos:.5f} |ref|={float(r.abs().max()):.3e}")


if __name__ == "__main__":
    main()
