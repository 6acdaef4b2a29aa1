"""debug: does any kernel of the training step read memory it has not written?  Poison the caching allocator (and, through
empty_cache, the pages the graph's private pool will get) with NaN patterns and compare against an unpoisoned run."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_train import NETWORK, TRAINING, oracle_model, _batch
from deadtrees_b200.network import SemSegment
from deadtrees_b200.train_graph import GraphedTrainStep

cin, n, T = 4, 2, 128
oracle = oracle_model(cin, 3)
batches = [_batch(n, cin, T, 3), _batch(n, cin, T, 3, seed=12), _batch(n, cin, T, 3)]
stats = [{"file": f"t{i}"} for i in range(n)]
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["GWDICE", "FOCAL"]

def make():
    seg = SemSegment(dict(NETWORK, in_channels=cin, precision="bf16", losses=names), dict(TRAINING, gradient_clip_val=0.5))
    seg.model.load_state_dict(oracle.state_dict())
    seg.cuda().train()
    (opt,), _ = seg.configure_optimizers()
    return seg, opt

REC = []


def eager():
    seg, opt = make()
    out, rec = [], {}
    for i, (a, b) in enumerate(batches):
        loss = seg.training_step({"main": (a.cuda(), b.cuda(), None, torch.zeros(n), stats)}, 0)
        loss.backward()
        if i == 0:
            rec["g"] = opt._flat["g"].clone()
            rec["logged"] = {k: float(v) for k, v in seg.logged.items() if hasattr(v, "item") or isinstance(v, float)}
        opt.step()
        if i == 0:
            rec["p"] = opt._flat["p"].clone()
            rec["scal"] = opt._flat["scratch"].tolist()
            rec["acc"] = opt._flat["acc"].tolist()
        out.append(float(loss.detach()))
    eng = seg.model.train_engine()
    rec["layout"] = [(nm, eng.reducer.flat_offset(nm), p.numel()) for nm, p in eng.params.items()]
    REC.append(rec)
    return out, opt._flat["p"].clone()


def compare(a, b):
    print("   logged:", {k: (a["logged"][k], b["logged"][k]) for k in a["logged"] if a["logged"][k] != b["logged"][k]})
    print("   adam scal", a["scal"], b["scal"], "sumsq", a["acc"], b["acc"])
    bad = []
    for nm, off, cnt in a["layout"]:
        dg = (a["g"][off:off + cnt] - b["g"][off:off + cnt]).abs().max().item()
        dp = (a["p"][off:off + cnt] - b["p"][off:off + cnt]).abs().max().item()
        ref = a["g"][off:off + cnt].abs().max().item()
        if dg > 1e-4 * ref + 1e-12 or dp > 1e-7:
            bad.append((nm, dg, ref, dp))
    print(f"   {len(bad)} of {len(a['layout'])} parameters differ after step 0:")
    for t in bad[:40]:
        print("     %-48s grad diff %.3e (max |g| %.3e)  param diff %.3e" % t)

def graphed():
    seg, opt = make()
    gs = GraphedTrainStep(seg, opt, n, T)
    out = [float(gs(a.pin_memory(), b.pin_memory())) for a, b in batches]
    return out, opt._flat["p"].clone()

def poison(release: bool):
    big = [torch.full(((256 << 20) // 4,), float("nan"), device="cuda") for _ in range(24)]          # 6 GB of large blocks
    mid = [torch.full(((2 << 20) // 4,), float("nan"), device="cuda") for _ in range(512)]           # 1 GB of 2 MB blocks
    small = [torch.full((s // 4,), float("nan"), device="cuda") for s in (512, 4096, 65536, 524288) for _ in range(2000)]
    torch.cuda.synchronize()
    del big, mid, small
    if release:
        torch.cuda.empty_cache()          # cudaFree: the next cudaMalloc (the graph's private pool) may get these pages back

base_l, base_p = eager()
print("eager baseline        ", base_l)
poison(False)
l, p = eager()
print("eager, poisoned cache ", l, "p diff", (p - base_p).abs().max().item())
compare(REC[0], REC[-1])
poison(False)
l, p = graphed()
print("graph, poisoned cache ", l, "p diff", (p - base_p).abs().max().item())
poison(True)
l, p = graphed()
print("graph, poisoned pages ", l, "p diff", (p - base_p).abs().max().item())
poison(True)
l, p = eager()
print("eager, poisoned pages ", l, "p diff", (p - base_p).abs().max().item())
compare(REC[0], REC[-1])
