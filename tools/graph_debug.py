import os, sys, types, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from csmpn_b200.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17
from csmpn_b200.train_step import DataParallelStep, GraphedDataParallelStep, FlatGradBucket
dev = torch.device("cuda:0")
fx = load_golden("models.pt")["md17"]
def mk():
    m = CliffordSharedSimplicialMPNN_md17(**fx["kwargs"]).to(dev)
    m.load_state_dict(fx["state_dict"], strict=False)
    g = types.SimpleNamespace(**{k: v.clone().to(dev) for k, v in fx["batch"].items()})
    return m, g
def run(graphed, sync_each, tag):
    m, g = mk()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    st = GraphedDataParallelStep(m, opt, g) if graphed else DataParallelStep(m, opt)
    out = []
    for _ in range(3):
        loss, _ = st(g)
        if sync_each: out.append(float(loss.detach()))
        else: out.append(loss.detach().clone())
    print(tag, [float(x) for x in out], flush=True)
run(True, True, "B graphed first, float() each step")
run(True, False, "B2 graphed, clone loss each step")
run(False, True, "eager")
run(True, True, "A graphed after eager")
