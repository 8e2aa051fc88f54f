import os, sys; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import contextlib, torch
import bench, yfv2, yfv2_engine as eng
dev = torch.device("cuda", 0)
model, _ = bench.random_state_dict()
model = model.to(dev).eval()
x = torch.rand(bench.BATCH, 3, bench.SIDE, bench.SIDE, generator=torch.Generator().manual_seed(1))
x = ((x * 255).to(torch.uint8) if os.environ.get("YFV2_PROF_F32") is None else x).to(dev)      # uint8 input like bench.py's default
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    preds = model(x)
    out = eng.decode_nms(preds, bench.cfg(), bench.CONF, bench.IOU)
torch.cuda.synchronize()
print("done", int(out[1].sum()))
