"""Per-phase clock64 breakdown of decode_nms_kernel on the bench workload (yfv2_debug_nms_profile)."""
import os, sys; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ctypes, json, torch
import bench, yfv2, yfv2_engine as eng
dev = torch.device("cuda", 0)
model, _ = bench.random_state_dict()
model = model.to(dev).eval()
x = torch.rand(bench.BATCH, 3, bench.SIDE, bench.SIDE, generator=torch.Generator().manual_seed(1))
x = (x * 255).to(torch.uint8).to(dev)
preds = model(x)
buf = torch.zeros(bench.BATCH, 16, dtype=torch.int64, device=dev)
lib = eng.lib()
lib.yfv2_debug_nms_profile(ctypes.c_void_p(buf.data_ptr()))
for _ in range(2):
    out = eng.decode_nms(preds, bench.cfg(), bench.CONF, bench.IOU)
torch.cuda.synchronize()
lib.yfv2_debug_nms_profile(None)
b = buf.cpu().double()
names = ["candidates", "sort", "chunk_load", "vs_kept", "in_chunk", "resolve", "append", "tail"]
tot = b[:, :8].sum(1)
res = {"clock_mhz_assumed": 1965, "images": int(b.shape[0]), "mean_total_us": float(tot.mean() / 1965), "max_total_us": float(tot.max() / 1965),
       "phases_mean_us": {n: round(float(b[:, i].mean() / 1965), 2) for i, n in enumerate(names)},
       "chunks_mean": float(b[:, 8].mean()), "candidates_mean": float(b[:, 9].mean()), "kept_mean": float(b[:, 10].mean())}
print(json.dumps(res))
