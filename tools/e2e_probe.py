"""H2D bandwidth and e2e variants (streams / buffers) for yfv2_detect_u8_host."""
import os, sys; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, time
import bench, yfv2, yfv2_engine as eng
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model, _ = bench.random_state_dict()
model = model.to(dev).eval()
g = torch.Generator().manual_seed(1)
B, S = bench.BATCH, bench.SIDE
x = (torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8).pin_memory()
xd = torch.empty_like(x, device=dev)
for _ in range(3): xd.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): xd.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s", 20 * x.numel() / (e0.elapsed_time(e1) / 1e3) / 1e9)
def run(nbuf, steps=30):
    plans = [eng.Plan(dev, B, S, S, bench.ANCHORS, bench.CLASSES, detect_max_det=eng.MAX_DET) for _ in range(nbuf)]
    params, bn = model._weight_tensors()
    for p_ in plans: p_.pack(params, bn)
    xs = [(torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8).pin_memory() for _ in range(nbuf)]
    outs = [torch.empty((B, eng.MAX_DET, 6), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    cnts = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(nbuf)]
    streams = [torch.cuda.Stream(dev) for _ in range(nbuf)]
    anc = eng.anchors_array(bench.cfg())
    def step(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            plans[b].detect_u8_host(xs[b], anc, bench.CONF, bench.IOU, outs[b], cnts[b])
    for i in range(4): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps): step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("nbuf", nbuf, "img/s", B * steps / dt)
for nb in (1, 2, 3): run(nb)
