import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ctypes, torch, bench, yfv2, yfv2_engine as eng
dev = torch.device("cuda", 0)
model, _ = bench.random_state_dict(); model = model.to(dev).eval()
x = torch.rand(256, 3, 352, 352).to(dev)
plan = model._plan_for(x); preds = plan.alloc_preds()
plan.forward(x, preds); torch.cuda.synchronize()
L = eng.lib(); L.yfv2_debug_timestamps.argtypes=[ctypes.c_void_p, ctypes.c_int]
names = plan.stage_names
for target in ("stage3.3", "stage2.2"):
    st = names.index(target)
    buf = (ctypes.c_longlong*64)()
    L.yfv2_debug_timestamps(None, 1)
    plan.forward_range(x, preds, st, st+1); torch.cuda.synchronize()
    L.yfv2_debug_timestamps(buf, 0)
    t = list(buf)
    t0 = t[0]
    print(target, "phase marks (cycles from item start): loads_start %d loads_done %d  B_done %d  after_sync %d  C_done %d  loop_end %d teardown %d" % tuple(t[i]-t0 for i in (1,2,3,4,5,6,7)))
    for k in range(min(8, t[62])):
        a,b,c = t[16+3*k:19+3*k]
        print("   tile %d: reach_dfull_wait %d  dfull_ready %d (+%d)  epilogue_done %d (+%d)" % (k, a-t0, b-t0, b-a, c-t0, c-b))
