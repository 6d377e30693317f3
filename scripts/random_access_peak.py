"""Measured throughput of random 32-byte sector writes/reads on this GPU (the bound of the
query sort and of the result scatter): out[perm[i]] = v[i] with 32-byte rows."""
import torch, json, sys
Q = 1 << 26
perm = torch.randperm(Q, device="cuda")
def tm(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
res = {}
for width in (4, 8):   # doubles per row: 32 B and 64 B
    v = torch.rand((Q, width), dtype=torch.float64, device="cuda")
    o = torch.empty_like(v)
    ms_w = tm(lambda: o.index_copy_(0, perm, v))
    ms_r = tm(lambda: torch.index_select(v, 0, perm, out=o))
    res["row_%dB" % (8 * width)] = {"scatter_ms": ms_w, "scatter_Grows_s": Q / ms_w / 1e6,
                                    "gather_ms": ms_r, "gather_Grows_s": Q / ms_r / 1e6}
    del v, o
print(json.dumps(res, indent=1))
