"""End-to-end (pinned host buffers through the C ABI) value+gradient time against batch size."""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import bsplineinterpolation_b200 as B
    n = 256
    t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3)
    fn = t.interpolate(torch.rand((n, n, n), dtype=torch.float64, device="cuda"))
    for lq in (16, 18, 20, 22, 24, 26):
        q = 1 << lq
        hp = torch.rand((q, 3), dtype=torch.float64).pin_memory(); ho = torch.empty((q, 4), dtype=torch.float64).pin_memory()
        a, b = hp.numpy(), ho.numpy()
        fn.value_grad(a, out=b); fn.value_grad(a, out=b)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps): fn.value_grad(a, out=b)
        dt = (time.perf_counter() - t0) / reps
        print("q=2^%d: %.3f ms  %.1f Mpts/s" % (lq, dt * 1e3, q / dt / 1e6), flush=True)
else:
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
