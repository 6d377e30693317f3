import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
print("H2D GB/s", n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9)
print("D2H GB/s", n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidirectional GB/s total", 2 * n / t(both) / 1e9)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("lscpu | head -20; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)

# The end-to-end evaluation moves 24 B in and 32 B out per query: the same copies with no kernels at
# all, in the chunking of the library's host pipe (2^22 queries), give the ceiling of any implementation.
q, chunk = 1 << 28, 1 << 22
hin = torch.empty(q * 24, dtype=torch.uint8, pin_memory=True); hout = torch.empty(q * 32, dtype=torch.uint8, pin_memory=True)
din = torch.empty(chunk * 24 * 3, dtype=torch.uint8, device="cuda"); dout = torch.empty(chunk * 32 * 3, dtype=torch.uint8, device="cuda")
def copies_only():
    for c in range(q // chunk):
        k = c % 3
        with torch.cuda.stream(s1):
            din[k * chunk * 24:(k + 1) * chunk * 24].copy_(hin[c * chunk * 24:(c + 1) * chunk * 24], non_blocking=True)
        with torch.cuda.stream(s2):
            hout[c * chunk * 32:(c + 1) * chunk * 32].copy_(dout[k * chunk * 32:(k + 1) * chunk * 32], non_blocking=True)
sec = t(copies_only, reps=2)
print("copies-only ceiling for 24 B in + 32 B out per query: %.3f s per 2^28 queries = %.0f Mpts/s" % (sec, q / sec / 1e6))
