import os, sys, subprocess
for mb in (4, 16, 24, 40, 64, 96, 160, 4096):
    env = dict(os.environ, BSPL_SWEEP_L2_BYTES=str(mb << 20))
    out = subprocess.run([sys.executable, "scripts/quick_bench.py", "solve"], env=env, capture_output=True, text=True).stdout
    print("== L2 budget", mb, "MB"); print(out, flush=True)
