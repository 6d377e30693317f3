"""Turn raw ncu outputs under gpurun_out/ into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(P, exist_ok=True)

# 1. launch list -> per-kernel shares
rows = [r for r in csv.reader(open(os.path.join(G, "%s_launches_bench.csv" % tag))) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
launches = [(r[ii], r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
with open(os.path.join(P, "%s_launches_bench.csv" % tag), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --no-cpu\n")
    fh.write("# per-launch device time (ns); cold-cache, serialised: compare shares, not absolutes\nid,kernel,ns\n")
    for i, k, v in launches:
        fh.write('%s,"%s",%.0f\n' % (i, k, v))
tot = {}
for _, k, v in launches:
    name = k.split("(")[0].split("::")[-1]
    tot.setdefault(name, [0, 0.0]); tot[name][0] += 1; tot[name][1] += v
mine = {k: v for k, v in tot.items() if any(s in k for s in ("kernel",)) and "at::" not in k}
with open(os.path.join(P, "%s_launch_shares.txt" % tag), "w") as fh:
    s = sum(v[1] for v in mine.values())
    fh.write("kernel, launches, total_ms, share_of_our_kernels\n")
    for k, (n, v) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
        fh.write("%s, %d, %.3f, %.1f%%\n" % (k, n, v / 1e6, 100 * v / s))
# the device-resident steps alone (warm-up + timed: everything up to the 5th full-size evaluation kernel); the rest of
# the list is the end-to-end leg's 2^22-query chunks, whose kernels have other proportions
big, seen = {}, 0
for _, k, v in launches:
    name = k.split("(")[0].split("::")[-1]
    if not any(t in name for t in ("key_count", "tile_totals", "plan_kernel", "key_cursor", "scatter_kernel", "eval_binned")):
        continue
    big.setdefault(name, [0, 0.0]); big[name][0] += 1; big[name][1] += v
    if "eval_binned" in name:
        seen += 1
        if seen == 5:
            break
with open(os.path.join(P, "%s_launch_shares.txt" % tag), "a") as fh:
    s5 = sum(v[1] for v in big.values())
    fh.write("\n# device-resident steps only (first 5 evaluate calls of 2^28 queries): kernel, launches, ms per step, share of the step\n")
    for k, (n, v) in sorted(big.items(), key=lambda kv: -kv[1][1]):
        fh.write("%s, %d, %.3f, %.1f%%\n" % (k, n, v / 1e6 / 5, 100 * v / s5))
print(open(os.path.join(P, "%s_launch_shares.txt" % tag)).read())

# 2. full-set summaries
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg']
traffic = {}
for rep in ("%s_eval_pipeline" % tag, "%s_solve" % tag, "%s_fields_contract" % tag):
    path = os.path.join(G, rep + ".ncu-rep")
    raw = os.path.join(G, rep + ".raw.csv")   # exported on the GPU box when the report itself is too large to bring back
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(path):
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    else:
        continue
    rr = list(csv.reader(out.splitlines())); h, u = rr[0], rr[1]
    with open(os.path.join(P, rep + "_summary.txt"), "w") as fh:
        fh.write("# ncu --set full --clock-control none (one launch each; raw page exported on the GPU box)\n")
        for r in rr[2:]:
            name = r[h.index('Kernel Name')]
            fh.write("== %s\n" % name)
            for w in WANT:
                if w in h:
                    fh.write("   %-72s %s %s\n" % (w, r[h.index(w)], u[h.index(w)]))
            def val(m):
                x = float(r[h.index(m)].replace(",", "")); un = u[h.index(m)]
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(un, 1)
            traffic[name.split("(")[0].split("::")[-1].split("<")[0]] = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
    print(open(os.path.join(P, rep + "_summary.txt")).read())
q = 1 << int(os.environ.get("PROFILE_LOG2_Q", "28"))
ev = sum(traffic.get(k, 0) for k in ("eval_binned_kernel", "scatter_kernel", "key_count_kernel"))
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"eval_bytes_per_query": ev / q if ev else None, "queries_in_capture": q, "commit": commit, "round": tag,
           "kernels": {k: v for k, v in traffic.items()},
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of key_count + scatter + eval_binned, one launch each, "
                   "ncu --set full --clock-control none, %d queries, 256^3 cubic fp64 value+gradient; solve kernels: 512^3" % q},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "traffic.json")).read())
