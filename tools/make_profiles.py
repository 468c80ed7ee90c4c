"""Turn the raw ncu artefacts under gpurun_out/ into the committed summaries under profiles/ (run on the CPU box).

  gpurun_out/fused_full.ncu-rep   (ncu --set full --import-source on, one launch of k_fused_assign_update)
  gpurun_out/launches.csv         (ncu --metrics gpu__time_duration.sum launch list of a short bench.py run)
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = os.path.join(ROOT, "gpurun_out", "fused_full.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
N_ROWS, D = int(os.environ.get("B2K_PROF_ROWS", 10_000_000)), 128

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def f(name):
    v, u = m[name]
    return float(v.replace(",", "")), u
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
summary = {}
for k in keep:
    if k in m:
        summary[k] = {"value": f(k)[0], "unit": f(k)[1]}
def to_bytes(name):
    v, u = f(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
js = {"kernel": "k_fused_assign_update<64,128>", "rows": N_ROWS, "d": D, "dram_bytes_per_launch": dram,
      "dram_bytes_per_row": dram / N_ROWS, "algorithmic_bytes_per_launch": 4.0 * N_ROWS * D, "metrics": summary,
      "source": f"ncu --set full --clock-control none -k regex:k_fused -s 3 -c 1 python bench.py --steps 4 --warmup 3 (tag {tag})"}
json.dump(js, open(os.path.join(out_dir, "fused_kernel_ncu_summary.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"{tag}_fused_kernel_ncu.md"), "w") as o:
    o.write(f"# {tag}: ncu --set full, k_fused_assign_update<64,128>, cfg2 (n=10M, d=128, k=64), one launch\n\n")
    o.write("Per-launch times under ncu are cold-cache and serialised (compare shares, not absolutes).\n\n| metric | value | unit |\n|---|---|---|\n")
    for k, v in summary.items():
        o.write(f"| `{k}` | {v['value']:.6g} | {v['unit']} |\n")
    o.write(f"\nDRAM traffic per launch = {dram/1e9:.4f} GB vs algorithmic 4*n*d = {4.0*N_ROWS*D/1e9:.4f} GB "
            f"(ratio {dram/(4.0*N_ROWS*D):.4f}: X is read exactly once)\n")
    # SASS evidence
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "spark_rapids_ml_b200", "libb2kmeans.so")], capture_output=True, text=True).stdout
    import re, collections
    cnt = collections.Counter(re.findall(r"\b(UTCHMMA|UTMALDG\.2D|UTMAPF[.A-Z0-9]*|LDTM\.x\d+|STTM\.x\d+|UTCBAR|SYNCS\.[A-Z0-9.]+|HMMA[.A-Z0-9]*)\b", sass))
    o.write("\nSASS mnemonics in libb2kmeans.so (all instantiations): " + ", ".join(f"{k}×{v}" for k, v in sorted(cnt.items())) + "\n")
lp = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lp):
    agg = {}
    rd = csv.reader(l for l in open(lp) if not l.startswith("=="))
    h = next(rd)
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    for r in rd:
        if len(r) <= vi: continue
        try: t = float(r[vi].replace(",", ""))
        except ValueError: continue
        a = agg.setdefault(r[ki][:110], [0.0, 0])
        a[0] += t; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    ours = {k: a for k, a in agg.items() if "k_" in k}
    tot_ours = sum(a[0] for a in ours.values())
    with open(os.path.join(out_dir, f"{tag}_launch_list.md"), "w") as o:
        o.write(f"# {tag}: ncu launch list (gpu__time_duration.sum) of `python bench.py --steps 4 --warmup 3`\n\n")
        o.write("Library kernels only (torch kernels of the synthetic-data generator excluded from the share):\n\n| kernel | launches | total us | share of library time |\n|---|---|---|---|\n")
        for k, a in sorted(ours.items(), key=lambda kv: -kv[1][0]):
            o.write(f"| `{k}` | {a[1]} | {a[0]/1e3:.1f} | {100*a[0]/tot_ours:.1f} % |\n")
print(json.dumps(js)[:400])
