"""Turn the raw ncu artefacts under gpurun_out/ into the committed summaries under profiles/ (run on the CPU box).

  python tools/make_profiles.py <tag> <cfg>        cfg in {cfg2, cfg3}
    gpurun_out/fused_full.ncu-rep / fused_t_full.ncu-rep   (ncu --set full --import-source on, one launch)
    gpurun_out/launches_<cfg>.csv                          (ncu --metrics gpu__time_duration.sum launch list)
Writes profiles/<tag>_<cfg>_ncu_summary.json (read by bench.py for roofline.traffic), <tag>_<cfg>_fused_kernel_ncu.md,
<tag>_<cfg>_launch_list.md and <tag>_<cfg>_source_hotspots.txt.
"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
SPEC = {"cfg2": dict(rep="fused_full.ncu-rep", kernel="k_fused_assign_update<64,128,PAIR,0> (3xTF32)", n=10_000_000, d=128, k=64),
        "cfg3": dict(rep="fused_t_full.ncu-rep", kernel="k_fused_t<8,true> (1xTF32 screening; near-tie rows deferred to k_fix_labels_t / k_fix_accum_t)", n=12_500_000, d=256, k=256)}[cfg]
rep = os.path.join(ROOT, "gpurun_out", SPEC["rep"])
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
N_ROWS, D, K = SPEC["n"], SPEC["d"], SPEC["k"]

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
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
summary = {}
for k in keep:
    if k in m:
        try:
            summary[k] = {"value": f(k)[0], "unit": f(k)[1]}
        except ValueError:
            pass
def to_bytes(name):
    v, u = f(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
alg = 4.0 * N_ROWS * D
js = {"kernel": SPEC["kernel"], "config": cfg, "rows": N_ROWS, "d": D, "k": K, "dram_bytes_per_launch": dram,
      "dram_bytes_per_row": dram / N_ROWS, "algorithmic_bytes_per_launch": alg, "metrics": summary,
      "source": f"ncu --set full --clock-control none --import-source on -s 4 -c 1 python bench.py --config {cfg} ... (tools/gpu_ncu.sh, tag {tag})"}
json.dump(js, open(os.path.join(out_dir, f"{tag}_{cfg}_ncu_summary.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"{tag}_{cfg}_fused_kernel_ncu.md"), "w") as o:
    o.write(f"# {tag}: ncu --set full, {SPEC['kernel']}, {cfg} (n={N_ROWS}, d={D}, k={K}), one launch\n\n")
    o.write("Per-launch times under ncu are cold-cache and serialised (compare shares, not absolutes).\n\n| metric | value | unit |\n|---|---|---|\n")
    for k, v in summary.items():
        o.write(f"| `{k}` | {v['value']:.6g} | {v['unit']} |\n")
    o.write(f"\nDRAM traffic per launch = {dram/1e9:.4f} GB vs algorithmic 4*n*d = {alg/1e9:.4f} GB (ratio {dram/alg:.4f})\n")
    so = os.path.join(ROOT, "spark_rapids_ml_b200", "libb2kmeans.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cnt = collections.Counter(re.findall(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.2D(?:\.2CTA)?|UTMAPF[.A-Z0-9]*|LDTM(?:\.16dp256bit)?\.x\d+|STTM\.x\d+|UTCBAR[.A-Z0-9]*|STAS[.0-9]*|USETMAXREG[.A-Z_]*|HMMA[.A-Z0-9]*)\b", sass))
    o.write("\nSASS mnemonics in libb2kmeans.so (all instantiations): " + ", ".join(f"{k}x{v}" for k, v in sorted(cnt.items())) + "\n")
# source hot spots (needs -lineinfo): top lines by sampled stalls
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
try:
    rd = list(csv.reader(io.StringIO(src)))
    h = rd[0]
    ci = {n: i for i, n in enumerate(h)}
    samp = next((n for n in h if n.startswith("# Samples") or n == "Warp Stall Sampling (All Samples)" or "Samples" in n), None)
    srcc = next((n for n in h if n == "Source"), None)
    if samp and srcc:
        rows_ = []
        for r in rd[1:]:
            try: rows_.append((float(r[ci[samp]].replace(",", "") or 0), r[ci[srcc]][:150]))
            except Exception: pass
        tot = sum(x[0] for x in rows_) or 1
        with open(os.path.join(out_dir, f"{tag}_{cfg}_source_hotspots.txt"), "w") as o:
            o.write(f"# {tag} {cfg}: top SASS/source lines of {SPEC['kernel']} by warp-stall samples ({samp}); total {tot:.0f}\n")
            for s_, line in sorted(rows_, key=lambda x: -x[0])[:40]:
                o.write(f"{100*s_/tot:6.2f} %  {line}\n")
except Exception as ex:
    print("source page skipped:", ex)
lp = os.path.join(ROOT, "gpurun_out", f"launches_{cfg}.csv")
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
    ours = {k: a for k, a in agg.items() if "k_" in k}
    tot_ours = sum(a[0] for a in ours.values())
    with open(os.path.join(out_dir, f"{tag}_{cfg}_launch_list.md"), "w") as o:
        o.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --config {cfg} --steps 3 --warmup 3 ...`\n\n")
        o.write("Library kernels (regex k_); per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total us | mean us | share of library time |\n|---|---|---|---|---|\n")
        for k, a in sorted(ours.items(), key=lambda kv: -kv[1][0]):
            o.write(f"| `{k}` | {a[1]} | {a[0]/1e3:.1f} | {a[0]/1e3/a[1]:.1f} | {100*a[0]/tot_ours:.1f} % |\n")
print(json.dumps({k: js[k] for k in ("kernel", "dram_bytes_per_launch", "algorithmic_bytes_per_launch")}), {k: v["value"] for k, v in summary.items()})
