"""Turn one evidence run's ncu outputs (gpurun_out/) into the tracked summaries under profiles/ (tools only).
usage: make_profile_summary.py TAG   (reads gpurun_out/raw_TAG.csv from `ncu -i prof_TAG.ncu-rep --page raw --csv`
and gpurun_out/TAG_launches.csv; writes profiles/TAG_ncu_full_summary.md, TAG_launches_summary.md, ncu_traffic.json)"""
import collections, csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_source_hash():   # bench.py quotes ncu_traffic.json only while this still matches the sources it runs
    h = hashlib.sha1()
    d = os.path.join(ROOT, "fea-large_b200", "csrc")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()



tag = sys.argv[1]
NOTES = {
    "element_kernel": "interleaved staging, stores straight from registers; DRAM writes at ~4.4 TB/s are now the busy unit (K_e staging 4.0 GB + F, sigma 0.72 GB + R_e 0.24 GB); 10 warps/SM",
    "gather_blocks_kernel [plain]": "upper triangle only + transposed store into the mirror slot: every staged block is read once; DRAM (read + write ~4.8 TB/s) and L2 are the busy units",
    "gather_blocks_kernel [Dirichlet flags folded in]": "what a bench step runs",
    "spmv_sell_kernel": "algorithmic 2.96 GB",
}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"raw_{tag}.csv"))))
hdr, units = rows[0], rows[1]
def val(r, m): return r[hdr.index(m)]
def scaled(r, m, table):
    return float(val(r, m)) * table[units[hdr.index(m)]]
picked, n_gather = collections.OrderedDict(), 0
for r in rows[2:]:
    name = val(r, "Kernel Name").split("(")[0].replace("void ", "").split("<")[0]
    key = name
    if name == "gather_blocks_kernel":   # profile_target.py: plain launches first, then the bench step's fused ones
        key = name + (" [plain]" if n_gather == 0 else " [Dirichlet flags folded in]")
        n_gather += 1
    picked[key] = r    # last launch of each kind
md = [f"# ncu --set full, kernels of `{tag}`: C3 = Kuhn 55^3, 998 250 tets, 4.10 M DOF, A5, one B200", "",
      "`ncu --set full --clock-control none --import-source on -k regex:... -c 14` on `python tools/profile_target.py 55`,",
      "after the same command exited 0 without ncu (the .ncu-rep itself stays in gpurun_out/, scratch).", "",
      "| kernel | time | DRAM read + write | L1TEX | L2 | FP64 pipe | warps active | note |", "|---|---|---|---|---|---|---|---|"]
traffic = {}
T = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
B = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
for key, r in picked.items():
    t = scaled(r, "gpu__time_duration.sum", T)
    rd, wr = scaled(r, "dram__bytes_read.sum", B), scaled(r, "dram__bytes_write.sum", B)
    pct = lambda m: float(val(r, m))
    md.append(f"| `{key}` | {t:.3f} ms | {rd:.2f} + {wr:.2f} GB | {pct('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} % | "
              f"{pct('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} % | {pct('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.0f} % | "
              f"{pct('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} % | {NOTES.get(key, '')} |")
    traffic[key] = (rd + wr) * 1e9
md.append("")
for key, r in picked.items():
    md.append(f"### {key}")
    md += [f"- {m}: {val(r, m)} {units[hdr.index(m)]}" for m in WANT if m in hdr]
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(md) + "\n")
json.dump({"source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, one launch each, C3 = Kuhn 55^3 on one B200)",
           "workload": {"n": 55, "n_gpus": 1}, "csrc_sha1": kernel_source_hash(),
           "dram_bytes_per_launch": {"element_kernel": traffic["element_kernel"],
                                     "gather_blocks_kernel": traffic["gather_blocks_kernel [Dirichlet flags folded in]"],
                                     "gather_residual_kernel": traffic.get("gather_residual_kernel"),
                                     "spmv_sell_kernel": traffic["spmv_sell_kernel"]}},
          open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)

rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv"))) if len(r) > 10]
h = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[h.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("fea::", ""), [0, 0.0])
    a[0] += 1
    a[1] += float(r[h.index("Metric Value")]) / 1e6
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list summary ({tag}_launches.csv): `bench.py --steps 3 --warmup 3 --newton-iters 1 --lin-max-iter 60 --no-cpu-baseline`", "",
       "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.", "",
       "| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
out += [f"| {k} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.4f} |" for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(md[5:14]))
print("\n".join(out[4:12]))
