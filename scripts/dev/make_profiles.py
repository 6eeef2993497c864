"""Turn the scratch output of scripts/collect_1gpu.sh (gpurun_out/<prefix>_*) into the committed summaries under
profiles/: bench lines as JSON, the ncu launch list, a table of the `ncu --set full` capture of the half-step kernels
and the DRAM-traffic record bench.py reads (tied to the source hash of the kernels it was taken from).
    python scripts/dev/make_profiles.py r02
"""
import csv, json, os, shutil, subprocess, sys
sys.path.insert(0, ".")
import bench

P = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, O = "gpurun_out", "profiles"

def last_json(path):
    out = None
    for l in open(path):
        if l.startswith("{"):
            out = l.strip()
    return out

for name in os.listdir(G):
    if name.startswith(P + "_bench_") and name.endswith(".log"):
        j = last_json(os.path.join(G, name))
        if j:
            open(os.path.join(O, name[:-4] + ".json"), "w").write(j + "\n")
for name in (P + "_ncu_launches.csv", P + "_pytest_gpu.log", P + "_parity.json"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(O, name))

rep = os.path.join(G, P + "_prof_half_step.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg"]
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(O, P + "_ncu_half_step_raw.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])

def val(r, k):
    i = hdr.index(k)
    x = float(r[i].replace(",", ""))
    u = units[i]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

# capture order (scripts/collect_1gpu.sh): user dual, user primal, (item: empty dual), item primal, (empty dual), item primal
names = ["user half-step, dual kernel", "user half-step, primal kernel", None, "item half-step, primal kernel"]
kern = {}
for r, nm in zip(data, names):
    if nm is None:
        continue
    kern[nm] = {"ms": val(r, "gpu__time_duration.sum"), "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "ipc": val(r, "sm__inst_executed.avg.per_cycle_active"), "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "eligible_warps_per_scheduler": val(r, "smsp__warps_eligible.avg.per_cycle_active"),
                "smem_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"), "registers": val(r, "launch__registers_per_thread")}
traffic = {"source_hash": bench.source_hash(), "epoch_dram_bytes": sum(k["dram_bytes"] for k in kern.values()),
           "what": "dram__bytes_read.sum + dram__bytes_write.sum of the three half-step kernels of one epoch (steady-state factors, "
                   "ML-20M shape), ncu --set full --clock-control none, scripts/collect_1gpu.sh",
           "kernels": kern}
json.dump(traffic, open(os.path.join(O, P + "_ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(O, P + "_ncu_half_step.md"), "w") as fh:
    fh.write(f"# ncu --set full, half-step kernels, ML-20M shape, factors after 5 epochs ({P})\n\n")
    fh.write("Command: `ncu --set full --clock-control none --import-source on -k regex:\"als_half_step_(dual|tc)_kernel\" -s 22 -c 6 "
             "python scripts/tc_profile.py --steady` (scripts/collect_1gpu.sh). Per-launch times are serialised and cold-cache; "
             "the bench line's CUDA-event times are the reference. Raw columns: `" + P + "_ncu_half_step_raw.csv`.\n\n")
    fh.write("| kernel | ms | DRAM MB (r+w) | L2 hit % | tensor pipe % | IPC | issue active % | eligible warps / scheduler | smem wavefronts % of peak | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for nm, k in kern.items():
        fh.write(f"| {nm} | {k['ms']/1e6 if k['ms'] > 1e3 else k['ms']:.3f} | {k['dram_bytes']/1e6:.0f} | {k['l2_hit_pct']:.1f} | {k['tensor_pipe_pct']:.1f} | {k['ipc']:.2f} | "
                 f"{k['issue_active_pct']:.1f} | {k['eligible_warps_per_scheduler']:.2f} | {k['smem_wavefronts_pct']:.1f} | {k['registers']:.0f} |\n")
    fh.write(f"\nDRAM traffic of the three kernels: {traffic['epoch_dram_bytes']/1e9:.2f} GB per epoch against 20.9 GB algorithmic "
             "(SURVEY.md 8d): the 85 MB of factors and the whitened copy stay in L2; the item side's traffic is the parked "
             "partial Grams of split rows.\n")
print(json.dumps(traffic, indent=1)[:1200])
