"""Turns the round-2 raw captures in gpurun_out/ into the committed summaries under profiles/ (run in the build
container; ncu is only used to READ the .ncu-rep)."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEYS = ['L0.spmv', 'L0.jacobi', 'L0.resid+restrict', 'L0.prolong', 'L0.spmv+dot', 'L0.jacobi+dot', 'L1.spmv', 'L1.jacobi',
        'L1.resid+restrict', 'L1.prolong']


def sweep_table(path, title, note):
    d = json.load(open(path))
    out = [f"## {title}", "", note, "", "| config | solve ms | V-cycle ms | " + " | ".join(KEYS) + " | SELL fill A per level |", "|---|---|---|" + "---|" * (len(KEYS) + 1)]
    for name, row in d.items():
        out.append(f"| {name} | {row['solve_ms']:.2f} | {row['vcycle_ms']:.3f} | " +
                   " | ".join(f"{row[k]['ms']:.4f} / {row[k]['frac']:.2f}" for k in KEYS) + f" | {row['fill']} |")
    return "\n".join(out) + "\n"


def kernel_sweep():
    txt = ["# r02 kernel sweep — Poisson 256^3, 1 GPU, L2 flushed before every launch (ms / fraction of the measured 6548 GB/s)",
           "", "`python scripts/kernel_sweep.py 256 out.json <configs>` under gpurun; each config re-uploads the same hierarchy with the",
           "environment switches of `scripts/kernel_sweep.py: ENVS`. `auto` = the shipped defaults. The solve has 22 iterations (random rhs).", ""]
    txt.append(sweep_table(os.path.join(G, "r2_sweep256_b.json"), "Final defaults and the switches that stayed off",
                           "auto-nolong: CSR-stream phase B with one thread per row (round 1); auto-renum/-w512/-w1024: upload-time renumbering of the "
                           "coarse rows by length in windows of 4096/512/1024; auto-psig*: sorting window of P; auto-rsig4096: R sorted; auto-sort1.1: "
                           "run-time row sorting from 1.1x padding; auto-p1: short-row SELL instantiation <2,ADD,U=2,5 CTAs/SM> for P."))
    txt.append(sweep_table(os.path.join(G, "r2_sweep256_c.json"), "L2 prefetch two slices ahead (PAMG_SELL_PF bit mask: 1 P, 2 long-row A, 4 R)",
                           "Persistent launches with `prefetch.global.L2` of the slice a warp reaches two iterations later. P0 gains 2 %; the long-row "
                           "operators lose 30-60 % because they have to run persistent for it."))
    txt.append(sweep_table(os.path.join(G, "r2_sweep256_a.json"), "First sweep of the round (code with the role code inlined into k_spmv_sell: 80 registers in three instantiations)",
                           "Kept because it shows the cost of that register change (auto-r1 = round-1 settings on that build: L0.prolong 0.53 ms "
                           "instead of 0.24, L0.jacobi 0.326 instead of 0.309) -- the unified-role code now lives in its own kernel."))
    txt += ["## Reading", "",
            "* Production kernels are unchanged from round 1 (72 registers, same SASS loop): L0 SpMV 0.99, Jacobi 0.99, residual+restriction 0.92.",
            "* Long-row phase B (16 lanes per row) lifts L1 residual+restriction 0.218 -> 0.210 ms; it is on by default.",
            "* Renumbering removes the padding of A1 (fill 1.159 -> 1.006 / 1.029) but the level-1 sweeps gain only 2-6 % and P0, whose columns are",
            "  renumbered with it, loses 18-30 us: net zero on the V-cycle (1.500 vs 1.501 ms). Off by default.",
            "* P0 (prolongation + correction on the fine level, 0.69) did not respond to more resident warps (48 registers, 5 CTAs/SM: 0.26 ms),",
            "  to an L2 prefetch (0.237), to smaller sorting windows (0.28-0.29) or to no sorting (0.2415 with 41 % padding). ncu: DRAM 56 %, L1 50 %,",
            "  L2 48 %, issue 20 %, long-scoreboard 24 per issue -- no unit is saturated; the slice chain (extents -> perm -> entries -> gather -> store)",
            "  is simply short of independent work per warp at 3.6 entries per row.", ""]
    open(os.path.join(P, "r02_kernel_sweep.md"), "w").write("\n".join(txt))


def traffic():
    rep = os.path.join(G, "r2_traffic.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    plain = [json.loads(l) for l in open(os.path.join(G, "r2_traffic_plain.log")) if l.startswith("{")]
    launches = []
    for rec in plain:   # two launches per record, SELL kernels only (the capture filter)
        if rec["format"] == 3:
            launches += [rec, rec]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "launch__registers_per_thread"]
    idx = {w: hdr.index(w) for w in want}
    units = rows[1]

    def val(r, w):
        v = float(r[idx[w]].replace(",", ""))
        u = units[idx[w]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    out = ["# r02 ncu --set full: DRAM traffic of the SELL kernels on three BASELINE operators (1 GPU, L2 flushed before each launch)", "",
           "Command: `python scripts/profile_traffic.py 2 && ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -o r2_traffic python scripts/profile_traffic.py 2`",
           "(gpurun call 6; report gpurun_out/r2_traffic.ncu-rep, 54 MB, not committed).  traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch;",
           "algorithmic = DESIGN.md section 4 (fp64 values, int32 columns, every vector element once).", "",
           "| workload | level | op | kernel | us | DRAM read | DRAM written | traffic / algorithmic | DRAM thr % | warps active % | L1 hit % | L2 hit % | regs |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    tj = {}
    body = rows[2:]
    for k, r in enumerate(body):
        if k >= len(launches):
            break
        rec = launches[k]
        name = re.search(r"k_spmv_sell<[^>]*>", r[hdr.index("Kernel Name")]).group(0)
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        if k % 2 == 1:   # second launch of each pair
            out.append(f"| {rec['workload']} | {rec['level']} | {rec['kind']} | `{name}` grid {r[hdr.index('Grid Size')]} | {float(r[idx['gpu__time_duration.sum']]):.1f} | "
                       f"{rd / 1e9:.3f} GB | {wr / 1e6:.1f} MB | {(rd + wr) / rec['algorithmic_bytes']:.3f} | {float(r[idx[want[3]]]):.1f} | {float(r[idx[want[4]]]):.1f} | "
                       f"{float(r[idx[want[5]]]):.1f} | {float(r[idx[want[6]]]):.1f} | {r[idx[want[7]]]} |")
            if rec["level"] == 0 and rec["kind"] == "spmv":
                tj[rec["workload"]] = {"1": dict(spmv_A0_dram_bytes_per_launch=int(rd + wr), algorithmic_bytes=rec["algorithmic_bytes"],
                                                 source="profiles/r02_ncu_traffic.md (ncu --set full, round 2, gpurun call 6)")}
    out += ["", "Reading: on all three operators the fine-level SpMV moves 0.94-1.00x its algorithmic bytes -- x is served by L1/L2 (every x element crosses",
            "the DRAM pins at most once), so there is nothing for a shared-memory staging of x or an L2 persisting window to recover (SURVEY 8 g2).",
            "The only launches above 1.0 are the level-1 Poisson / jump sweeps (SELL padding 1.159 / 1.036, see r02_kernel_sweep.md for the renumbering experiment)."]
    open(os.path.join(P, "r02_ncu_traffic.md"), "w").write("\n".join(out) + "\n")
    tj["_format"] = "workload -> number of GPUs -> capture; bench.py reports roofline.traffic only for an exact (workload, N) match, else null"
    json.dump(tj, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)


def sass_and_regs():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    os.chdir(ROOT)
    import sass_extract
    fn = sass_extract.functions()
    out = ["# r02 SASS of the SELL entry loop (cuobjdump -sass parallel_amg_b200/libpamg.so; skeleton = memory / control / fp64 lines)", "",
           "DESIGN.md 4.1 argues from this schedule: per step the four 64-bit column loads and the four 128-bit value loads (LDG.E.EF = ld.global.cs)",
           "are issued back to back, then the eight x gathers (LDG.E.64.CONSTANT), then the DMUL/DADD chain in column order; nothing else is live.", ""]
    for want in ("k_spmv_sell<2, 0, false, 4, 3, 0>", "k_spmv_sell<2, 2, false, 4, 3, 0>"):
        for name, body in fn.items():
            if want in name:
                lines = []
                for line in body:
                    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                    if m and re.search(r"LDG|STG|LDS|STS|BAR|BRA|DADD|DMUL|DFMA|MEMBAR|ATOM|CCTL|EXIT", m.group(2)):
                        lines.append(f"{m.group(1)}  {m.group(2).strip()}")
                # the entry loop = the window around the LDG.E.EF.128 loads
                k128 = [i for i, l in enumerate(lines) if "LDG.E.EF.128" in l]
                lo, hi = max(0, k128[0] - 8), min(len(lines), k128[-1] + 34)
                out += [f"## {name}  ({len(body)} SASS lines; loop excerpt)", "```"] + lines[lo:hi] + ["```", ""]
    open(os.path.join(P, "r02_sass_sell_loop.md"), "w").write("\n".join(out))
    tab = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ptxas_table.py")], capture_output=True, text=True).stdout
    open(os.path.join(P, "r02_ptxas_registers.txt"), "w").write(
        "# ptxas -v of every kernel in libpamg.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3); CTAs/SM = min over registers, 2048 threads, shared memory\n" + tab)


if __name__ == "__main__":
    kernel_sweep()
    traffic()
    sass_and_regs()
    print("profiles written")
