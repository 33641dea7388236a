"""Turns the raw captures of the second half of round 2 (value-indexed SELL, GPU aggregation, FGMRES; files gpurun_out/r3_*) into
the committed summaries under profiles/r03_* (run in the build container; ncu is only used to READ the .ncu-rep)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
KEYS = ['L0.spmv', 'L0.jacobi', 'L0.resid+restrict', 'L0.prolong', 'L0.spmv+dot', 'L0.jacobi+dot', 'L1.spmv', 'L1.jacobi',
        'L1.resid+restrict', 'L1.prolong']


def sweep_rows(path, only=None, rename=None, keys=None):
    d = json.load(open(path))
    out = []
    for name in (only or list(d)):
        row = d[name]
        cells = [f"{row[k]['ms']:.4f} / {row[k]['frac']:.2f}" if k in row and (keys is None or k in keys) else "–" for k in KEYS]
        out.append(f"| {(rename or {}).get(name, name)} | {row['solve_ms']:.2f} | {row['vcycle_ms']:.3f} | " + " | ".join(cells) + f" | {row.get('value_indexed')} |")
    return out


def kernel_sweep():
    hdr = ["| config | solve ms | V-cycle ms | " + " | ".join(KEYS) + " | value_indexed per level |", "|---|---|---|" + "---|" * (len(KEYS) + 1)]
    txt = ["# r03 kernel sweep — value-indexed SELL storage, Poisson 256^3, 1 GPU (ms / fraction of the measured 6548 GB/s, **on the fp64-CSR bytes of "
           "SURVEY 8(d)** so that the rows compare; a value-indexed kernel moves 5 (6) bytes per entry instead of 12, so fractions above 1 mean the kernel "
           "is faster than ANY kernel that reads fp64 values could be)", "",
           "`python scripts/kernel_sweep.py 256 out.json <configs>` under gpurun, L2 flushed before every launch, 22 iterations (random rhs). "
           "value_indexed: bit mask per level, 1 A / 2 P / 4 R one index byte, 8 / 16 / 32 two index bytes.", "",
           "## The variants, in the order they were built", ""] + hdr
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi.json"), ("auto-novi", "auto", "auto-vi1"),
                      {"auto-novi": "fp64 values (round-2 kernels, PAMG_VALUE_INDEX=0)", "auto": "variant 0: two rows per lane, U = 4, 3 CTAs/SM (P sorted)",
                       "auto-vi1": "variant 1: two rows per lane, U = 8, 2 CTAs/SM"})
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi2.json"), ("auto", "auto-vi2"),
                      {"auto": "variant 0 again, P unsorted (padding costs 5 B: 1.41x fill beats the permutation)", "auto-vi2": "variant 2: software-pipelined (next slice's columns in flight)"})
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi3.json"), ("auto-vi3",), {"auto-vi3": "variant 3: FOUR interleaved rows per lane, loads pinned by a data dependence"})
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi4.json"), ("auto-vi8only", "auto"),
                      {"auto-vi8only": "variant 3, one-byte indices only (PAMG_VALUE_INDEX=1)", "auto": "**shipped default**: variant 3 + two-byte indices for A1 (285..2000 distinct values)"})
    txt += ["", "## Look-ahead and level-1 options (same code, later boxes)", ""] + hdr
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi5.json"), ("auto", "auto-ahead"),
                      {"auto": "variant 3 without look-ahead", "auto-ahead": "**shipped**: + extents one iteration early, L2 prefetch of the next slice (PAMG_VI_AHEAD=1)"})
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi6.json"), ("auto", "auto-plong", "auto-w512", "auto-renum", "auto-plong-w512"),
                      {"auto": "shipped defaults (another box)", "auto-plong": "level-1 A as one resident wave + look-ahead (PAMG_VI_PERSIST_LONG=1)",
                       "auto-w512": "coarse rows renumbered by length, window 512 (PAMG_RENUMBER=1)", "auto-renum": "renumbered, window 4096",
                       "auto-plong-w512": "both"})
    txt += sweep_rows(os.path.join(G, "r3_sweep_vi7.json"), ("auto", "auto-occ1", "auto-occ2"),
                      {"auto": "3 CTAs/SM (U = 4, 80 registers; level 0 only in this sweep)", "auto-occ1": "U = 4 squeezed into 64 registers, 4 CTAs/SM (spills in the loop)",
                       "auto-occ2": "**shipped**: U = 2, 64 registers, 4 CTAs/SM, no spill (PAMG_VI_OCC=2)"}, keys=KEYS[:6])
    txt += ["", "None of the level-1 options pays: with two-byte indices the padding of A1 costs 6 bytes per entry, and the renumbering that removes it (fill 1.159 -> 1.006) "
            "scatters the x gathers of A1 and the columns of P0 / R0 (L1 SpMV 0.146 -> 0.179 ms).", ""]
    txt += ["", "## Reading", "",
            "* The 7-point Poisson matrix has 2 distinct values, its smoothed-aggregation P and R 9, the level-1 Galerkin matrix a few hundred to two thousand; "
            "the Q1 elasticity matrix of config 4 has 22, the jump-coefficient matrix of config 5 20 (P: 51, coarse levels 1100-1500). "
            "Storing a byte (or two) per entry that indexes a dictionary of the ORIGINAL doubles keeps every product and every sum bit-identical.",
            "* Halving the bytes alone bought 17 % (0.267 -> 0.221 ms): ncu (profiles/r03_ncu_vi.md) showed the two-row kernel latency bound -- DRAM 48 %, L1 49 %, "
            "0.37 eligible warps per scheduler, 14.7 long-scoreboard stalls per issue. More loads per warp at lower occupancy (variant 1) and a software "
            "pipeline (variant 2) did not help; four rows per lane, 32 apart, did (0.159 ms): twice the rows go through the same number of dependent phases and "
            "the lanes of a warp touch 32 consecutive rows per access (half the L1 wavefronts per row).",
            "* That variant only works when ptxas issues every column / index load of a step before the first x gather. It interleaved them in 4 of 6 modes "
            "(column load, its gathers, next column load, ...); `asm volatile(\"\" ::: \"memory\")` and `__syncwarp()` do not hold non-coherent loads back. "
            "What does: the gather index is offset by the sign of the OR of all the step's columns and indices (zero, but not provably): first run of variant 3 "
            "Jacobi 0.295 ms, with the dependence 0.2225 ms.",
            "* With the loads pinned, U = 2 fits 64 registers without a spill: 4 CTAs/SM, SpMV 0.163 -> 0.144 ms, solve 33.7 -> 31.4 ms (last table).",
            "* Whole solve 43.1 -> 31.4 ms (random rhs, 22 iterations); with the benchmark's rhs (19 iterations) 37.6 -> 27.0 ms.", ""]
    open(os.path.join(P, "r03_kernel_sweep.md"), "w").write("\n".join(txt))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum"]


def ncu_table(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        out.append(f"### `{name}` grid {r[hdr.index('Grid Size')]}")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {r[i]} | {units[i]} |")
        out.append("")
    return out


def ncu():
    txt = ["# r03 ncu --set full: the value-indexed SELL kernels on Poisson 256^3 (one GPU, `scripts/profile_spmv.py 256 1`, L2 flushed before each launch)", "",
           "MODE template argument: 0 y = A x, 1 residual, 2 Jacobi sweep, 3 prolongation + correction (P0), 4 restriction (R0). "
           "Times under ncu are cold-clock and serialised; the shipped numbers are CUDA-event times (profiles/r03_kernel_sweep.md).", "",
           "## Variant 0 (two rows per lane): the capture that showed the latency bound", ""]
    txt += ncu_table(os.path.join(G, "r3_vi_full.ncu-rep"))
    txt += ["## Variant 3 (four interleaved rows per lane, shipped)", ""]
    txt += ncu_table(os.path.join(G, "r3_vi4_full.ncu-rep"))
    txt += ["## Reading", "",
            "* y = A x: 211.6 -> 150.7 us; DRAM traffic 0.827 -> 0.824 GB (format bytes 5 nnz + 16 n = 0.854 GB: 0.97x, nothing re-read), DRAM throughput 48 -> 67 % "
            "of ncu's peak, long-scoreboard stalls per issue 14.7 -> 9.7, same 3 CTAs/SM (68 -> 80 registers).",
            "* The fp64-valued kernel of round 2 moved 1.64 GB for the same product in 255 us (profiles/r02_ncu_traffic.md).", ""]
    open(os.path.join(P, "r03_ncu_vi.md"), "w").write("\n".join(txt))


def traffic_json():
    rep = os.path.join(G, "r3_vi4_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    r = rows[2]   # first launch: k_spmv_sell_vi4<0,...> = y = A x on level 0

    def val(name):
        v, u = float(r[hdr.index(name)].replace(",", "")), rows[1][hdr.index(name)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]

    total = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    path = os.path.join(P, "ncu_traffic.json")
    d = json.load(open(path))
    d["poisson3d-256"]["1"] = dict(spmv_A0_dram_bytes_per_launch=total, algorithmic_bytes=5 * 117047296 + 16 * 16777216,
                                   fp64_csr_bytes=1740111876,
                                   source="profiles/r03_ncu_vi.md (ncu --set full of k_spmv_sell_vi4<MUL>, value-indexed storage)")
    for k in ("elasticity3d-96", "diffusion-jump-3d-256"):   # captured with the fp64-valued kernels of round 2: not the kernel that runs now
        d.pop(k, None)
    json.dump(d, open(path, "w"), indent=1)
    print("traffic", total)


def traffic_three():
    """r3_traffic.ncu-rep: level-0/1 launches of the three BASELINE operators in one process (scripts/profile_traffic.py 1)."""
    rep = os.path.join(G, "r3_traffic.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, name):
        v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "%": 1.0}.get(u, 1.0)

    plain = [json.loads(l) for l in open(os.path.join(G, "r3_traffic_plain.log")) if l.startswith("{")]
    sell = [p for p in plain if p["format"] == 3]          # the capture filter keeps the SELL launches only
    launches = rows[2:]
    assert len(sell) == len(launches), (len(sell), len(launches))
    out = ["# r03 ncu --set full: DRAM traffic of the level-0/1 kernels of the three BASELINE operators with value-indexed storage", "",
           "`ncu --set full --clock-control none -k regex:k_spmv_sell -o r3_traffic python scripts/profile_traffic.py 1` (one launch per kind, L2 flushed before it). "
           "format bytes = what the storage that runs has to move (value-indexed: 5 or 6 bytes per entry + the vectors; fp64 values: 12 + row pointers); "
           "fp64-CSR bytes = SURVEY 8(d).", "",
           "| workload | level | op | kernel | µs (ncu) | DRAM read + written (GB) | format bytes (GB) | traffic / format | fp64-CSR bytes (GB) | DRAM throughput % |", "|---|---|---|---|---|---|---|---|---|---|"]
    tj = json.load(open(os.path.join(P, "ncu_traffic.json")))
    for pl, r in zip(sell, launches):
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        vi = "sell_vi4" in name
        ib = 2 if vi and name.split(",")[2].strip() == "2" else 1
        n, nnz = pl["rows"], pl["nnz"]
        extra = {"spmv": 16 * n, "jacobi": 32 * n, "prolong": 16 * n}[pl["kind"]]     # x + y (+ b, w) ; prolong: x read + written (+ 8 n_c, ignored)
        fmt = ((4 + ib) * nnz + extra) if vi else pl["algorithmic_bytes"]
        out.append(f"| {pl['workload']} | {pl['level']} | {pl['kind']} | `{name}` | {val(r, 'gpu__time_duration.sum'):.1f} | {dram / 1e9:.3f} | {fmt / 1e9:.3f} | "
                   f"{dram / fmt:.2f} | {pl['algorithmic_bytes'] / 1e9:.3f} | {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
        if pl["level"] == 0 and pl["kind"] == "spmv":
            tj[pl["workload"]] = {"1": dict(spmv_A0_dram_bytes_per_launch=int(dram), algorithmic_bytes=int(fmt), fp64_csr_bytes=int(pl["algorithmic_bytes"]),
                                          source="profiles/r03_ncu_traffic.md (ncu --set full, value-indexed storage, gpurun call 17)")}
    out += ["", "Reading: on all three operators the DRAM traffic of the fine-level product is within 3 % of the bytes its storage holds (x crosses the pins once; nothing for a "
            "shared-memory staging of x or an L2 window to recover, SURVEY 8 g2) -- and those bytes are 0.49 (Poisson, jump coefficients) and 0.42 (elasticity) of what a "
            "kernel reading fp64 values moves. The prolongator of the elasticity hierarchy (47 k distinct values) keeps fp64 values: 1.10 GB at 83.5 % DRAM throughput.", ""]
    open(os.path.join(P, "r03_ncu_traffic.md"), "w").write("\n".join(out))
    json.dump(tj, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    kernel_sweep()
    ncu()
    traffic_json()
    traffic_three()
