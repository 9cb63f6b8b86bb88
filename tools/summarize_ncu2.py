"""Turns an `ncu --set full --import-source on` capture into the committed summaries (round 2 onwards):
  <out_dir>/ncu_full_summary.json   selected metrics per kernel (all sm__inst_executed_pipe_* / sm__pipe_*_cycles_active included)
  <out_dir>/opcodes_<kernel>.json   dynamic SASS opcode histogram (warp-instructions executed per opcode, from the source page)
  <out_dir>/smem_conflicts_<kernel>.txt  the instructions with shared-memory bank conflicts
  profiles/traffic.json             DRAM bytes and warp-instructions per unit at the profiled batch (read by bench.py)
  profiles/pipes.json               issue-slot and pipe utilisation + instruction mix per kernel (read by bench.py)
usage: python tools/summarize_ncu2.py <file.ncu-rep> <out_dir> <batch> [log_n L]"""
import collections, csv, io, json, os, re, subprocess, sys

rep, out_dir, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
LOG_N, L = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (13, 4)
os.makedirs(out_dir, exist_ok=True)


def ncu_page(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def short_name(name):
    variant = "fast" if "dpfhe::fast" in name else "gen"
    if "ntt_kernel" in name:
        targs = [int(t) for t in re.findall(r"(?:\((?:int|bool)\))?(\d+)", re.search(r"ntt_kernel<([^>]*)>", name).group(1))]
        logn, inv = targs[0], targs[3] == 1
        base = "ntt_kernel_inv" if inv else "ntt_kernel_fwd"
        return (base if logn == 13 else "%s_n%d" % (base, 1 << logn)), variant
    if "ntt_inv_tma_kernel" in name:   # the inverse direction's kernel since the TMA-fed load (same passes, same summary key)
        logn = int(re.search(r"ntt_inv_tma_kernel<(?:\(int\))?(\d+)", name).group(1))
        return ("ntt_kernel_inv" if logn == 13 else "ntt_kernel_inv_n%d" % (1 << logn)), variant
    if "ks_fused_kernel" in name:
        return "ks_fused_kernel_mul_relin", variant
    return re.sub(r"[^a-z_]", "", name.split("<")[0].split("::")[-1]), variant


src = ncu_page("source")
blocks = re.split(r'(?m)^"Kernel Name",', src)[1:]
# the raw page prints kernel names without their namespace; the source page keeps it: variant (gen / fast) per kernel from there
VARIANT = {}
for blk in blocks:
    nm = next(csv.reader(io.StringIO('"Kernel Name",' + blk.split("\n", 1)[0])))[1]
    VARIANT[short_name(nm)[0]] = short_name(nm)[1]

rows = list(csv.reader(io.StringIO(ncu_page("raw"))))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [h for h in hdr if h in (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio")
    or re.match(r"sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active$", h)
    or re.match(r"sm__pipe_[a-z0-9_]+_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$", h)
    or re.match(r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio$", h)]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
summary, traffic, pipes = {}, {}, {}
for r in data:
    name = r[idx["Kernel Name"]]
    short, variant = short_name(name)
    variant = VARIANT.get(short, variant)
    d = {"kernel": name[:160], "variant": variant}
    for w in WANT:
        v = num(r[idx[w]])
        if isinstance(v, float) and (v != 0 or "stall" not in w):
            d[w] = {"value": v, "unit": units[idx[w]]}
    summary[short] = d
    rd = d["dram__bytes_read.sum"]["value"] * SCALE.get(d["dram__bytes_read.sum"]["unit"], 1)
    wr = d["dram__bytes_write.sum"]["value"] * SCALE.get(d["dram__bytes_write.sum"]["unit"], 1)
    logn = int(re.search(r"_n(\d+)$", short).group(1)).bit_length() - 1 if re.search(r"_n(\d+)$", short) else LOG_N
    is_ntt = "ntt" in short
    units_per_launch = B * 2 * L if is_ntt else B
    alg = units_per_launch * ((2 << logn) * 8 if is_ntt else 6 * L * (8 << logn))
    traffic[short] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes_per_launch": alg,
                      "ratio": (rd + wr) / alg, "profiled_batch": B, "units_per_launch": units_per_launch,
                      "dram_bytes_per_unit": (rd + wr) / units_per_launch,
                      "warp_instructions_per_unit": d["smsp__inst_executed.sum"]["value"] / units_per_launch, "variant": variant,
                      "source": os.path.basename(rep)}
    g = lambda k: d.get(k, {}).get("value")
    pipes[short] = {"issue_slots_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fmaheavy_cycles_active_pct_of_elapsed": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    "fma_cycles_active_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                    "alu_cycles_active_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                    "lsu_inst_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                    "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "active_over_elapsed_cycles": (g("sm__cycles_active.avg") or 0) / (g("sm__cycles_elapsed.avg") or 1),
                    "warp_instructions_per_unit": traffic[short]["warp_instructions_per_unit"], "variant": variant,
                    "source": os.path.basename(rep)}
    print("%-28s %s dram/alg %.3f  inst/unit %.0f  issue %.1f%%  fmaheavy %.1f%% (elapsed)  alu %.1f%%" % (
        short, variant, traffic[short]["ratio"], traffic[short]["warp_instructions_per_unit"], pipes[short]["issue_slots_active_pct"] or 0,
        pipes[short]["fmaheavy_cycles_active_pct_of_elapsed"] or 0, pipes[short]["alu_cycles_active_pct"] or 0))

# ---- source page: dynamic opcode histogram and bank conflicts per kernel
for blk in blocks:
    lines = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
    name = lines[0][1]
    short, variant = short_name(name)
    h = lines[1]
    i_src, i_exec = h.index("Source"), h.index("Instructions Executed")
    i_conf = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None   # wavefronts beyond the ideal count
    i_wave = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
    hist, conflicts, total = collections.Counter(), [], 0
    for row in lines[2:]:
        if len(row) <= i_exec:
            continue
        txt = row[i_src].strip()
        toks = txt.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        n = num(row[i_exec])
        if not isinstance(n, float):
            continue
        hist[op] += n
        total += n
        if i_conf is not None and isinstance(num(row[i_conf]), float) and num(row[i_conf]) > 0:
            conflicts.append((num(row[i_conf]), num(row[i_wave]) if i_wave is not None else 0, n, txt))
    cls = collections.Counter()
    for op, n in hist.items():
        if op.startswith("IMAD.WIDE"):
            cls["IMAD.WIDE (64-bit product)"] += n
        elif op.startswith("IMAD.HI"):
            cls["IMAD.HI"] += n
        elif op.startswith(("IMAD.MOV", "IMAD.X", "IMAD.IADD", "IMAD.SHL")):
            cls["IMAD used as move / add"] += n
        elif op.startswith("IMAD"):
            cls["IMAD (32-bit)"] += n
        elif op.startswith(("IADD3", "VIADD", "LOP3", "SHF", "SEL", "ISETP", "MOV", "LEA", "PRMT", "IABS", "HFMA2", "CS2R")):
            cls["integer ALU (add / logic / select / move)"] += n
        elif op.startswith(("LDS", "STS")):
            cls["shared memory"] += n
        elif op.startswith(("LDG", "STG", "LD.", "ST.", "LDL", "STL", "LDC", "ATOM", "RED", "CCTL", "MEMBAR", "ERRBAR")):
            cls["global / local / constant memory"] += n
        else:
            cls["control / other"] += n
    units_per_launch = traffic.get(short, {}).get("units_per_launch", B)
    out = {"kernel": name[:160], "variant": variant, "warp_instructions": total, "per_unit": total / units_per_launch,
           "classes_share": {k: round(v / total, 4) for k, v in cls.most_common()},
           "opcodes": {k: v for k, v in hist.most_common(40)}}
    json.dump(out, open(os.path.join(out_dir, "opcodes_%s.json" % short), "w"), indent=1)
    if short in pipes:
        pipes[short]["instruction_mix"] = out["classes_share"]
    with open(os.path.join(out_dir, "smem_conflicts_%s.txt" % short), "w") as f:
        f.write("shared-memory accesses with excess wavefronts, i.e. bank conflicts (excess wavefronts, wavefronts, executions, SASS)\n")
        for c in sorted(conflicts, reverse=True)[:40]:
            f.write("%12.0f %12.0f %12.0f  %s\n" % c)
    print("%-28s instruction mix:" % short, {k: round(v / total, 3) for k, v in cls.most_common()})

json.dump(summary, open(os.path.join(out_dir, "ncu_full_summary.json"), "w"), indent=1)
for path, new in (("profiles/traffic.json", traffic), ("profiles/pipes.json", pipes)):
    try:
        old = json.load(open(path))
    except Exception:
        old = {}
    old.update(new)
    json.dump(old, open(path, "w"), indent=1)
