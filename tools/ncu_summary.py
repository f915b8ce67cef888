#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture for profiles/: key metrics of every captured launch and, for
one kernel, where its instructions and stall samples go by SOURCE LINE (SASS addresses of the report joined with
`nvdisasm -g` line info of the library the report was taken from -- so run it against the same libmfcas.so build).

    python tools/ncu_summary.py gpurun_out/r02_deflate2.ncu-rep [--lines mf_deflate2_kernel] > profiles/r02_deflate2_ncu.md
"""
import argparse, collections, csv, os, re, subprocess, sys, tempfile

csv.field_size_limit(10 ** 9)
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--lines", default=None, help="kernel name (substring) for the per-source-line breakdown")
    ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metaflow_b200", "csrc", "libmfcas.so"))
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    rows = list(csv.reader(ncu(["-i", a.report, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s\n" % os.path.basename(a.report))
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## %s\n\n```" % r[ik].split("(")[0])
        for i, h in enumerate(hdr):
            stall = "stalled" in h and h.endswith("per_issue_active.ratio")
            if h in KEYS or (stall and float(r[i] or 0) >= 0.25):
                print("%-88s %16s %s" % (h, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            print("dram read+write per launch = %.4f GB" % ((rd * scale[ur] + wr * scale[uw]) / 1e9))
        except (ValueError, KeyError):
            pass
        print("```\n")
    if not a.lines:
        return
    # ---- per source line
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", a.lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    addr2line, cur, on = {}, None, False
    for l in dis:
        if l.startswith("//--------------------- .text."):
            on = a.lines in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
    src_rows = list(csv.reader(ncu(["-i", a.report, "--page", "source", "--csv", "-k", "regex:" + a.lines]).splitlines()))
    h = [r for r in src_rows if "Address" in r][0]
    body = src_rows[src_rows.index(h) + 1:]
    ia, ii, isamp, iwf = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
    ibar = h.index("stall_barrier")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    base = None
    tot = tots = 0
    for r in body:
        try:
            ad = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = ad
        key = addr2line.get(ad - base, ("?", 0))
        v = agg[key]
        v[0] += int(r[ii] or 0); v[1] += int(r[isamp] or 0); v[2] += int(r[iwf] or 0); v[3] += int(r[ibar] or 0)
        tot += int(r[ii] or 0); tots += int(r[isamp] or 0)
    srcdir = os.path.dirname(a.lib)
    cache = {}
    print("## %s by source line (%d warp instructions, %d stall samples)\n" % (a.lines, tot, tots))
    print("| % instr | % samples | smem wavefronts | barrier samples | line | source |\n|---|---|---|---|---|---|")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
        f, ln = key
        if f not in cache:
            p = os.path.join(srcdir, f)
            cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = cache[f][ln - 1].strip()[:100].replace("|", "\\|") if 0 < ln <= len(cache[f]) else ""
        print("| %.1f | %.1f | %.1f M | %d | %s:%d | `%s` |" % (100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(tots, 1), v[2] / 1e6, v[3], f, ln, text))


if __name__ == "__main__":
    main()
