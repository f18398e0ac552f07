"""Summarise an .ncu-rep: key metrics + instruction / stall-sample share by function.
usage: ncu_summary.py REPORT [top_lines [CELLSTEPS "title / command line"]]
With CELLSTEPS (cell-steps the profiled launch processed) the header carries the DRAM traffic and the executed FP64 operations per cell-step."""
import csv, sys, re, subprocess, collections, bisect, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cellsteps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
title = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warp_latency_per_inst_issued.ratio']
if cellsteps:
    g = lambda name: float(vals[hdr.index(name)]) if name in hdr and vals[hdr.index(name)] not in ("", "n/a") else 0.0
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = g("dram__bytes_read.sum") * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
    wr = g("dram__bytes_write.sum") * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
    for line in title.split("|"):
        if line.strip():
            print("# " + line.strip())
    print("# DRAM traffic per cell-step: (%.2f + %.2f) GB / %.0f = %.0f B" % (rd / 1e9, wr / 1e9, cellsteps, (rd + wr) / cellsteps))
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    sh = next((r for r in srows if "Thread Instructions Executed" in r), None)
    if sh:
        it, isrc = sh.index("Predicated-On Thread Instructions Executed"), sh.index("Source")
        ops = collections.Counter()
        for r in srows:
            if len(r) == len(sh) and r is not sh:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
                if m and m.group(2) in ("DFMA", "DMUL", "DADD"):
                    try: ops[m.group(2)] += int(r[it] or 0)
                    except ValueError: pass
        print("# executed FP64 per cell-step: DFMA %.0f, DMUL %.0f, DADD %.0f = %.0f flops" % (
            ops["DFMA"] / cellsteps, ops["DMUL"] / cellsteps, ops["DADD"] / cellsteps, (2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / cellsteps))
print("metric,unit,value")
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(vals[i] or 0) > 0.05):
        print("%s,%s,%s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections = []; cur = None
for r in rows:
    if r and r[0] == 'File Path': cur = {'file': r[1], 'rows': [], 'hdr': None}; sections.append(cur)
    elif r and r[0] == 'Function Name': continue
    elif r and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']): cur['rows'].append(r)
def fnranges(path):
    out = []
    try:
        for i, l in enumerate(open(path), 1):
            m = re.match(r'^(?:__device__|__global__|static|extern|MEK_HD|MEK_HDN).*?\b(\w+)\s*\(', l)
            if m: out.append((i, m.group(1)))
    except OSError: pass
    return out
tot = collections.Counter(); samp = collections.Counter(); wait = collections.Counter(); lsb = collections.Counter(); lines = []
for s in sections:
    f = s['file']
    if not f.endswith(('.cuh', '.cu')): continue
    fr = fnranges(f); starts = [a for a, _ in fr]
    h = s['hdr']; iL = h.index('Line No'); iI = h.index('Instructions Executed'); iS = h.index('# Samples'); iw = h.index('stall_wait'); il = h.index('stall_long_sb'); isrc = h.index('Source')
    for r in s['rows']:
        try: ln = int(r[iL]); ins = int(r[iI] or 0); sm = int(r[iS] or 0)
        except ValueError: continue
        k = bisect.bisect_right(starts, ln) - 1
        fn = f.split('/')[-1] + ':' + (fr[k][1] if k >= 0 else '?')
        tot[fn] += ins; samp[fn] += sm; wait[fn] += int(r[iw] or 0); lsb[fn] += int(r[il] or 0)
        lines.append((sm, f.split('/')[-1], ln, int(r[il] or 0), int(r[iw] or 0), ins, r[isrc].strip()[:100]))
T = sum(tot.values()) or 1; S = sum(samp.values()) or 1
print("# by function: share of executed instructions / of stall samples (wait, long scoreboard)")
for fn, v in tot.most_common(16):
    print("# %-40s inst %5.1f%%  samples %5.1f%%  wait %5.1f%%  long_sb %5.1f%%" % (fn, 100 * v / T, 100 * samp[fn] / S, 100 * wait[fn] / S, 100 * lsb[fn] / S))
if top:
    lines.sort(reverse=True)
    for sm, f, ln, l, w, ins, srcl in lines[:top]:
        print("# %5.2f%% lsb %5.2f%% wait %5.2f%% inst %10d %s:%d %s" % (100 * sm / S, 100 * l / S, 100 * w / S, ins, f, ln, srcl))
