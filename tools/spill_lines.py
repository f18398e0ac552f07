"""Source lines of an .ncu-rep's kernel that execute local-memory loads / stores (register spills, arrays in local memory), per
cell-step.  usage: spill_lines.py REPORT CELLSTEPS [TOP]"""
import subprocess, csv, io, re, collections, sys
rep = sys.argv[1]; cs = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur_file = None; cur_line = None; srcs = {}
by = collections.defaultdict(collections.Counter)
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] != '' and r[0].isdigit():
        cur_line = (cur_file, int(r[0])); srcs[cur_line] = r[1]; continue
    if len(r) > 7 and r[2].startswith('0x'):
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        if m:
            try: n = int(r[7] or 0)
            except ValueError: n = 0
            by[cur_line][m.group(2).split('.')[0]] += n
for op in ("LDL", "STL"):
    print("== %s: %.1f per cell-step" % (op, sum(c[op] for c in by.values()) * 32 / cs))
    for k, c in sorted(by.items(), key=lambda kv: -kv[1][op])[:top]:
        if c[op]: print("  %6.1f %s:%d  %s" % (c[op] * 32 / cs, k[0], k[1], srcs[k][:120]))
