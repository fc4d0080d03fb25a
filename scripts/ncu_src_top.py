"""Top stall lines per source file from an ncu report (development aid).  usage: ncu_src_top.py report.ncu-rep [launch_index_from_end]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
secs = txt.split('"File Path",')[1:]
# group sections per launch: a new launch starts when the first file path repeats
first = secs[0].splitlines()[0]
starts = [i for i, s in enumerate(secs) if s.splitlines()[0] == first]
starts.append(len(secs))
lo, hi = starts[-1 - which], starts[-which]
names = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_membar", "stall_mio",
         "stall_lg", "stall_math", "stall_selected", "stall_not_selected", "stall_sleep", "stall_dispatch", "stall_misc", "stall_drain", "stall_tex"]
for sec in secs[lo:hi]:
    lines = sec.splitlines()
    path = lines[0].strip('"')
    print("=====", path, lines[1][:110])
    rdr = list(csv.reader(io.StringIO("\n".join(lines[2:]))))
    hdr = rdr[0]
    iS = hdr.index("# Samples")
    idx = {n: hdr.index(n) for n in names if n in hdr}
    per, cur, tot = collections.OrderedDict(), None, 0
    for r in rdr[1:]:
        if len(r) < len(hdr):
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:100])
            per.setdefault(cur, [0, collections.Counter()])
        elif cur is not None:
            try:
                s = int(r[iS] or 0)
            except ValueError:
                s = 0
            per[cur][0] += s
            tot += s
            for n, i in idx.items():
                try:
                    per[cur][1][n] += int(r[i] or 0)
                except ValueError:
                    pass
    print("   total samples", tot)
    for (ln, src), (s, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:12]:
        if s == 0:
            break
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  L{ln:>4} {src[:86]:86s} {dict(c.most_common(3))}")
