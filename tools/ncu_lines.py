#!/usr/bin/env python
"""Top source lines by warp-stall samples for each kernel of an ncu report captured with --import-source on.
   python tools/ncu_lines.py report.ncu-rep [name-substring] [N]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern = None; hdr = None; cur = None; agg = None; stall_cols = None
def flush():
    if kern is None or agg is None or sub not in kern: return
    tot = sum(v[0] for v in agg.values())
    print("====", kern[:140], "samples", tot)
    for (ln, text), (n, st, ninst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        top = ", ".join(f"{k}={v}" for k, v in st.most_common(3))
        print(f"{n:6d} {100.0*n/max(tot,1):5.1f}%  inst={ninst:9d}  L{ln:<5} {text[:70]:70s} | {top}")
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        flush(); kern = r[1]; agg = collections.defaultdict(lambda: [0, collections.Counter(), 0]); continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
        stall_cols = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and "(Not" not in h]
        continue
    if r[0] in ("File Path",): continue
    if hdr is None or agg is None: continue
    if r[0] != "":
        cur = (r[0], r[1].strip()); continue
    if cur is None: continue
    try: n = int(r[si]); ni = int(r[ii])
    except (ValueError, IndexError): continue
    a = agg[cur]; a[0] += n; a[2] += ni
    for name, i in stall_cols:
        try:
            v = int(r[i] or 0)
        except ValueError: v = 0
        if v: a[1][name] += v
flush()
