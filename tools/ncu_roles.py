#!/usr/bin/env python
"""Warp-stall samples of one kernel of an `ncu --set full --import-source on` report, attributed to source lines and grouped into
the kernel's warp roles by source-line ranges (given as name:first-last,...).  Usage:
  python tools/ncu_roles.py report.ncu-rep file.cu "producer:100-200,mma:201-260,..." [kernel-name-regex] > out.json"""
import collections
import csv
import io
import json
import subprocess
import sys


def main():
    rep, fname, spec = sys.argv[1], sys.argv[2], sys.argv[3]
    kern = sys.argv[4] if len(sys.argv) > 4 else None
    roles = []
    for part in spec.split(","):
        name, rng = part.split(":")
        a, b = rng.split("-")
        roles.append((name, int(a), int(b)))
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kern:
        cmd += ["-k", "regex:" + kern, "-c", "1"]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = [i for i, r in enumerate(rows[:12]) if "# Samples" in r][0]
    hdr = rows[hi]
    stall = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    cur, curfile, seen = None, "?", set()
    per_role = collections.defaultdict(lambda: collections.Counter())
    per_line = collections.Counter()
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            curfile = r[1].split("/")[-1]
            continue
        if not r or r[0] == "Line No":
            continue
        if r[0] != "":
            cur = (curfile, int(r[0]) if r[0].isdigit() else -1, r[1][:90])
            continue
        try:
            n, addr = int(r[4]), r[2]
        except (ValueError, IndexError):
            continue
        if addr in seen or cur is None:
            continue
        seen.add(addr)
        role = "other (inlined helpers: %s)" % cur[0] if cur[0] != fname else "unassigned"
        if cur[0] == fname:
            for name, a, b in roles:
                if a <= cur[1] <= b:
                    role = name
        c = per_role[role]
        c["samples"] += n
        for h, i in stall:
            v = int(r[i] or 0)
            if v:
                c[h] += v
        per_line[(cur[0], cur[1], cur[2])] += n
    total = sum(c["samples"] for c in per_role.values())
    out = {"report": rep, "kernel": kernel[:160], "total_samples": total, "roles": {}, "top_lines": []}
    for role, c in sorted(per_role.items(), key=lambda kv: -kv[1]["samples"]):
        s = c.pop("samples")
        out["roles"][role] = {"samples": s, "share": round(s / max(1, total), 4),
                              "stalls": {k: round(v / max(1, s), 3) for k, v in c.most_common(6)}}
    for (f, l, src), n in per_line.most_common(14):
        out["top_lines"].append({"file": f, "line": l, "source": src, "samples": n, "share": round(n / max(1, total), 4)})
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
