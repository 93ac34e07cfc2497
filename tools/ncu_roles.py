#!/usr/bin/env python
"""Warp-stall samples of one kernel of an `ncu --set full --import-source on` report, grouped into the kernel's warp roles.

The kernels of this library branch on the warp index into roles (TMA producer, MMA issuer, epilogue, ...); every SASS instruction
belongs to exactly one role's code (inlined helpers are copied per call site).  ncu's combined "cuda,sass" source view does not say
which FILE a source line comes from, so roles are assigned in SASS address order: an instruction whose source line is a line of the
kernel's own .cu file (line number and text both match) takes the role of that line's range; instructions from inlined helpers
(mbarrier waits, the shared epilogue, ...) inherit the role of the nearest preceding own-file instruction.
  python tools/ncu_roles.py report.ncu-rep path/to/kernel.cu "producer:100-200,mma:201-260,..." [kernel-name-regex] > out.json"""
import collections
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, src_path, spec = sys.argv[1], sys.argv[2], sys.argv[3]
    kern = sys.argv[4] if len(sys.argv) > 4 else None
    roles = []
    for part in spec.split(","):
        name, rng = part.split(":")
        a, b = rng.split("-")
        roles.append((name, int(a), int(b)))
    own = open(src_path).read().splitlines()
    norm = lambda t: "".join(t.split())[:40]
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kern:
        cmd += ["-k", "regex:" + kern, "-c", "1"]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = [i for i, r in enumerate(rows[:12]) if "# Samples" in r][0]
    hdr = rows[hi]
    stall = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    insts, cur, seen = [], None, set()
    for r in rows[hi + 1:]:
        if not r or r[0] in ("Line No", "File Name"):
            continue
        if r[0] != "":
            ln = int(r[0]) if r[0].isdigit() else -1
            is_own = 0 < ln <= len(own) and norm(own[ln - 1]) != "" and norm(own[ln - 1]) == norm(r[1])[:len(norm(own[ln - 1]))]
            cur = (ln, r[1][:90], is_own)
            continue
        try:
            n, addr = int(r[4]), int(r[2], 16)
        except (ValueError, IndexError):
            continue
        if addr in seen or cur is None:
            continue
        seen.add(addr)
        insts.append((addr, n, {h: int(r[i] or 0) for h, i in stall if (r[i] or "0") != "0"}, cur))
    insts.sort()
    per_role = collections.defaultdict(collections.Counter)
    per_line = collections.Counter()
    role = "setup"
    for addr, n, st, (ln, text, is_own) in insts:
        if is_own:
            for name, a, b in roles:
                if a <= ln <= b:
                    role = name
        c = per_role[role]
        c["samples"] += n
        for h, v in st.items():
            c[h] += v
        per_line[(role, ln, text, "own" if is_own else "inlined")] += n
    total = sum(c["samples"] for c in per_role.values())
    out = {"report": os.path.basename(rep), "kernel": kernel[:160], "source": os.path.basename(src_path), "role_line_ranges": spec,
           "total_samples": total, "roles": {}, "top_lines": []}
    for name, c in sorted(per_role.items(), key=lambda kv: -kv[1]["samples"]):
        s = c.pop("samples")
        out["roles"][name] = {"samples": s, "share": round(s / max(1, total), 4),
                              "stalls": {k: round(v / max(1, s), 3) for k, v in c.most_common(6)}}
    for (name, ln, text, kind), n in per_line.most_common(16):
        out["top_lines"].append({"role": name, "line": ln, "file": kind, "source": text, "samples": n, "share": round(n / max(1, total), 4)})
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
