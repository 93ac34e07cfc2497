#!/usr/bin/env python
"""Copies the judged evidence from gpurun_out/ (scratch) into profiles/ (tracked): bench JSON lines, the ncu launch list,
per-kernel summaries extracted from the ncu --set full reports, and ncu_traffic.json (DRAM bytes per launch of the body-layer
kernel, which bench.py reports as roofline.traffic).  Usage: python tools/make_profiles.py r01"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, SRC = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


ROLL_ROLES = ""     # filled in below from the current sources (line ranges of the role branches)
CHAIN_ROLES = ""


def role_spec(path, markers, after=0):
    """name:first-last,... from marker comments in a kernel source: a role starts at its marker line and ends before the next one."""
    lines = open(path).read().splitlines()
    pos = []
    for name, needle in markers:
        hit = [i + 1 for i, l in enumerate(lines) if needle in l and i + 1 > after]
        if hit:
            pos.append((hit[0], name))
    pos.sort()
    out = []
    for k, (ln, name) in enumerate(pos):
        end = pos[k + 1][0] - 1 if k + 1 < len(pos) else len(lines)
        out.append(f"{name}:{ln}-{end}")
    return ",".join(out)


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarize(rep, tag, out_name):
    hdr, units, rows = raw_page(rep)
    seen, out = {}, []
    for r in rows:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        short = name.split("(")[0].split("::")[-1]
        if short in seen:
            continue
        seen[short] = 1
        rec = {"kernel": name[:160]}
        for k in hdr:
            kk = k.split("TriageCompute.")[-1]
            if kk in KEEP or ("warps_issue_stalled" in kk and kk.endswith("per_issue_active.ratio")):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if "stalled" in kk and v < 0.2:
                    continue
                rec[kk + (" [" + units[hdr.index(k)] + "]" if units[hdr.index(k)] else "")] = v
        out.append(rec)
    json.dump(out, open(os.path.join(OUT, f"{tag}_{out_name}"), "w"), indent=1)
    return out


def main():
    global ROLL_ROLES, CHAIN_ROLES
    csrc = os.path.join(ROOT, "pnp-pds_b200", "csrc")
    CHAIN_ROLES = role_spec(os.path.join(csrc, "dncnn_chain.cu"),
                            [("setup", "extern __shared__ uint8_t smem_raw[];"), ("producer", "// ------------------------------------------------------------ producer"),
                             ("mma_issuer", "// ------------------------------------------------------------ MMA issuer"),
                             ("weight_forwarder", "// ------------------------------------------------------------ CTA 1: tell"),
                             ("publisher", "// ------------------------------------------------------------ publisher"),
                             ("epilogue", "// ------------------------------------------------------------ epilogue"), ("teardown", "  tc_fence_before();\n  __syncthreads();\n  cluster_sync_all();")])
    roll = os.path.join(csrc, "dncnn_roll.cu")
    start = [i + 1 for i, l in enumerate(open(roll).read().splitlines()) if "conv_roll_d_kernel(const" in l]
    ROLL_ROLES = role_spec(roll, [("setup", "conv_roll_d_kernel("), ("producer", "// ------------------------------------------------------------ TMA producer"),
                                  ("mma_issuer", "// ------------------------------------------------------------ MMA issuer"),
                                  ("epilogue", "// ------------------------------------------------------------ epilogue"),
                                  ("converters", "// ------------------------------------------------------------ converters")], after=(start[0] - 1 if start else 276))
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(OUT, exist_ok=True)
    for name in ("default", "cfg1", "cfg1_per_layer", "cfg2", "cfg3", "cfg2b", "cfg5"):
        p = os.path.join(SRC, f"BENCH_{name}.json")
        if os.path.exists(p):
            shutil.copy(p, os.path.join(OUT, f"{tag}_bench_{name}.json"))
    if os.path.exists(os.path.join(SRC, "BENCH_ref.json")):
        shutil.copy(os.path.join(SRC, "BENCH_ref.json"), os.path.join(OUT, f"{tag}_bench_reference_arm.json"))
    for n in (2, 4, 8):
        p = os.path.join(SRC, f"SCALE_{n}.json")
        if os.path.exists(p):
            lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if lines:
                open(os.path.join(OUT, f"{tag}_bench_cfg4_{n}gpu.json"), "w").write(lines[-1] + "\n")
    if os.path.exists(os.path.join(SRC, "launches.csv")):
        shutil.copy(os.path.join(SRC, "launches.csv"), os.path.join(OUT, f"{tag}_launches_cfg4_b8.csv"))
    rep = os.path.join(SRC, "prof_conv_layers.ncu-rep")
    if os.path.exists(rep):
        recs = summarize(rep, tag, "ncu_conv_layers.json")
        for r in recs:
            if "conv_roll_d_kernel" in r["kernel"]:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}

                def val(prefix):
                    k = [k for k in r if k.startswith(prefix + " [")][0]
                    return r[k] * scale[k.split("[")[1].rstrip("]")]
                rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
                px = 8 * 1024 * 1024
                json.dump({"kernel": "roll::conv_roll_d_kernel (row-streaming body layer), cfg4 shape, 8 images (8,388,608 px) per launch",
                           "source": f"ncu --set full --clock-control none (profiles/{tag}_ncu_conv_layers.json)",
                           "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "px_per_launch": px,
                           "bytes_per_px": (rd + wr) / px, "gpu_time_ms": val("gpu__time_duration.sum")},
                          open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
                # bench.py reads roofline.traffic from ncu_traffic.json; the default bench line of the same GPU call was printed
                # before this capture existed, so give it the traffic measured in that call
                bp = os.path.join(OUT, f"{tag}_bench_default.json")
                if os.path.exists(bp):
                    line = json.loads(open(bp).read().strip().splitlines()[-1])
                    if line.get("roofline") and "conv_roll_d_kernel" in line["roofline"].get("kernel", ""):
                        line["roofline"]["traffic"] = rd + wr
                        open(bp, "w").write(json.dumps(line) + "\n")
    rep = os.path.join(SRC, "prof_pointwise.ncu-rep")
    if os.path.exists(rep):
        summarize(rep, tag, "ncu_pointwise.json")
    if os.path.exists(os.path.join(SRC, "hbm_rw.json")):
        shutil.copy(os.path.join(SRC, "hbm_rw.json"), os.path.join(OUT, f"{tag}_hbm_read_write.json"))
    for name, outn in (("prof_chain", "ncu_chain.json"), ("prof_blur", "ncu_blur.json"), ("prof_first_im2col", "ncu_first_im2col.json")):
        rep = os.path.join(SRC, name + ".ncu-rep")
        if os.path.exists(rep):
            summarize(rep, tag, outn)
    if os.path.exists(os.path.join(SRC, "launches_cfg1.csv")):
        shutil.copy(os.path.join(SRC, "launches_cfg1.csv"), os.path.join(OUT, f"{tag}_launches_cfg1.csv"))
    for name in ("timeline_256.txt",):
        if os.path.exists(os.path.join(SRC, name)):
            shutil.copy(os.path.join(SRC, name), os.path.join(OUT, f"{tag}_chain_{name}"))
    # warp-stall samples by warp role (source-line ranges of the kernels' role branches)
    tc = os.path.join(csrc, "dncnn_tc.cu")
    first_at = [i + 1 for i, l in enumerate(open(tc).read().splitlines()) if "conv_first2_kernel(const" in l]
    first_roles = role_spec(tc, [("setup", "conv_first2_kernel(const"), ("record_builders", "// ------------------------------------------------------------ record builders"),
                                 ("tma_producer", "// ------------------------------------------------------------ TMA producer: one (136"),
                                 ("mma_issuer", "// ------------------------------------------------------------ MMA issuer: nine"),
                                 ("epilogue", "// ------------------------------------------------------------ epilogue: four groups"),
                                 ("teardown", "}  // namespace first2")], after=(first_at[0] - 1 if first_at else 0))
    roles = [("prof_conv_layers", "dncnn_roll.cu", ROLL_ROLES, "conv_roll_d_kernel", "ncu_roll_roles.json"),
             ("prof_conv_layers", "dncnn_tc.cu", first_roles, "conv_first2_kernel", "ncu_first_roles.json"),
             ("prof_chain", "dncnn_chain.cu", CHAIN_ROLES, "conv_chain_kernel", "ncu_chain_roles.json")]
    for name, src, spec, kern, outn in roles:
        rep = os.path.join(SRC, name + ".ncu-rep")
        if os.path.exists(rep) and spec:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_roles.py"), rep, os.path.join(ROOT, "pnp-pds_b200", "csrc", src), spec, kern],
                               capture_output=True, text=True)
            if r.returncode == 0 and r.stdout.strip():
                open(os.path.join(OUT, f"{tag}_{outn}"), "w").write(r.stdout)
            else:
                print("ncu_roles failed for", name, r.stderr[-400:])
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
