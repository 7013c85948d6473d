#!/usr/bin/env python
"""Summarises an ncu report (run here, no GPU needed): one row per kernel launch with the figures the roofline uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md profiles/x.md] [--raw-gz profiles/x_raw.csv.gz -k regex]

Columns: duration, DRAM bytes read / written, achieved DRAM GB/s (bytes / duration), DRAM % of ncu's peak, tensor-pipe % of
peak (sm__pipe_tensor_cycles_active ... pct_of_peak_sustained_active), warps active %, registers, grid."""
import argparse
import csv
import gzip
import io
import re
import subprocess
import sys

MUL = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TMUL = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--md", default="")
    ap.add_argument("--raw-gz", default="", help="also store the raw page (filtered by -k) gzip-compressed")
    ap.add_argument("-k", default="", help="regex on the kernel name for --raw-gz")
    ap.add_argument("--title", default="")
    ap.add_argument("--merge", type=int, default=1, help="1: merge launches with the same (kernel, grid) into one row (mean)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(h)}

    def val(r, name, mul=None):
        i = col.get(name)
        if i is None or i >= len(r) or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        if mul is not None:
            v *= mul.get(units[i], 1.0)
        return v

    out = []
    for r in data:
        name = re.sub(r"^void (<unnamed>::)?", "", r[col["Kernel Name"]])
        name = re.sub(r"\(.*$", "", name)
        dur = val(r, "gpu__time_duration.sum", TMUL)
        rd, wr = val(r, "dram__bytes_read.sum", MUL) or 0.0, val(r, "dram__bytes_write.sum", MUL) or 0.0
        out.append(dict(kernel=name, grid=r[col["Grid Size"]], block=r[col["Block Size"]], us=dur, rd=rd, wr=wr,
                        dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        warps_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                        regs=val(r, "launch__registers_per_thread"), n=1))
    if a.merge:
        merged = {}
        for o in out:
            key = (o["kernel"], o["grid"])
            m = merged.get(key)
            if m is None:
                merged[key] = dict(o)
            else:
                for f in ("us", "rd", "wr", "dram_pct", "tensor_pct", "warps_pct"):
                    if o[f] is not None and m[f] is not None:
                        m[f] = (m[f] * m["n"] + o[f]) / (m["n"] + 1)
                m["n"] += 1
        out = list(merged.values())
    lines = []
    if a.title:
        lines += [f"# {a.title}", ""]
    lines += ["| kernel | grid | block | launches | time us | dram read MB | dram write MB | achieved GB/s | dram % of ncu peak | tensor pipe % | warps active % | regs |",
              "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    f = lambda v, d=1: "-" if v is None else f"{v:.{d}f}"
    for o in out:
        gbs = (o["rd"] + o["wr"]) / (o["us"] * 1e-6) / 1e9 if o["us"] else None
        lines.append(f"| `{o['kernel'][:70]}` | {o['grid']} | {o['block']} | {o['n']} | {f(o['us'])} | {f(o['rd'] / 1e6, 2)} | {f(o['wr'] / 1e6, 2)} | {f(gbs, 0)} | "
                     f"{f(o['dram_pct'])} | {f(o['tensor_pct'])} | {f(o['warps_pct'])} | {f(o['regs'], 0)} |")
    text = "\n".join(lines) + "\n"
    if a.md:
        with open(a.md, "a" if a.title else "w") as fh:
            fh.write(text + "\n")
    sys.stdout.write(text)
    if a.raw_gz:
        keep = [r for r in data if re.search(a.k, r[col["Kernel Name"]])] if a.k else data
        with gzip.open(a.raw_gz, "wt") as fh:
            w = csv.writer(fh)
            w.writerow(h); w.writerow(units); w.writerows(keep)


if __name__ == "__main__":
    main()
