#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS) export by CUDA source line, using nvdisasm line info.

    python tools/ncu_lines.py <src.csv from ncu> <object or cubin> <kernel substring> [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_map(obj, kernel_sub):
    tmp = tempfile.mkdtemp()
    if obj.endswith(".cubin"):
        cubins = [obj]
    else:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    out = subprocess.run(["nvdisasm", "-g", "-c", cubins[0]], capture_output=True, text=True).stdout
    amap, cur_line, cur_file, in_k, inline = {}, None, None, False, ""
    for ln in out.splitlines():
        m = re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            in_k = kernel_sub in m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_file, cur_line, inline = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and in_k:
            amap[int(m.group(1), 16)] = (cur_file, cur_line, inline.strip())
    return amap


def main():
    src_csv, obj, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(obj, ksub)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0, 0])
    base = None
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[idx["Address"]], 16)
            ie = int(r[idx["Instructions Executed"]] or 0)
            ns = int(r[idx["# Samples"]] or 0)
            te = int(r[idx["Thread Instructions Executed"]] or 0)
        except ValueError:
            continue
        if base is None:
            base = addr
        key = amap.get(addr - base, ("?", 0, ""))
        a = agg[(key[0], key[1])]
        a[0] += ie; a[1] += ns; a[2] += te
    ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    srcs = {}
    print("total warp-instructions %d, samples %d" % (ti, ts))
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        path = os.path.join("blueberry_b200", "csrc", f)
        if os.path.exists(path):
            if path not in srcs:
                srcs[path] = open(path).read().splitlines()
            if 0 < l <= len(srcs[path]):
                text = srcs[path][l - 1].strip()[:90]
        print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%-4d %s" % (100.0 * v[0] / ti, 100.0 * v[1] / max(ts, 1), v[2] / max(v[0], 1), f, l, text))


if __name__ == "__main__":
    main()
