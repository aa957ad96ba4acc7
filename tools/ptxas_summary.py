#!/usr/bin/env python
"""Summarise the ptxas -v logs kept next to the objects: kernel -> registers / spills / smem."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerf_texture_b200/lib/obj")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(log).read()
    names = re.findall(r"Compiling entry function '(\S+)'", txt)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    blocks = re.split(r"ptxas info\s+: Compiling entry function", txt)[1:]
    for n, b in zip(dem, blocks):
        if pat and pat not in n:
            continue
        regs = re.search(r"Used (\d+) registers", b)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        smem = re.search(r"(\d+) bytes smem", b)
        short = re.sub(r"\(.*", "", n)
        print("%-28s %-70s regs=%s spill=%s/%s smem=%s" % (os.path.basename(log)[:-10], short[:70], regs.group(1) if regs else "?",
              spill.group(1) if spill else "?", spill.group(2) if spill else "?", smem.group(1) if smem else 0))
