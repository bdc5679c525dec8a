"""Summarise a `nvcc -Xptxas -v` log: registers / spills per kernel.  usage: ptxas_summary.py LOG [filter...]"""
import re, subprocess, sys
log = open(sys.argv[1]).read()
filt = sys.argv[2:]
pat = re.compile(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
rows = pat.findall(log)
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(len(rows), "kernels; max regs", max(int(r[4]) for r in rows), "; spilling:", sum(1 for r in rows if int(r[2]) > 0))
for nm, r in zip(names, rows):
    short = nm.replace("cagp::", "").split("(")[0]
    if int(r[2]) > 0 or (filt and all(f in short for f in filt)):
        print(f"{short:90s} regs={r[4]} stack={r[1]} spill={r[2]}/{r[3]}")
