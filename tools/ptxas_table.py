"""Register / spill / stack table of every kernel from the ptxas -v logs the Makefile keeps (build/obj/*.ptxas)."""
import glob, re, subprocess, sys
rows = []
for f in sorted(glob.glob("build/obj/*.ptxas")):
    txt = open(f).read().splitlines()
    name = None
    for i, l in enumerate(txt):
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            name = m.group(1); continue
        m = re.search(r"Used (\d+) registers", l)
        if m and name:
            prev = txt[i - 1]
            sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", prev)
            sm = re.search(r"(\d+) bytes smem", l)
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"lk::\(anonymous namespace\)::", "", dem); dem = re.sub(r"\(.*", "", dem)
            if dem.startswith("void cub::") or "cub::" in dem: name = None; continue
            rows.append((f.split("/")[-1].replace(".ptxas", ".cu"), dem[:70], int(m.group(1)), int(sp.group(1)) if sp else 0, int(sp.group(2)) if sp else 0, int(sp.group(3)) if sp else 0, int(sm.group(1)) if sm else 0))
            name = None
print("%-18s %-70s %5s %6s %7s %7s %7s" % ("file", "kernel", "regs", "stack", "spill_st", "spill_ld", "static_smem"))
for r in rows: print("%-18s %-70s %5d %6d %7d %7d %7d" % r)
