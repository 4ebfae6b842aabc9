"""Per-kernel SASS statistics of the shipped library: instruction count / code bytes and an opcode histogram of the
instructions that prove the memory path (UBLKCP = 1-D TMA bulk copy, LDGSTS = cp.async, SYNCS = mbarrier, DFMA...)."""
import re, subprocess, sys, collections, json
so = sys.argv[1] if len(sys.argv) > 1 else "leg-kilo_b200/liblegkilo_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None; cnt = collections.OrderedDict(); ops = {}
for l in txt.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        name = m.group(1); cnt[name] = 0; ops[name] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if name and m:
        cnt[name] += 1
        ops[name][m.group(1).split('.')[0]] += 1
def short(k):
    out = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    out = re.sub(r'lk::\(anonymous namespace\)::', '', out)
    return re.sub(r'\(.*', '', out)[:80]
rows = []
KEYS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "SHFL", "LDG", "STG", "LDS", "STS", "BAR", "ATOMG", "RED", "MUFU", "CALL"]
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]):
    if v < 64: continue
    r = dict(kernel=short(k), instructions=v, code_kb=round(v * 16 / 1024, 1))
    for o in KEYS: r[o] = ops[k].get(o, 0)
    rows.append(r)
if "--json" in sys.argv: print(json.dumps(rows, indent=1))
else:
    print("%-70s %7s %7s " % ("kernel", "instr", "KB") + " ".join("%6s" % o for o in KEYS))
    for r in rows: print("%-70s %7d %7.1f " % (r["kernel"][:70], r["instructions"], r["code_kb"]) + " ".join("%6d" % r[o] for o in KEYS))
