"""Turn an `ncu --set full` report into the committed summaries: one metric-per-line CSV per kernel and r2_traffic.json."""
import csv, json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]
def export(rep, kernel, out_csv, workload, alg_bytes, traffic):
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out_csv, "w") as f:
        f.write("# ncu --set full --clock-control none, %s; one column per captured launch\n" % workload)
        f.write("metric,unit," + ",".join("launch%d" % i for i in range(len(rows))) + "\n")
        for h in hdr:
            vals = [r[idx[h]] for r in rows]
            if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"): continue
            f.write('"%s","%s",%s\n' % (h, units[idx[h]], ",".join('"%s"' % v for v in vals)))
    def num(r, k): return float(r[idx[k]].replace(",", ""))
    def to_bytes(r, k):
        u = units[idx[k]].lower(); v = num(r, k)
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    d = sum(to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in rows) / len(rows)
    dur = sum(num(r, "gpu__time_duration.sum") for r in rows) / len(rows)
    traffic[kernel] = dict(dram_bytes_per_launch=d, workload=workload, algorithmic_bytes_per_launch=alg_bytes, launches_captured=len(rows),
                           gpu_time_under_ncu=[r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]] for r in rows],
                           source=os.path.basename(out_csv))
traffic = {}
export(os.path.join(ROOT, "gpurun_out/r2_fused.ncu-rep"), "k_scan_fused", os.path.join(ROOT, "profiles/r2_ncu_k_scan_fused.csv"),
       "leg_fusion_b1 (bench.py default: one 28.8k-point scan x 3 iterations per launch, ring of 512 scans, 1.2M-voxel map)", 304 * 28792 * 3, traffic)
export(os.path.join(ROOT, "gpurun_out/r2_stream2.ncu-rep"), "k_residual_stream2", os.path.join(ROOT, "profiles/r2_ncu_k_residual_stream2.csv"),
       "synth100k_b1024 shard (128 scans x 102 399 points per launch, 10M-voxel map)", 304 * 128 * 102399, traffic)
if os.path.exists(os.path.join(ROOT, "gpurun_out/r2_fallback.ncu-rep")):
    export(os.path.join(ROOT, "gpurun_out/r2_fallback.ncu-rep"), "k_residual_fallback", os.path.join(ROOT, "profiles/r2_ncu_k_residual_fallback.csv"),
           "synth100k_b1024 shard: the ~5 % of the launch's points the hot-image pass could not finish (no algorithmic-byte figure of its own: its points are counted in k_residual_stream2's 304 B)", 0, traffic)
json.dump(traffic, open(os.path.join(ROOT, "profiles/r2_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
