"""Text summaries of the round's ncu captures (run here on the .ncu-rep files gpurun brings back in gpurun_out/).

    python profiles/summarize_ncu.py mlp  gpurun_out/r02_fwd_nosave.ncu-rep gpurun_out/r02_fwd_save.ncu-rep gpurun_out/r02_bwd.ncu-rep
    python profiles/summarize_ncu.py ray  gpurun_out/r02_ray_kernels.ncu-rep
    python profiles/summarize_ncu.py list gpurun_out/r02_launches_bench.csv
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(PEAKS.get("hbm_gbs", 6551.0))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]


def mlp(reps):
    flop = {"forward": 4096 * 192 * 2 * 593408, "backward": 4096 * 192 * 2 * 557696}
    print("ncu --set full --clock-control none, fine network, 4096 rays x 192 samples (one launch each; times under ncu are ~5-10 % above un-profiled)")
    for rep in reps:
        hdr, units, rows = raw(rep)
        r = rows[0]
        g = lambda name: (r[hdr.index(name)], units[hdr.index(name)])
        us = to_us(*g("gpu__time_duration.sum"))
        rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
        name = r[hdr.index("Kernel Name")]
        f = flop["backward" if "backward" in name else "forward"]
        print(f"\n== {os.path.basename(rep)}: {name[:70]}")
        print(f"   duration {us:.1f} us   SM clock {float(g('sm__cycles_elapsed.avg.per_second')[0]):.3f} GHz   algorithmic {f / 1e9:.1f} GFLOP -> {f / us / 1e6:.0f} TFLOP/s")
        for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
                  "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
            if m in hdr:
                print(f"   {m:92s} {float(r[hdr.index(m)]):6.1f} %")
        print(f"   dram read {rd / 1e6:.1f} MB  write {wr / 1e6:.1f} MB  -> {(rd + wr) / us / 1e3:.0f} GB/s = {(rd + wr) / us / 1e3 / HBM:.2f} of the measured HBM peak ({HBM:.0f} GB/s)")
        print(f"   instructions executed (warp level) {float(r[hdr.index('smsp__inst_executed.sum')]) / 1e6:.0f} M")
        top = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_top.py"), rep, "12"], capture_output=True, text=True).stdout
        print("   top SASS lines by warp samples (address, samples, executions, instruction, dominant stalls):")
        for line in top.splitlines()[1:15]:
            print("     " + line[:170])


def ray(reps):
    print(f"ncu --set full --clock-control none (caches flushed before every kernel), 32768 rays, 64 + 128 samples; HBM peak {HBM:.0f} GB/s (MEASURED_PEAKS.json)")
    print("algorithmic bytes per SURVEY 8(d); dram bytes as counted by ncu (writes that stay in the 126 MB L2 do not reach DRAM inside the kernel)")
    alg = {("composite_fwd", 64): 32768 * (20 * 64 + 12 + 4 * 64 + 24), ("composite_fwd", 192): 32768 * (20 * 192 + 12 + 24),
           ("composite_bwd", 64): 32768 * (36 * 64 + 24), ("composite_bwd", 192): 32768 * (36 * 192 + 24)}
    for rep in reps:
        hdr, units, rows = raw(rep)
        seen = collections.Counter()
        print(f"{'kernel':34s} {'grid':>6s} {'us':>8s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s (dram)':>12s} {'of peak':>8s} {'L2 hit %':>9s} {'warps active %':>15s}")
        for r in rows:
            g = lambda name: (r[hdr.index(name)], units[hdr.index(name)])
            name = r[hdr.index("Kernel Name")].split("(")[0]
            us = to_us(*g("gpu__time_duration.sum"))
            rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
            seen[name] += 1
            print(f"{name[:34]:34s} {r[hdr.index('launch__grid_size')]:>6s} {us:8.1f} {rd / 1e6:11.2f} {wr / 1e6:11.2f} {(rd + wr) / us / 1e3:12.0f} {(rd + wr) / us / 1e3 / HBM:8.2f} "
                  f"{float(r[hdr.index('lts__t_sector_hit_rate.pct')]):9.1f} {float(r[hdr.index('sm__warps_active.avg.pct_of_peak_sustained_active')]):15.1f}")


def launches(paths):
    for path in paths:
        rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
        hdr = rows[0]
        ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            if r[im] == "gpu__time_duration.sum":
                k = r[ik].split("(")[0].replace("void ", "")
                tot[k] += float(r[iv].replace(",", "")) / 1e3
                cnt[k] += 1
        s = sum(tot.values())
        print(f"{os.path.basename(path)}: {sum(cnt.values())} launches, {s:.0f} us of kernel time (cold-cache, serialised)")
        for k, v in tot.most_common(30):
            print(f"  {k[:80]:80s} {cnt[k]:5d} launches {v:10.1f} us  {100 * v / s:5.1f} %")


if __name__ == "__main__":
    {"mlp": mlp, "ray": ray, "list": launches}[sys.argv[1]](sys.argv[2:])
