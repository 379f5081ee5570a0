"""Turn ncu artefacts under gpurun_out/ into the small, tracked summaries under profiles/.
usage: summarize_profiles.py launches <csv> <out.md>   |   summarize_profiles.py kernel <ncu-rep> <out.json> [name]"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(val.replace(",", "")) * mult


def kernel(rep, out, name=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    res = []
    for val in rows[2:]:
        d = {"kernel": val[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = {"value": val[i], "unit": unit[i]}
        rd = d.get("dram__bytes_read.sum")
        wr = d.get("dram__bytes_write.sum")
        if rd and wr:
            d["dram_bytes_per_launch"] = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        res.append(d)
    # stall breakdown from the source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    stalls = collections.Counter()
    if len(srows) > 2:
        # the source page repeats a (kernel, header) pair per profiled launch; take the first launch
        h = srows[1]
        for r in srows[2:]:
            if len(r) != len(h) or r[0] == "Kernel Name":
                break
            for i, k in enumerate(h):
                if k.startswith("stall_") and "Not Issued" not in k:
                    try:
                        stalls[k] += int(r[i] or 0)
                    except ValueError:
                        pass
    summary = {"source_report": rep, "name": name, "launches": res,
               "warp_stall_samples_first_launch": dict(stalls.most_common())}
    if res and "dram_bytes_per_launch" in res[0]:
        summary["dram_bytes_per_launch"] = res[0]["dram_bytes_per_launch"]
    json.dump(summary, open(out, "w"), indent=1)
    print("wrote", out)


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0] != "ID" and not r[0].startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name = r[4]
        short = name.split("(")[0][-90:]
        ns = float(r[14])
        a = agg.setdefault((short, r[8]), [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(out, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        fh.write("source: `%s`  (%d launches, %.1f us total)\n\n" % (path, len(rows), total / 1e3))
        fh.write("| kernel | launches | total us | avg us | share | block | grid |\n|---|---|---|---|---|---|---|\n")
        for (k, _g), (n, ns, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.1f | %.1f | %.1f%% | %s | %s |\n" % (k, n, ns / 1e3, ns / 1e3 / n, 100 * ns / total, blk, grd))
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
