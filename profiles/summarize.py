"""Turns an ncu report (+ launch list) brought back in gpurun_out/ into the tracked summaries.

    python profiles/summarize.py gpurun_out/r1_prof.ncu-rep gpurun_out/r1_launches_bench.csv r1

Writes profiles/<tag>_kernels.json (per captured launch: duration, DRAM bytes, throughput,
occupancy, issue utilisation, ...), profiles/<tag>_launches.csv (the launch list as captured)
and profiles/<tag>_summary.md.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
              "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


OURS = re.compile(r"(encode_|mse_|decode_|oks_|rescore_|pack_|heatmap_acc|scale_inplace|train_geometry|transform_joints|box_affine|center_scale_affine|step_|eval_rows|person_rows)\w*kernel")


def short(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    if rep.endswith(".csv"):        # `ncu -i X.ncu-rep --page raw --csv` already run on the GPU box (reports > 64 MB do not travel)
        with open(rep) as fh:
            raw = fh.read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for m in KEEP:
            if m in idx:
                try:
                    v = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[m]]
                if m.startswith("dram__bytes") or m.startswith("gpu__time"):
                    v *= UNIT_SCALE.get(u, 1.0)
                k[m] = v
        if "gpu__time_duration.sum" in k:
            t = k["gpu__time_duration.sum"]
            k["dram_traffic_bytes"] = k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)
            k["dram_GBps"] = k["dram_traffic_bytes"] / t / 1e9
        kernels.append(k)
    with open(os.path.join(HERE, tag + "_kernels.json"), "w") as fh:
        json.dump(kernels, fh, indent=1)

    agg = collections.OrderedDict()
    with open(launches) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    with open(os.path.join(HERE, tag + "_launches.csv"), "w") as fh:
        fh.writelines(lines)
    skipped = 0
    for r in csv.reader(lines):
        if not r or not r[0].isdigit():
            continue
        if not OURS.search(r[4]):
            skipped += 1            # torch kernels that build the synthetic inputs before the timed steps
            continue
        agg.setdefault(short(r[4]), []).append(float(r[-1]))
    # one step of bench.py launches each of its three kernels once per batch; the same command also
    # runs the fused-training-step loop (encode_mse_* + decode), so shares are taken from the
    # per-launch averages of the three step kernels, not from totals
    step = [k for k in agg if k.startswith(("encode_refine", "mse_ring", "mse_fwd_bwd", "decode_tma", "decode_generic"))]
    if any(k.startswith("step_kernel") for k in agg):       # round 2: the headline step is ONE kernel
        step = [k for k in agg if k.startswith("step_kernel")]
    step_sum = sum(sum(agg[k]) / len(agg[k]) for k in step)
    md = ["# ncu summary %s" % tag, "",
          "Launch list (`%s_launches.csv`, `ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and"
          " serialised: compare shares, not absolutes; %d launches of torch/NCCL setup kernels that generate the synthetic"
          " inputs before the steps are left out of the shares):" % (tag, skipped), "",
          "| kernel | launches | avg us | share of step |", "|---|---|---|---|"]
    for k, v in agg.items():
        share = ("%.3f" % (sum(v) / len(v) / step_sum)) if k in step else "- (fused step)"
        md.append("| `%s` | %d | %.1f | %s |" % (k, len(v), sum(v) / len(v) / 1e3, share))
    md += ["", "Full capture (`ncu --set full --clock-control none --import-source on`, `profiles/prof_driver.py`;"
           " last captured launch per kernel/grid; `%s_kernels.json` has every launch):" % tag, "",
           "| kernel | grid x block | us | DRAM read MB | DRAM write MB | DRAM GB/s | dram % of ncu peak | issue % | warps active % | regs | fp64 pipe % |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    # prof_driver.py launches each kernel PROF_REPS (=2) times at 64x48 (1024 persons/launch), then
    # PROF_REPS times at 96x72 (512 persons/launch); label by order of appearance per kernel
    reps = int(os.environ.get("PROF_REPS", "2"))
    order = collections.Counter()
    last = collections.OrderedDict()
    traffic = {}
    for k in kernels:
        i = order[k["kernel"]]
        order[k["kernel"]] += 1
        shape = "64x48,P=1024" if i < reps else "96x72,P=512"
        if k["kernel"].startswith("encode_mse_tile_kernel<"):       # the template argument names the map width
            shape = "64x48,P=1024" if k["kernel"].startswith("encode_mse_tile_kernel<12") else "96x72,P=512"
        if k["kernel"].startswith("step_kernel") and i >= 2 * reps:
            shape = "64x48,P=128"
        if k["kernel"].split("<")[0] in ("rescore_kernel", "oks_nms_kernel"):
            shape = "512 images, ~10.7k persons"
        if k["kernel"].split("<")[0] in ("eval_rows_nms_kernel", "box_affine_kernel") or (k["kernel"].startswith("decode_tma_kernel<0") and i >= 2 * reps):
            shape = "cfg5 / 8: 12.9k persons, 619 images"
        if k["kernel"].split("<")[0] in ("train_geometry_kernel",):
            shape = "8192 persons"
        k["config"] = shape
        last[(k["kernel"], shape, 0, 0)] = k
        traffic["%s @ %s" % (k["kernel"], shape)] = {"dram_bytes_per_launch": k["dram_traffic_bytes"],
                                                     "us": k["gpu__time_duration.sum"] * 1e6}
    # the library these numbers were measured on: bench.py ignores the file once the sources change
    sys.path.insert(0, os.path.dirname(HERE))
    from simple_pose_b200 import build as _build
    traffic["_lib_fingerprint"] = _build._fingerprint()
    with open(os.path.join(HERE, tag + "_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
    with open(os.path.join(HERE, tag + "_kernels.json"), "w") as fh:
        json.dump(kernels, fh, indent=1)
    for (name, grid, block, smem), k in last.items():
        name = "%s` `%s" % (name, k["config"])
        md.append("| `%s` | %s x %s | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d | %.1f |" % (
            name, k["grid"], k["block"], k["gpu__time_duration.sum"] * 1e6, k.get("dram__bytes_read.sum", 0) / 1e6,
            k.get("dram__bytes_write.sum", 0) / 1e6, k.get("dram_GBps", 0),
            k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
            k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
            k.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0), int(k.get("launch__registers_per_thread", 0)),
            k.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0)))
    with open(os.path.join(HERE, tag + "_summary.md"), "w") as fh:
        fh.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
