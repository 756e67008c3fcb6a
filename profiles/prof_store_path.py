"""Write side on the TMA or not: the masked-MSE kernel and the fused encode+loss kernel, default
(st.global.v4 from registers) against the opt-in bulk-store variants (quads staged in the drained ring slot, one
cp.async.bulk.global.shared::cta per chunk), same inputs, for one `ncu --set full` capture:

    ncu --set full --clock-control none -k regex:"mse_ring|encode_mse_tile" -f -o /tmp/store python profiles/prof_store_path.py
    ncu -i /tmp/store.ncu-rep --page raw --csv > gpurun_out/r2store_raw.csv
    python profiles/prof_store_path.py --summarize gpurun_out/r2store_raw.csv      # here: writes profiles/r2_store_path.md
"""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = [                                    # (label, kernel, environment) in launch order, per shape
    ("loss, register stores (default)", "loss", {}),
    ("loss, TMA bulk stores", "loss", {"SP_LOSS_BULK_STORE": "1"}),
    ("fused encode+loss, register stores (default)", "fused", {}),
    ("fused encode+loss, TMA bulk stores", "fused", {"SP_TRAIN_BULK_STORE": "1", "SP_TRAIN_TILE_CFG": "2"}),
]
SHAPES = ((64, 48, 1024), (96, 72, 512))
METRICS = [
    ("us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("instructions", "smsp__inst_executed.sum"),
    ("stall long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
    ("stall sleeping", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"),
    ("stall short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
]


def launch():
    import torch
    from simple_pose_b200 import _abi, synth
    from simple_pose_b200.pipeline import HeatmapHotPath
    dev = torch.device("cuda:0")
    for (h, w, batch) in SHAPES:
        hp = HeatmapHotPath(batch, 17, h, w, device=dev)
        joints = synth.joints(batch, height=h, width=w, seed=1, device=dev)
        pred = synth.heatmaps(batch, height=h, width=w, seed=1, device=dev)
        hp.encode(joints)
        torch.cuda.synchronize()
        for label, kernel, env in VARIANTS:
            for k in ("SP_LOSS_BULK_STORE", "SP_TRAIN_BULK_STORE", "SP_TRAIN_TILE_CFG"):
                os.environ.pop(k, None)
            os.environ.update(env)
            _abi.reload_tuning()
            if kernel == "loss":
                hp.loss_fwd_bwd(pred)
            else:
                hp.train_fused(joints, pred)
            torch.cuda.synchronize()
    print("done")


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    # the encode kernel that fills the targets is not captured (regex), so launches arrive in VARIANTS order per shape
    want = len(VARIANTS) * len(SHAPES)
    assert len(data) == want, "expected %d captured launches, got %d" % (want, len(data))
    out = ["# Write side on the TMA or not (`ncu --set full --clock-control none`, `profiles/prof_store_path.py`)", "",
           "One captured launch per row; stall columns are ncu's `smsp__average_warps_issue_stalled_*_per_issue_active`"
           " (warps waiting per issued instruction). ncu serialises launches and runs them cold, so the durations are not"
           " the bench numbers; the warm back-to-back timings of the same variants are in `profiles/r2_sweeps/`.", ""]
    i = 0
    for (h, w, batch) in SHAPES:
        out += ["## %dx%d, %d persons per launch" % (h, w, batch), "",
                "| variant | kernel | " + " | ".join(m[0] for m in METRICS) + " |", "|---|---|" + "---|" * len(METRICS)]
        for label, _, _ in VARIANTS:
            r = data[i]
            i += 1
            name = r[idx["Kernel Name"]]
            name = name[:name.index("(")] if "(" in name else name
            cells = []
            for title, metric in METRICS:
                v = r[idx[metric]].replace(",", "") if metric in idx else ""
                try:
                    f = float(v)
                    if metric.startswith("gpu__time"):
                        unit = rows[1][idx[metric]]
                        f *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
                    cells.append("%.0f" % f if f >= 1000 else "%.2f" % f)
                except ValueError:
                    cells.append(v or "-")
            out.append("| %s | `%s` | %s |" % (label, name.replace("|", "\\|"), " | ".join(cells)))
        out.append("")
    out += ["## Reading", "",
            "* Neither kernel is issue-bound with register stores (loss 24 % issue-active, fused 31-39 %), so there are no issue"
            " slots for the TMA store path to win back. Staging the quads costs one `st.shared.v4` per quad plus the fence, commit"
            " and wait per chunk: the instruction count goes UP (loss +3.6 %, fused +3.6-7.1 %), `short_scoreboard` (shared-memory"
            " latency) stalls rise from 0.62 to 1.14 warps per issue in the loss kernel and 0.83 -> 1.10 / 1.24 -> 1.83 in the fused"
            " one, and DRAM throughput falls (loss 73-74 % -> 66-67 % of ncu's peak).",
            "* Registers: the loss kernel drops from 56 to 48 per thread, which buys nothing at 128 threads per CTA; the fused"
            " kernel stays at its 128-register cap either way (the pressure is the float64 factor arithmetic, not store addresses).",
            "* Under ncu (cold L2, serialised) the bulk-store loss kernel is 12 % slower; warm and back to back the difference is"
            " within +-1 % at these shapes (`profiles/r2_sweeps/r2c_ubench.log`, `r2h_ubench.log`) (fused kernel at 96x72: 2 % slower with"
            " the retuned tile, 27 % slower without it). The register-store variants stay the default; the bulk-store ones remain selectable"
            " (`SP_LOSS_BULK_STORE`, `SP_TRAIN_BULK_STORE`) and parity-tested.", ""]
    text = "\n".join(out)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_store_path.md"), "w") as fh:
        fh.write(text)
    print(text)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarize":
        summarize(sys.argv[2])
    else:
        launch()
