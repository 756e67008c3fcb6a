"""Reduces `ncu -i X.ncu-rep --page source --csv -k regex:step_kernel` (SASS view, one section per captured launch) to
where the warp-state samples of the one-launch step kernel fall: by instruction class and the top instructions.

    python profiles/summarize_source.py gpurun_out/r2final_source_step.csv r2final      # writes profiles/<tag>_step_source.md
"""
import collections
import csv
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LABELS = ["64x48, 1024 persons", "96x72, 512 persons", "64x48, 128 persons (batch-128 step)"]      # prof_driver.py's order


def klass(op):
    o = op.split(".")[0]
    if o == "SYNCS":
        return "mbarrier ops (SYNCS)"
    if o.startswith("UBLKCP") or o.startswith("UTMA"):
        return "TMA issue"
    if o in ("STG", "LDS", "STS"):
        return o
    if o in ("LDG", "LD", "LDC", "LDCU", "ULDC"):
        return "LDG / constant loads"
    if o in ("DADD", "DMUL", "DFMA", "DSETP", "F2F", "I2F", "F2I", "MUFU", "DMNMX"):
        return "float64 / convert / MUFU"
    if o in ("SHFL", "REDUX", "VOTE", "MATCH", "WARPSYNC"):
        return "warp shuffle / redux"
    if o in ("ATOMG", "ATOMS", "RED", "ATOM"):
        return "atomics"
    if o in ("BAR", "MEMBAR", "FENCE", "ERRBAR", "NANOSLEEP", "BSYNC", "BSSY", "CCTL"):
        return "barriers / fences / reconvergence"
    if o in ("BRA", "EXIT", "RET", "CALL", "BRX", "JMP"):
        return "branches (incl. the mbarrier try_wait loop and __syncthreads exits)"
    if o.startswith("F"):
        return "float32 ALU"
    return "integer / other ALU"


def main():
    path, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[starts[0]]
    stall_cols = [c for c in range(len(hdr)) if hdr[c].startswith("stall_") and "Not Issued" not in hdr[c]]
    sections, seen = [], set()
    for j, s in enumerate(starts):
        e = (starts[j + 1] - 1) if j + 1 < len(starts) else len(rows)
        body = [r for r in rows[s + 1:e] if len(r) >= len(hdr)]
        key = tuple((r[1], r[2]) for r in body[:400])
        if key in seen:             # ncu prints a matched launch once per -k match; keep one copy
            continue
        seen.add(key)
        sections.append(body)
    out = ["# Step kernel, warp-state samples by SASS instruction (`ncu --set full --import-source on`, %s)" % tag, "",
           "`step_kernel<1, 1, 0>` (gradient + targets written, no HeatMapAcc). Samples = `Warp Stall Sampling (All Samples)`.", ""]
    for label, body in zip(LABELS, sections):
        total = sum(int(r[2] or 0) for r in body)
        insts = sum(int(r[5] or 0) for r in body)
        by, by_inst, reasons = collections.Counter(), collections.Counter(), collections.Counter()
        for r in body:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
            c = klass(m.group(2) if m else r[1])
            by[c] += int(r[2] or 0)
            by_inst[c] += int(r[5] or 0)
            for col in stall_cols:
                reasons[hdr[col]] += int(r[col] or 0)
        out += ["## %s: %d samples, %.1f M warp instructions" % (label, total, insts / 1e6), "",
                "Stall reasons: " + ", ".join("%s %.0f %%" % (k, 100.0 * v / max(1, sum(reasons.values())))
                                              for k, v in reasons.most_common(7)), "",
                "| instruction class | % of samples | % of warp instructions |", "|---|---|---|"]
        for k, v in by.most_common():
            if v or by_inst[k]:
                out.append("| %s | %.1f | %.1f |" % (k, 100.0 * v / total, 100.0 * by_inst[k] / insts))
        out += ["", "| samples | instruction | stall reasons |", "|---|---|---|"]
        for r in sorted(body, key=lambda r: -int(r[2] or 0))[:8]:
            why = ", ".join("%s %s" % (hdr[c], r[c]) for c in stall_cols if r[c] not in ("0", ""))
            out.append("| %s | `%s` | %s |" % (r[2], r[1].strip(), why))
        out.append("")
    out += ["## Reading", "",
            "* `SYNCS.PHASECHK.TRANS64.TRYWAIT` + `@!P0 BRA` (stall_long_sb) is the warp waiting for its map's TMA copy: 7.8 % of"
            " the samples at 1024 x 64x48, 4.7 % at 512 x 96x72 -- the only place the kernel waits for HBM reads; the float64 factor"
            " computation between issuing the copy and this wait covers the rest of the copy's latency.",
            "* `FADD R37, R32, -R29` (stall_long_sb, 5.1 % / 4.2 % / 10.6 % of the samples) is the first use of the map's joint"
            " coordinates, three `LDG.E.CONSTANT` issued a few instructions earlier: a dependent global load per map whose latency"
            " is exposed. Loading the NEXT map's joint and inverse affine one map ahead (the map index is known once the claim"
            " returns) would hide it in multi-round launches; in the single-round batch-128 launch it is part of the start-up chain"
            " joint -> factors -> wait for the copy. Not done this round (DESIGN.md section 7).",
            "* `@P1 BRA` (stall_barrier, 3.2 % / 4.1 % / 10.2 %) is `__syncthreads`: warps that ran out of maps waiting in the loss"
            " reduction for the CTA's last warp, i.e. the tail imbalance inside a CTA, plus the two start-up barriers.",
            "* The `DMUL` rows (stall_short_sb) wait for the float64 factors read from shared memory in the loss pass; with 9 warps"
            " per SM (2.25 per scheduler) shared-memory latency is not covered by other warps. More warps cost more than they hide"
            " (`profiles/r2_sweeps/`: 12+ warps lose 1-4 %), because every warp adds a read stream and two write streams.",
            "* Stores are not a stall site (STG 0.6 % of samples for 2.1 % of the instructions), nor are the TMA issue and the work"
            " claim (atomics 0.0 %).", ""]
    text = "\n".join(out)
    with open(os.path.join(HERE, tag + "_step_source.md"), "w") as fh:
        fh.write(text)
    print(text)


if __name__ == "__main__":
    main()
