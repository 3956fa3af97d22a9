#!/usr/bin/env python
"""Summarise ONE steady-state training step out of an ncu launch list
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_step.py
(STEPS steps of the bench workload incl. fused Adam): the launches from the LAST step's first pack_kernel (parameters ->
packed operands, the first launch of a step) to the end of the list, grouped by kernel.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/r2_launches_steady_state.txt"""
import csv
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], v))
    packs = [i for i, r in enumerate(rows) if "pack_kernel" in r[1]]
    # a step = pack (forward) ... unpack (backward) ... Adam; the last step starts at the second-to-last pack_kernel
    start = packs[-2]
    step = rows[start:]
    tot = sum(r[2] for r in step)
    groups = {}
    for _, name, us in step:
        short = name.split("(")[0].replace("void ", "").replace("wnb::", "")[:84]
        g = groups.setdefault(short, [0, 0.0])
        g[0] += 1
        g[1] += us
    print("# One steady-state training step (the last of the run) of\n#   ncu --metrics gpu__time_duration.sum --clock-control none "
          "python tools/prof_step.py\n# (%s, launches %d..%d).  ncu serialises the launches and runs them cold, so the SHARES "
          "are comparable with the\n# bench line, not the absolute times.\n" % (path, step[0][0], step[-1][0]))
    ours = 0.0
    for name, (n, us) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        lib = "torch" if name.startswith("at::") or "at::" in name else "libwnb200"
        if lib == "libwnb200":
            ours += us
        print("%-86s %3d launches %9.1f us  %5.1f %%  %s" % (name, n, us, 100 * us / tot, lib))
    print("\nsum of kernel time %.1f us in %d launches: libwnb200 %.1f us, torch %.1f us (scalar helpers: loss scale, the optimizer step counter)"
          % (tot, len(step), ours, tot - ours))


if __name__ == "__main__":
    main(sys.argv[1])
