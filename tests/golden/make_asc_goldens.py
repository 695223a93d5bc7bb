"""Copies the reference's two adaptive-sparse-contacts application cases — parameter file (input
data) plus, from the golden .output, the logged statistics every 100 iterations and the final
per-cell mobility status that `set type = mobility_status` prints in deal.II intermediate format
(dem.cc:771-783) — into tests/golden/apps/ (run where /root/reference is mounted).

    python tests/golden/make_asc_goldens.py
"""
import json
import os
import re
import shutil

REF = "/root/reference/applications_tests/lethe-particles"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "apps")
CASES = {"mobility_status": "mobility_status.output", "load_balancing_mobility_status": "load_balancing_mobility_status.mpirun=2.output"}
ROWS = ("Contact list generation", "Velocity magnitude", "Angular velocity magnitude", "Translational kinetic energy", "Rotational kinetic energy")


def parse(path):
    lines = open(path).read().splitlines()
    log, it, sync = [], None, False
    for ln in lines:
        m = re.match(r"Transient iteration:\s+(\d+)", ln)
        if m:
            it = int(m.group(1))
            log.append({"iteration": it})
        if "Synchronized particle statistics" in ln:
            sync = True
            log.append({"iteration": "synchronized"})
        for name in ROWS:
            if ln.startswith("| " + name):
                vals = [float(v) for v in ln.split("|")[2:6]]
                log[-1][name] = vals
    # patches: vertex coordinates line (24 numbers), ..., data line of 8 equal values
    cells = []
    k = 0
    while k < len(lines):
        if lines[k].startswith("[deal.II intermediate Patch<3,3>]"):
            coords = [float(v) for v in lines[k + 2].split()]
            data = [float(v) for v in lines[k + 7].split()]
            assert len(coords) == 24 and len(data) == 8 and len(set(data)) == 1, (k, coords, data)
            cells.append({"lo": coords[0:3], "hi": coords[21:24], "status": int(data[0])})
            k += 8
        else:
            k += 1
    return {"log": log, "cells": cells}


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {}
    for case, output in CASES.items():
        shutil.copy(f"{REF}/{case}.prm", f"{OUT}/{case}.prm")
        gold[case] = parse(f"{REF}/{output}")
        print(case, len(gold[case]["log"]), "log blocks", len(gold[case]["cells"]), "cells")
    with open(os.path.join(OUT, "mobility_status.json"), "w") as f:
        json.dump(gold, f, indent=0)


if __name__ == "__main__":
    main()
