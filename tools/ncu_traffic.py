#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of an `ncu --set full` capture ->
JSON that bench.py's roofline.traffic reads.

    python tools/ncu_traffic.py SESSION_TAG out.json workload=capture.ncu-rep [workload=capture.ncu-rep ...]

Written into gpurun_out/traffic_session.json by the session scripts (same gpurun call as the bench line that quotes it) and
copied to profiles/traffic_r02.json."""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    tag, out = sys.argv[1], sys.argv[2]
    res = {"session": tag}
    for spec in sys.argv[3:]:
        wl, rep = spec.split("=", 1)
        try:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        except Exception as e:  # noqa: BLE001
            print("skip", rep, e, file=sys.stderr)
            continue
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        for r in rows[2:]:
            name = r[ki].split("<")[0].split("(")[0].replace("void ", "").strip()
            b = float(r[ri]) * UNIT.get(units[ri], 1.0) + float(r[wi]) * UNIT.get(units[wi], 1.0)
            acc.setdefault(name, []).append(b)
        res.setdefault(wl, {}).update({k: sum(v) / len(v) for k, v in acc.items()})
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
