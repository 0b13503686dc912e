#!/usr/bin/env python
"""Per-kernel memory traffic of the frame chain from an ncu metrics CSV.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum \
        --cache-control none --clock-control none -k regex:"planes_kernel|score_kernel|segment_decide|mask_dilate|iou_kernel" \
        --launch-skip 100 --launch-count 100 --csv --log-file traffic.csv python bench.py --steps 60 ...
    python tools/ncu_traffic.py traffic.csv profiles/r02_traffic.json

Writes, per kernel, the MEAN per launch of DRAM bytes read / written, L2 (lts) bytes and duration, next to the
algorithmic bytes of DESIGN.md section 4 for the 1080p full-frame crop."""
import collections
import csv
import json
import re
import sys

NPX = 1920 * 1080
# algorithmic bytes per pixel: what the stage must read and write at least once
ALGO = {"planes": 3 + 7, "score": 3 + 8, "segment_decide": 0, "mask_dilate": 4 + 1, "iou": 2}
NAMES = [("planes_kernel", "planes"), ("score_kernel", "score"), ("segment_decide", "segment_decide"),
         ("mask_dilate", "mask_dilate"), ("iou_kernel", "iou")]


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        key = next((short for pat, short in NAMES if pat in r[ik]), None)
        if key is None:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        acc[key][r[im]].append(v * scale)
    out = {"source": "ncu --cache-control none --clock-control none, metrics dram__bytes_read/write.sum, lts__t_bytes.sum; mean per launch "
                     "in steady state of bench.py (1080p full-frame crop, blended forests)", "kernels": {}}
    tot_dram = tot_lts = tot_algo = 0
    for _, k in NAMES:
        m = acc.get(k)
        if not m:
            continue
        mean = {name: sum(v) / len(v) for name, v in m.items()}
        dr, dw = mean.get("dram__bytes_read.sum", 0.0), mean.get("dram__bytes_write.sum", 0.0)
        rec = {"launches": len(next(iter(m.values()))), "dram_bytes_read": round(dr), "dram_bytes_write": round(dw),
               "dram_bytes": round(dr + dw), "lts_bytes": round(mean.get("lts__t_bytes.sum", 0.0)),
               "duration_us": round(mean.get("gpu__time_duration.sum", 0.0) / 1e3, 2), "algorithmic_bytes": ALGO[k] * NPX}
        out["kernels"][k] = rec
        tot_dram += rec["dram_bytes"]; tot_lts += rec["lts_bytes"]; tot_algo += rec["algorithmic_bytes"]
    compulsory = (3 + 4 + 1 + 1 + 1) * NPX      # frame in, labels in, mask out, mask + truth read by the IoU
    out["frame"] = {"dram_bytes": tot_dram, "lts_bytes": tot_lts, "algorithmic_bytes_sum_of_kernels": tot_algo,
                    "compulsory_bytes": compulsory, "dram_over_compulsory": round(tot_dram / compulsory, 3),
                    "lts_over_compulsory": round(tot_lts / compulsory, 3)}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
