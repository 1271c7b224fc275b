"""Build profiles/r2_traffic.json (what bench.py's roofline.traffic reads) from ncu CSV extracts.

usage: python tools/ncu_traffic.py profiles/<extract>.csv:<signature>:<rows> [...]

<extract>.csv is a tools/summarize_ncu.py extract of an `ncu --set full` capture; <signature> is the
kernel's name as the library reports it (anml_model_main_kernel, e.g. "fused_hess_kernel<4, 2, 0, 0, 1, 32>")
and must occur in the CSV's kernel column (spaces ignored); <rows> is the row count of the captured launch.
The table maps signature -> DRAM bytes (read + write) per row."""
import csv
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def squash(s):
    return re.sub(r"\s+", "", s).replace("anml::", "")


def main():
    out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in sys.argv[1:]:
        path, sig, rows = spec.rsplit(":", 2)
        rows = float(rows)
        total, seen = 0.0, set()
        for r in csv.DictReader(open(path)):
            # ncu prints template booleans as 0/1 or false/true depending on the version
            name = squash(r["kernel"]).replace("false", "0").replace("true", "1")
            if squash(sig) not in name:
                continue
            if r["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and r["metric"] not in seen:
                total += float(r["value"]) * UNIT[r["unit"]]
                seen.add(r["metric"])
        if len(seen) != 2:
            raise SystemExit(f"{path}: no dram__bytes_read/write for a kernel matching '{sig}'")
        table[sig] = {"dram_bytes_per_row": total / rows, "csv": os.path.relpath(path, ROOT), "rows": rows}
        print(sig, table[sig])
    json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
