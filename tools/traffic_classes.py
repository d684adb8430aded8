"""Per-instruction memory traffic of a kernel from an ncu report (source page): L1 tag requests and the sectors each load / store /
reduction asks of L2, grouped into the traffic classes of the raster path. usage: traffic_classes.py rep kernel-regex"""
import csv
import io
import subprocess
import sys


def rows_of(rep: str, kernel: str):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]

    def col(name):
        return hdr.index(name) if name in hdr else None
    c = {k: col(k) for k in ("Source", "Instructions Executed", "L1 Tag Requests Global", "L2 Theoretical Sectors Global",
                             "L2 Theoretical Sectors Global Ideal", "L2 Theoretical Sectors Local")}
    out = []
    for n, r in enumerate(rows[2:]):
        if len(r) <= c["Instructions Executed"] or not r[c["Instructions Executed"]].isdigit():
            continue
        s = r[c["Source"]].strip()
        if not any(k in s for k in ("LDG", "STG", "RED", "ATOMG", "LDL", "STL")):
            continue

        def val(k):
            i = c[k]
            return int(r[i]) if i is not None and i < len(r) and r[i].isdigit() else 0
        out.append({"n": n, "exec": val("Instructions Executed"), "tags": val("L1 Tag Requests Global"), "sectors": val("L2 Theoretical Sectors Global"),
                    "ideal": val("L2 Theoretical Sectors Global Ideal"), "local": val("L2 Theoretical Sectors Local"), "sass": s})
    return out


if __name__ == "__main__":
    for o in rows_of(sys.argv[1], sys.argv[2]):
        if o["exec"] > 1000:
            print(f"#{o['n']:4d} exec={o['exec']:9d} tags={o['tags']:9d} sectors={o['sectors']:9d} ideal={o['ideal']:9d} local={o['local']:8d} {o['sass'][:70]}")
