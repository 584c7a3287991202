"""DRAM traffic of the contraction launches of ONE UNet forward, from an ncu CSV taken with
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file profiles/rNN_dram_unet_p64.csv python tools/profile_unet.py --patches 64 --iters 1
Writes the JSON bench.py reads for roofline.traffic (bytes per contraction launch):
    python tools/traffic_from_ncu.py profiles/rNN_dram_unet_p64.csv > profiles/rNN_traffic.json"""
import csv
import json
import sys


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main(path):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    # launches in order; the last forward starts at the last timestep-embedding linear_kernel triple
    ids = []
    by_id = {}
    for r in rows:
        i = int(r["ID"])
        if i not in by_id:
            by_id[i] = {"name": r["Kernel Name"]}
            ids.append(i)
        by_id[i][r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    lin = [i for i in ids if "linear_kernel" in by_id[i]["name"]]
    start = lin[-3] if len(lin) >= 3 else ids[0]
    sel = [by_id[i] for i in ids if i >= start and "gemm_tc" in by_id[i]["name"]]
    rd = sum(to_bytes(*k["dram__bytes_read.sum"]) for k in sel)
    wr = sum(to_bytes(*k["dram__bytes_write.sum"]) for k in sel)
    out = {"source": f"{path} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one UNet call, P=64, bf16)",
           "contraction_launches": len(sel), "dram_read_bytes": rd, "dram_write_bytes": wr,
           "traffic_bytes_per_launch": (rd + wr) / max(len(sel), 1)}
    if sel and "gpu__time_duration.sum" in sel[0]:
        out["serialised_us"] = sum(float(k["gpu__time_duration.sum"][0].replace(",", "")) for k in sel) / 1e3
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
