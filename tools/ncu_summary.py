"""Summarise an ncu capture (read here, without a GPU) into the small JSON files kept under profiles/.

    python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep   profiles/ncu_attention_latest.json   [kernel-regex]
    python tools/ncu_summary.py shares gpurun_out/launches.csv   profiles/r2_ncu_launch_shares.json

`full`: the metrics the roofline needs of the first matching kernel of a `--set full` report, plus the SHA-256 of the kernel
sources the report was taken from (bench.py only uses `roofline.traffic` when that hash matches the sources it runs).
`shares`: per-kernel share of the device time in a `--metrics gpu__time_duration.sum` launch list."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct|sm__pipe_tensor.*cycles_active.*pct|"
                  r"sm__inst_executed_pipe_xu.*pct|sm__inst_issued\.avg\.pct|sm__pipe_(alu|fma|fmaheavy)_cycles_active.*pct|"
                  r"sm__warps_active\.avg\.pct|launch__(registers_per_thread|grid_size|block_size|shared_mem)|lts__t_sector_hit_rate|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active|smsp__inst_executed\.sum$|sm__cycles_elapsed\.avg\.per_second|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__inst_executed_pipe_(uniform|tc|tmem)")


def full(rep, out, pattern="attn_fwd"):
    from bench import ATTENTION_SOURCES, source_sha256
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    name_col = head.index("Kernel Name")
    row = next(r for r in rows[2:] if re.search(pattern, r[name_col]))
    rec = {"kernel": row[name_col][:160], "source": os.path.basename(rep), "source_sha256": source_sha256(ATTENTION_SOURCES)}
    for h, u, v in zip(head, units, row):
        if KEEP.search(h):
            rec[h] = f"{v} {u}".strip()
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


def shares(csv_path, out):
    lines = [l for l in open(csv_path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot, by = 0.0, {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
        d = by.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v * scale
        tot += v * scale
    rec = {"launches": sum(d[0] for d in by.values()), "total_ms": tot,
           "kernels": {k: {"launches": d[0], "ms": d[1], "share": d[1] / tot} for k, d in sorted(by.items(), key=lambda kv: -kv[1][1])}}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1)[:2000])


if __name__ == "__main__":
    {"full": full, "shares": shares}[sys.argv[1]](*sys.argv[2:])
