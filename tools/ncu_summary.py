"""Summarise ncu artefacts into markdown for profiles/ (run here, no GPU needed).
    python tools/ncu_summary.py launches gpurun_out/p_launches.csv
    python tools/ncu_summary.py report gpurun_out/p_gemm.ncu-rep [...]
    python tools/ncu_summary.py traffic gemm=gpurun_out/p_gemm.ncu-rep attention=... residual_ln=...  > profiles/ncu_traffic.json"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster size"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM read rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem LSU wavefronts"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem tensor-core wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe %"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("glc::<unnamed>::", "").replace("void ", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        tot[name][0] += 1
        tot[name][1] += v
    allus = sum(v[1] for v in tot.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {100 * us / allus:.1f}% |")
    print(f"\nTotal {allus / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches.")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, zip(units, r)))
        name = d.get("Kernel Name", ("", "?"))[1]
        print(f"\n### `{re.sub(r'[(].*', '', name)}`  ({path.split('/')[-1]})\n\n| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in d:
                u, v = d[k]
                print(f"| {label} (`{k}`) | {v} {u} |")


def traffic(specs):
    """dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches captured in each report -> the JSON
    bench.py reads (profiles/ncu_traffic.json)"""
    import json
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res, detail = {}, {}
    for spec in specs:
        fam, path = spec.split("=", 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        tot, per = 0.0, []
        for r in rows[2:]:
            d = dict(zip(hdr, zip(units, r)))
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                u, v = d[k]
                b += float(v.replace(",", "")) * mult[u]
            name = re.sub(r"[(].*", "", d["Kernel Name"][1]).replace("glc::<unnamed>::", "").replace("void ", "")
            per.append({"kernel": name[:80], "dram_bytes": b, "duration_us_under_ncu": float(d["gpu__time_duration.sum"][1].replace(",", "")) *
                        {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[d["gpu__time_duration.sum"][0]]})
            tot += b
        res[fam] = tot / max(1, len(per))
        detail[fam] = per
    print(json.dumps({"bytes_per_launch": res, "launches": detail,
                      "note": "dram__bytes_read.sum + dram__bytes_write.sum from `ncu --set full --clock-control none` captures of the "
                              "bench.py workload (scripts/gpu_profiles.sh), averaged over the captured launches of each family; "
                              "cold-cache and serialised, so an upper bound on the in-step DRAM traffic"}, indent=1))


if __name__ == "__main__":
    mode, paths = sys.argv[1], sys.argv[2:]
    if mode == "traffic":
        traffic(paths)
    else:
        for p in paths:
            (launches if mode == "launches" else report)(p)
