"""Summarise an `ncu --set full` report into a small JSON for profiles/: per captured launch the duration, DRAM bytes,
tensor-pipe / issue / shared-memory utilisation and the dominant warp-stall reasons.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/rNN_name.json ["note"]"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}


def main(rep, dst, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        e = {"kernel": r[ix["Kernel Name"]].split("(")[0]}
        for k, name in WANT.items():
            if k in ix and r[ix[k]] != "":
                e[name] = {"value": float(r[ix[k]].replace(",", "")), "unit": units[ix[k]]}
        stalls = {}
        for h, i in ix.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") \
                    and "not_issued" not in h and r[i] not in ("", "n/a"):
                v = float(r[i])
                if v >= 0.3:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
        e["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        out.append(e)
    json.dump({"source": rep, "note": note, "launches": out}, open(dst, "w"), indent=1)
    print(f"{len(out)} launches -> {dst}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
