"""Condense an .ncu-rep (ncu --set full) into the handful of counters the design discussion uses.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md      (runs here: ncu -i needs no GPU)"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % path.split("/")[-1])
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(H, U, r)}
        print("## %s  (grid %s, block %s)\n" % (d.get("Kernel Name", ("", "?"))[1], d.get("Grid Size", ("", "?"))[1],
                                                d.get("Block Size", ("", "?"))[1]))
        print("| counter | value | unit |\n|---|---|---|")
        for k, name in KEYS:
            if k in d:
                print("| %s (`%s`) | %s | %s |" % (name, k, d[k][1], d[k][0]))
        print()


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(path, label, batch, out="profiles/ncu_traffic.json"):
    """--traffic: append {label, batch, dram bytes read / written, ncu duration, source} of the FIRST kernel of an
    `ncu --set full` capture to profiles/ncu_traffic.json -- what bench.py reports as roofline.traffic."""
    import json
    import os

    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U, r = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(H, U, r)}

    def nbytes(k):
        return float(d[k][1].replace(",", "")) * UNIT[d[k][0]]

    ent = {"label": label, "batch": int(batch), "kernel": d["Kernel Name"][1], "grid": d["Grid Size"][1],
           "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
           "ncu_duration": "%s %s" % (d["gpu__time_duration.sum"][1], d["gpu__time_duration.sum"][0]),
           "source": "ncu --set full --clock-control none, %s" % os.path.basename(path)}
    cur = json.load(open(out)) if os.path.exists(out) else []
    cur = [e for e in cur if not (e["label"] == label and e["batch"] == int(batch))] + [ent]
    json.dump(cur, open(out, "w"), indent=1)
    print(json.dumps(ent))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":     # python tools/ncu_summary.py --traffic x.ncu-rep "attention N=25088 d=32" 64
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1])
