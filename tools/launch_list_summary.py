"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list for ONE step of bench.py
--profile-mode:  python tools/launch_list_summary.py launches.csv <launches_per_step> [step_index] > profiles/x.md"""
import collections
import csv
import sys


def main(path, per_step, step=1):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    H = rows[0]
    ik, iv, ig, ib = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
    data = rows[1:][step * per_step:(step + 1) * per_step]
    tot = sum(float(r[iv].replace(",", "")) for r in data)
    agg = collections.OrderedDict()
    for r in data:
        a = agg.setdefault(r[ik].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    print("# ncu launch list of `python bench.py --profile-mode --steps 1` (step %d, %d launches)\n" % (step, len(data)))
    print("Cold-cache, serialised per-launch durations (`gpu__time_duration.sum`, `--clock-control none`); only the"
          " SHARES are comparable with the CUDA-event times of bench.py.\n")
    print("step total: %.3f ms\n" % (tot / 1e6))
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, n, v / 1e6, 100 * v / tot))
    print("\n## every launch, in order\n\n| # | kernel | grid | block | us |\n|---|---|---|---|---|")
    for i, r in enumerate(data):
        print("| %d | `%s` | %s | %s | %.1f |" % (i, r[ik].split("(")[0], r[ig], r[ib], float(r[iv].replace(",", "")) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 1)
