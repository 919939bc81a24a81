"""Turn the ncu outputs of tools/profile_r2.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_list():
    rows = [l for l in open(os.path.join(G, "r2_launches.csv")) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = collections.OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit.replace("second", "s").replace("nsecond", "ns"), 1e-6)
        k = r["Kernel Name"].split("(")[0][:60]
        c, t = per.get(k, (0, 0.0))
        per[k] = (c + 1, t + ms)
    tot = sum(t for _, t in per.values())
    n = sum(c for c, _ in per.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 2500  python bench.py --steps 1 --warmup 1 --no-cpu   (c2, 1 x B200)",
           "per-launch times are cold-cache and serialised: compare SHARES.  First 2500 launches = warm-up cycle + part of the timed cycle",
           "total device time %.1f ms over %d launches" % (tot, n)]
    for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:15]:
        out.append("%-62s %6d launches %10.3f ms %5.1f %%" % (k, c, t, 100 * t / tot))
    open(os.path.join(P, "r2_launch_list_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:8]))


def full_capture(steps):
    rows = list(csv.reader(open(os.path.join(G, "r2_ppo_persist_raw.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__icc_request_hit_rate.pct", "sm__inst_executed.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    out = ["ncu --set full --clock-control none --cache-control none --import-source on -k regex:ppo_persist -s 1 -c 1  python tools/persist_once.py 256",
           "ppo_persist_kernel<false>, c2 networks (actor + 2 critics, 2x256), 256 envs x 300 steps = %d minibatch steps in the launch; B200" % steps,
           "(--cache-control none: L2 is NOT flushed between replays, i.e. the steady state the kernel runs in)", ""]
    for k in want:
        if k in d:
            out.append("%-90s %14s %s" % (k, d[k][0], d[k][1]))

    def num(k):
        v, u = d[k]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    dur_v, dur_u = d["gpu__time_duration.sum"]
    dur_us = float(dur_v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(dur_u.replace("second", "s"), 1)
    out += ["", "per minibatch step: %.1f us under ncu; DRAM traffic %.0f B read + %.0f B written per step (operand images, partials and flags "
            "live in L2)" % (dur_us / steps, rd / steps, wr / steps)]
    open(os.path.join(P, "r2_ppo_persist_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
    json.dump({"dram_bytes_per_launch": rd + wr, "launch": "%d minibatch steps (256 envs x 300 steps), ncu --set full --cache-control none" % steps,
               "steps_in_launch": steps}, open(os.path.join(P, "r2_ppo_persist_traffic.json"), "w"))
    print("\n".join(out[-3:]))


if __name__ == "__main__":
    launch_list()
    full_capture(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
