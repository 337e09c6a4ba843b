#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics per launch and the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__waves_per_multiprocessor',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def limiter_json(hdr, units, row, out_path, config):
    """profiles/r2_k3_limiter.json: what bench.py needs from the profiler and cannot measure itself — DRAM bytes per
    launch and the utilisation figures that name the kernel's real limiter (all for ONE launch = one step)."""
    import json

    def get(name, scale_units=True):
        if name not in hdr:
            return None
        i = hdr.index(name)
        try:
            v = float(row[i])
        except ValueError:
            return None
        if scale_units:
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
        return v
    stalls = {n[len("smsp__pcsamp_warps_issue_stalled_"):]: float(row[i]) for i, n in enumerate(hdr)
              if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("not_issued") and row[i] not in ("", "n/a")}
    tot = sum(stalls.values()) or 1.0
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
    sm = {"issue_slots_busy_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
          "lsu_data_pipe_pct": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
          "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False),
          "l2_throughput_pct": get("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
          "dram_throughput_pct": get("dram__throughput.avg.pct_of_peak_sustained_elapsed", False),
          "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active", False),
          "shared_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
          "shared_bank_conflict_wavefronts": get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False),
          "l2_hit_rate_pct": get("lts__t_sector_hit_rate.pct", False),
          "registers_per_thread": get("launch__registers_per_thread", False),
          "top_stalls_pct_of_samples": {k: round(100 * v / tot, 1) for k, v in top}}
    pipes = {"sm_lsu": sm["lsu_data_pipe_pct"] or 0, "sm_issue": sm["issue_slots_busy_pct"] or 0, "sm_alu": sm["alu_pipe_pct"] or 0,
             "l2": sm["l2_throughput_pct"] or 0, "hbm": sm["dram_throughput_pct"] or 0}
    name = max(pipes, key=pipes.get)
    out = {"config": config, "kernel": row[hdr.index("Kernel Name")][:60],
           "gpu_time_ms": get("gpu__time_duration.sum", False),
           "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
           "limiter": name, "limiter_frac": round(pipes[name] / 100, 4),
           "limiter_is": "the busiest unit of the launch (ncu pct_of_peak_sustained): sm_lsu = LSU data pipe (shared-memory "
                         "accumulators + posting loads), sm_issue = issue slots, sm_alu = integer pipe, l2, hbm",
           "sm": sm, "source": "ncu --set full --clock-control none, one launch of the kernel; see the .txt beside this file"}
    json.dump(out, open(out_path, "w"), indent=1)
    print("wrote", out_path, {k: out[k] for k in ("limiter", "limiter_frac", "dram_bytes_read")})


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    if "--json" in sys.argv:       # --json OUT.json --config '{"docs": ..., "queries": ..., "top_k": ..., "n_gpus": 1, "unique_terms": 120}'
        import json
        limiter_json(hdr, units, rows[2], sys.argv[sys.argv.index("--json") + 1], json.loads(sys.argv[sys.argv.index("--config") + 1]))
    print("== raw metrics (one column per captured launch)")
    name_i = hdr.index("Kernel Name")
    print("kernels:", [r[name_i][:40] for r in rows[2:]])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:78s} {units[i]:10s}", [r[i] for r in rows[2:]])
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, kernel, agg = None, 0, []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Kernel Name':
            kernel += 1
            continue
        if kernel > 1:
            break
        if r[0].isdigit() and len(r) > 8:
            try:
                agg.append((cur, int(r[0]), r[1].strip()[:78], float(r[6] or 0), float(r[7] or 0)))
            except ValueError:
                pass   # a source line whose text contains the csv separator
    ti = sum(a[4] for a in agg) or 1
    ts = sum(a[3] for a in agg) or 1
    print(f"\n== hottest source lines of the first captured launch (warp instructions {ti:.0f}, stall samples {ts:.0f})")
    for a in sorted(agg, key=lambda a: -a[4])[:n_lines]:
        print(f"{a[0]:16s}:{a[1]:4d} inst={a[4] / ti * 100:5.1f}% samp={a[3] / ts * 100:5.1f}%  {a[2]}")


if __name__ == "__main__":
    main()
