#!/usr/bin/env python3
"""Summarise every kernel of an .ncu-rep (one `ncu --set full` capture) as markdown + JSON.

usage: python tools/ncu_summary.py report.ncu-rep [out.md] [out.json]
Per kernel: duration, DRAM bytes read + written (the `traffic` figure of bench.py's roofline), achieved DRAM GB/s,
issue rate, occupancy, registers, shared memory, and the top warp-stall reasons from the PC samples.
"""
import json
import sys

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402


def main():
    rep = ncu_report.load_report(sys.argv[1])
    rng = rep.range_by_idx(0)
    rows = []
    for k in range(rng.num_actions()):
        act = rng.action_by_idx(k)

        def g(name, default=float("nan")):
            try:
                return act.metric_by_name(name).as_double()
            except Exception:
                return default

        stalls = {}
        for n in act.metric_names():
            if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
                m = act.metric_by_name(n)
                stalls[n[33:]] = sum(m.as_uint64(i) for i in range(m.num_instances()))
        tot = sum(stalls.values()) or 1
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
        dur_us = g("gpu__time_duration.sum") / 1e3
        dram = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
        rows.append({
            "kernel": act.name(),
            "grid": [int(g("launch__grid_size"))], "block": [int(g("launch__block_size"))],
            "duration_us": dur_us,
            "dram_bytes": dram,
            "dram_bytes_read": g("dram__bytes_read.sum"),
            "dram_bytes_write": g("dram__bytes_write.sum"),
            "dram_gbs": dram / (dur_us * 1e-6) / 1e9 if dur_us > 0 else 0.0,
            "ipc_per_sm": g("sm__inst_executed.avg.per_cycle_elapsed"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": int(g("launch__registers_per_thread")),
            "smem_per_block": int(g("launch__shared_mem_per_block_dynamic", 0) + g("launch__shared_mem_per_block_static", 0)),
            "warp_inst": g("smsp__inst_executed.sum"),
            "thread_inst_per_warp_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "thread_inst_pred_on_per_warp_inst": g("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
            "stalls": {k2: v / tot for k2, v in top},
        })
    md = ["| kernel | grid x block | time (us) | DRAM MB | DRAM GB/s | IPC/SM | active warps % | regs | smem/CTA | thr/inst | top stalls |",
          "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    for r in rows:
        st = ", ".join(f"{k} {v * 100:.0f}%" for k, v in r["stalls"].items())
        md.append(f"| `{r['kernel']}` | {r['grid'][0]} x {r['block'][0]} | {r['duration_us']:.0f} | {r['dram_bytes'] / 1e6:.1f} | "
                  f"{r['dram_gbs']:.0f} | {r['ipc_per_sm']:.2f} | {r['warps_active_pct']:.0f} | {r['registers']} | "
                  f"{r['smem_per_block']} | {r['thread_inst_per_warp_inst']:.1f} | {st} |")
    text = "\n".join(md)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    if len(sys.argv) > 3:
        json.dump(rows, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
