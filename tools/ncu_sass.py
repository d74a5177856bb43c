#!/usr/bin/env python3
"""Top SASS instructions of a kernel by stall samples (needs `ncu --set full --import-source on`).

usage: python tools/ncu_sass.py report.ncu-rep [kernel-index] [top-n]
Prints pc, share of samples, the two main stall reasons, the SASS text and the source line it maps to.  A stall is
attributed to the instruction that could not issue, i.e. the CONSUMER of a pending load, not the load.
"""
import sys
from collections import defaultdict

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402

rep = ncu_report.load_report(sys.argv[1])
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
act = rep.range_by_idx(0).action_by_idx(kidx)
print("kernel:", act.name())
names = [n for n in act.metric_names() if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
by_pc = defaultdict(lambda: defaultdict(float))
for n in names:
    m = act.metric_by_name(n)
    ids = m.correlation_ids()
    for i in range(m.num_instances()):
        v = m.as_uint64(i)
        if v:
            pc = ids.as_uint64(i)
            by_pc[pc]["samp"] += v
            by_pc[pc][n.replace("smsp__pcsamp_warps_issue_stalled_", "")] += v
tot = sum(v["samp"] for v in by_pc.values()) or 1
base = min(by_pc) if by_pc else 0
for pc, v in sorted(by_pc.items(), key=lambda kv: -kv[1]["samp"])[:topn]:
    top = sorted(((k, x) for k, x in v.items() if k != "samp"), key=lambda t: -t[1])[:2]
    tops = " ".join(f"{k}:{x / v['samp'] * 100:.0f}%" for k, x in top)
    info = act.source_info(pc)
    where = f"{info.file_name().split('/')[-1]}:{info.line()}" if info else "?"
    try:
        sass = act.sass_by_pc(pc)
    except Exception:
        sass = "?"
    print(f"+{pc - base:06x} {v['samp'] / tot * 100:5.1f}%  [{tops}]  {sass}   <- {where}")
