#!/usr/bin/env python3
"""Per-source-line summary of an .ncu-rep (needs -lineinfo builds and `ncu --import-source on`).

usage: python tools/ncu_lines.py report.ncu-rep [kernel-index] [top-n]
Prints, for the top source lines by stall samples: samples, share, warp instructions executed, avg active threads.
"""
import sys
from collections import defaultdict

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402

rep = ncu_report.load_report(sys.argv[1])
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rng = rep.range_by_idx(0)
act = rng.action_by_idx(kidx)
print("kernel:", act.name(), "duration_us:", act.metric_by_name("gpu__time_duration.sum").as_double() / 1e3)
samples = act.metric_by_name("smsp__pcsamp_sample_buffer")  # may not exist
m_inst = act.metric_by_name("inst_executed")
m_tinst = act.metric_by_name("thread_inst_executed")
m_samp = act.metric_by_name("smsp__pcsamp_warps_issue_stalled_all") if "smsp__pcsamp_warps_issue_stalled_all" in act.metric_names() else None
names = [n for n in act.metric_names() if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
by_line = defaultdict(lambda: defaultdict(float))
n_inst = m_inst.num_instances()
pcs = m_inst.correlation_ids()
for i in range(n_inst):
    pc = pcs.as_uint64(i)
    info = act.source_info(pc)
    key = (info.file_name().split("/")[-1], info.line()) if info else ("?", 0)
    by_line[key]["inst"] += m_inst.as_uint64(i)
    by_line[key]["tinst"] += m_tinst.as_uint64(i)
for n in names:
    m = act.metric_by_name(n)
    ids = m.correlation_ids()
    for i in range(m.num_instances()):
        pc = ids.as_uint64(i)
        info = act.source_info(pc)
        key = (info.file_name().split("/")[-1], info.line()) if info else ("?", 0)
        v = m.as_uint64(i)
        by_line[key]["samp"] += v
        by_line[key][n.replace("smsp__pcsamp_warps_issue_stalled_", "")] += v
tot_s = sum(v["samp"] for v in by_line.values()) or 1
tot_i = sum(v["inst"] for v in by_line.values()) or 1
print(f"total warp-inst {tot_i:.3e}  total samples {tot_s:.0f}")
src_cache = {}
def src(f, l):
    try:
        if f not in src_cache:
            import glob
            cand = glob.glob(f"/root/repo/**/{f}", recursive=True)
            src_cache[f] = open(cand[0]).read().split("\n") if cand else []
        return src_cache[f][l - 1].strip()[:90]
    except Exception:
        return ""
for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1]["samp"])[:topn]:
    top = sorted(((k, x) for k, x in v.items() if k not in ("inst", "tinst", "samp")), key=lambda t: -t[1])[:2]
    tops = " ".join(f"{k}:{x / max(v['samp'], 1) * 100:.0f}%" for k, x in top)
    print(f"{f}:{l:4d} samp {v['samp'] / tot_s * 100:5.1f}%  inst {v['inst'] / tot_i * 100:5.1f}%  thr/inst {v['tinst'] / max(v['inst'], 1):4.1f}  [{tops}]  {src(f, l)}")
