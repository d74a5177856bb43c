#!/usr/bin/env python3
"""profiles/traffic.json from an ncu summary (tools/ncu_summary.py JSON of a 262144-stream roundtrip48 capture).

usage: python tools/make_traffic.py profiles/<tag>_roundtrip48_ncu_summary.json [streams]
Per kernel: (dram__bytes_read.sum + dram__bytes_write.sum) of one launch / streams = DRAM bytes per stream-frame.  bench.py
scales that figure by the streams of its own launch for `roofline.traffic`.  Kernels bench.py times together (the post
filter with the synthesis kernel, the three bitstream kernels) are summed under bench.py's name for the group.
"""
import json
import sys

src = sys.argv[1]
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
rows = json.load(open(src))
per = {}
for r in rows:
    per["lc3b::" + r["kernel"].split("<")[0]] = per.get("lc3b::" + r["kernel"].split("<")[0], 0.0) + r["dram_bytes"] / streams
groups = {
    "lc3b::synth_kernel+ltpf_kernel": ("lc3b::synth_kernel", "lc3b::ltpf_kernel"),
    "lc3b::enc_bs_prepare+range_coder+bs_finish": ("lc3b::enc_bs_prepare_kernel", "lc3b::enc_range_coder_kernel", "lc3b::enc_bs_finish_kernel"),
}
for name, parts in groups.items():
    if all(p in per for p in parts):
        per[name] = sum(per[p] for p in parts)
per = {k: round(v, 1) for k, v in per.items()}
dec = {k: v for k, v in per.items() if not k.startswith("lc3b::enc_")}
enc = {k: v for k, v in per.items() if k.startswith("lc3b::enc_")}
out = {"dram_bytes_per_stream_frame": {"roundtrip48": per, "decode48": dec, "encode48": enc},
       "source": f"{src} (ncu --set full, {streams} streams, 48 kHz 10 ms 150 B; dram__bytes_read.sum + dram__bytes_write.sum per "
                 "launch / streams). encode48 (120 B frames) reuses the 150 B figures.",
       "streams_profiled": streams}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(per, indent=1))
