#!/usr/bin/env python3
"""profiles/traffic.json from ncu summaries (tools/ncu_summary.py JSON of 262144-stream captures).

usage: python tools/make_traffic.py <roundtrip48 summary.json> [<decode48 summary.json>] [streams]
Per kernel: (dram__bytes_read.sum + dram__bytes_write.sum) of one launch / streams = DRAM bytes per stream-frame.  bench.py
scales that figure by the streams of its own launch for `roofline.traffic`.  Kernels bench.py times together (the post
filter with the synthesis kernel, the three bitstream kernels) are summed under bench.py's name for the group.
The decoder kernels' figures come from the decode48 capture when one is given: inside a round trip the first decoder
kernel is billed for the write-back of what the encoder's last kernels left dirty in L2.
"""
import json
import sys

args = [a for a in sys.argv[1:] if not a.isdigit()]
nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
streams = nums[0] if nums else 262144


def per_kernel(path):
    per = {}
    for r in json.load(open(path)):
        name = "lc3b::" + r["kernel"].split("<")[0].replace("synth_classic_kernel", "synth_kernel")
        per[name] = per.get(name, 0.0) + r["dram_bytes"] / streams
    groups = {
        "lc3b::synth_kernel+ltpf_kernel": ("lc3b::synth_kernel", "lc3b::ltpf_kernel"),
        "lc3b::enc_bs_prepare+range_coder+bs_finish": ("lc3b::enc_bs_prepare_kernel", "lc3b::enc_range_coder_kernel", "lc3b::enc_bs_finish_kernel"),
    }
    for name, parts in groups.items():
        have = [p for p in parts if p in per]
        if have and parts[0] in per:                      # the post-filter kernel is not launched when min_nbytes rules it out
            per[name] = sum(per[p] for p in have)
    return {k: round(v, 1) for k, v in per.items()}


rt = per_kernel(args[0])
dec = per_kernel(args[1]) if len(args) > 1 else {k: v for k, v in rt.items() if not k.startswith("lc3b::enc_")}
enc = {k: v for k, v in rt.items() if k.startswith("lc3b::enc_")}
out = {"dram_bytes_per_stream_frame": {"roundtrip48": {**rt, **dec}, "decode48": dec, "encode48": enc},
       "source": f"{' + '.join(args)} (ncu --set full, {streams} streams, 48 kHz 10 ms 150 B; dram__bytes_read.sum + dram__bytes_write.sum per "
                 "launch / streams). encode48 (120 B frames) reuses the 150 B figures.",
       "streams_profiled": streams}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out["dram_bytes_per_stream_frame"]["decode48"], indent=1))
