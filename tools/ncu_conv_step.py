"""One pipeline call (B frame pairs of 1080p, default 8 = one sub-batch of the bench step) between cudaProfilerStart / Stop, for
    ncu --profile-from-start off --kernel-name regex:conv_split --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/conv_traffic.csv python tools/ncu_conv_step.py
and, with a CSV argument, the summary of such a log:  python tools/ncu_conv_step.py gpurun_out/conv_traffic.csv [B]  ->  JSON with the
measured DRAM bytes of all convolution launches of one bench step (2 sub-batches of 8) next to their algorithmic bytes."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]

if len(sys.argv) > 1 and sys.argv[1].endswith(".csv"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    per = {}
    for r in rows[1:]:
        if "conv_split" not in r[ik]:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
                 "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1.0)
        per.setdefault(r[iid], {})[r[im]] = v * scale
    rd = sum(p.get("dram__bytes_read.sum", 0.0) for p in per.values())
    wr = sum(p.get("dram__bytes_write.sum", 0.0) for p in per.values())
    ms = sum(p.get("gpu__time_duration.sum", 0.0) for p in per.values())
    k = 16.0 / B                                     # the bench step = 16 frame pairs
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over every conv_split_kernel launch "
                                "of ONE pipeline call on %d frame pairs (tools/ncu_conv_step.py), scaled to the 16 pairs of a bench step" % B,
                      "launches_per_step": int(len(per) * k), "dram_read_bytes_per_step": rd * k, "dram_write_bytes_per_step": wr * k,
                      "dram_bytes_per_step": (rd + wr) * k, "dram_bytes_per_launch": (rd + wr) / max(len(per), 1),
                      "ncu_serialized_ms_per_step": ms * k}))
    sys.exit(0)

import torch
from fvfi.pipeline import FusionPipeline
from fvfi import synth as fp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=24)
pipe.max_batch = 8
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()
pipe(d1, d2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe(d1, d2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
