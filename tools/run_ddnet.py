"""DDnet deep demosaic at the mid-scale size (8 mosaics 512x512): time per call, per-layer conv profile, and the
accuracy of the TF32 tensor-core engine against the fp32 FFMA engine on the same device."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from adaptivepnp_sci_b200 import engine as eng_mod
from adaptivepnp_sci_b200.fastdvdnet_adapter import DataParallelLike
from adaptivepnp_sci_b200.network_demosaicking import DDnet
from adaptivepnp_sci_b200.synthetic import ddnet_synthetic_state_dict, make_case


def build(impl):
    os.environ["SCI_CONV_IMPL"] = impl
    m = DataParallelLike(DDnet())
    m.load_state_dict({"module." + k: v for k, v in ddnet_synthetic_state_dict().items()})
    return m.eval().cuda()


H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = 8
_, _, orig = make_case(H, W, B, 3000, bayer=True)
mosaic = torch.from_numpy(orig).permute(2, 0, 1).contiguous().cuda()
tc, ref = build("tc").module.engine(), build("ref").module.engine()
y_tc = tc.forward(mosaic).clone()
y_ref = ref.forward(mosaic).clone()
err = float((y_tc - y_ref).abs().max())
for _ in range(3):
    tc.forward(mosaic)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    tc.forward(mosaic)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
tc.profile = []
tc.forward(mosaic)
torch.cuda.synchronize()
rows = {}
for a, b, fl, tag in tc.profile:
    t = a.elapsed_time(b)
    r = rows.setdefault(tag, [0, 0.0, 0.0])
    r[0] += 1; r[1] += t; r[2] += fl
tc.profile = None
conv_ms = sum(r[1] for r in rows.values())
flops = sum(r[2] for r in rows.values())
out = {"size": [B, H, W], "ms_per_call": ms, "conv_ms_instrumented": conv_ms, "alg_gflop": flops / 1e9,
       "alg_tflops": flops / ms / 1e9, "max_abs_tc_vs_fp32": err,
       "layers": [{"layer": k, "launches": v[0], "ms": round(v[1], 4), "tflops": round(v[2] / v[1] / 1e9, 1)}
                  for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ddnet_%d.json" % H, "w"), indent=1)
