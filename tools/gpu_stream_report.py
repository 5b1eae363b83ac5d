"""Per-stream concurrency report of one graph-replayed training step (Kineto/CUPTI): how busy every stream is, how much of
the step has 1, 2, 3... kernels in flight, the largest idle gaps of the busiest stream, and the kernels on it.
usage: python tools/gpu_stream_report.py [H W]"""
import os, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from b200seg.module import B200SegModule
from b200seg.optim import FusedSGD

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 2048)
torch.manual_seed(0)
net = B200SegModule("ocrnet.HRNet_Mscale", 19).cuda().train()
with torch.no_grad():
    for n_, p_ in net.named_parameters():
        if p_.dim() == 4 and n_.startswith("backbone"):
            p_.normal_(0, (2.0 / (p_.shape[1] * p_.shape[2] * p_.shape[3])) ** 0.5)
opt = FusedSGD(net.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
images = torch.randn(1, 3, H, W, device="cuda")
gts = torch.randint(0, 19, (1, H, W), device="cuda")


def step():
    opt.zero_grad(set_to_none=True)
    loss = net({"images": images, "gts": gts})
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()

raw = prof.profiler.kineto_results.events()
ev = []
for e in raw:
    try:
        if "cuda" not in str(e.device_type()).lower():
            continue
        t0 = e.start_ns() if hasattr(e, "start_ns") else e.start_us() * 1000
        d = e.duration_ns() if hasattr(e, "duration_ns") else e.duration_us() * 1000
        ev.append((int(e.device_resource_id()), int(t0), int(t0 + d), e.name().split("(")[0][:48]))
    except Exception:
        continue
if not ev:
    print("no device events captured")
    sys.exit(0)
T0, T1 = min(a for _, a, _, _ in ev), max(b for _, _, b, _ in ev)
span = (T1 - T0) / 1e6
print("events %d  span %.3f ms" % (len(ev), span))
per = defaultdict(list)
for s, a, b, n in ev:
    per[s].append((a, b, n))
print("%-8s %7s %10s %10s" % ("stream", "kernels", "busy ms", "busy/span"))
for s, lst in sorted(per.items(), key=lambda kv: -sum(b - a for a, b, _ in kv[1])):
    busy = sum(b - a for a, b, _ in lst) / 1e6
    print("%-8d %7d %10.3f %9.1f%%" % (s, len(lst), busy, 100 * busy / span))
# concurrency histogram: sweep over start/end points
pts = sorted([(a, 1) for _, a, _, _ in ev] + [(b, -1) for _, _, b, _ in ev])
level, last, hist = 0, T0, defaultdict(int)
for t, d in pts:
    hist[level] += t - last
    last, level = t, level + d
print("kernels in flight -> share of the step:", {k: "%.1f%%" % (100.0 * v / (T1 - T0)) for k, v in sorted(hist.items())})
busiest = max(per.items(), key=lambda kv: sum(b - a for a, b, _ in kv[1]))[1]
busiest.sort()
gaps = sorted(((busiest[i + 1][0] - busiest[i][1], busiest[i][2], busiest[i + 1][2]) for i in range(len(busiest) - 1)),
              reverse=True)[:12]
print("largest idle gaps on the busiest stream (us, after -> before):")
for g, a, b in gaps:
    print("  %8.1f  %s -> %s" % (g / 1e3, a, b))
agg = defaultdict(lambda: [0, 0])
for a, b, n in busiest:
    agg[n][0] += 1
    agg[n][1] += b - a
print("kernels on the busiest stream:")
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("  %-48s n=%5d total %8.3f ms" % (n, c, d / 1e6))

# union coverage per kernel class: share of the step during which at least one kernel of the class is in flight, and the
# summed durations (a class whose union ~ its sum runs serialised; a sum far above the union overlaps with itself)
def union(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot


classes = {"tensor-core conv (halo/igemm/wgrad)": ("conv3x3_halo", "conv_igemm", "wgrad_igemm"),
           "wgrad_reduce": ("wgrad_reduce",), "batchnorm": ("bn_",), "resample/fuse": ("fuse_fwd", "upsample", "masked_accum"),
           "everything": ("",)}
print("%-40s %8s %10s %10s" % ("class", "kernels", "sum ms", "union ms"))
for cname, keys in classes.items():
    iv = [(a, b) for _, a, b, n in ev if any(k in n for k in keys)]
    print("%-40s %8d %10.3f %10.3f" % (cname, len(iv), sum(b - a for a, b in iv) / 1e6, union(iv) / 1e6))
byname = defaultdict(lambda: [0, 0])
for _, a, b, n in ev:
    byname[n][0] += 1
    byname[n][1] += b - a
print("per kernel (all streams):")
for n, (c, d) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:16]:
    print("  %-48s n=%5d total %8.3f ms avg %6.1f us" % (n, c, d / 1e6, d / c / 1e3))
