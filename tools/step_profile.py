"""In-situ kernel timeline of training steps at the bench shape (torch.profiler / CUPTI, warm caches, real overlap):
per-kernel totals, GPU busy time per stream and the idle gaps on the main stream.  Not a bench number."""
import os, sys, json, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench as B
import mmdfn_b200
from mmdfn_b200.dp import FlatAdamTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(2021)
model = mmdfn_b200.DialogueGNNModel(
    "LSTM", B.D_T, 150, 150, 100, 100, 100, 100, n_speakers=B.SPEAKERS, max_seq_len=200, window_past=10, window_future=10,
    n_classes=B.CLASSES, dropout=B.DROPOUT, graph_type="GDF", alpha=0.2, lamda=0.5, D_m_v=B.D_V, D_m_a=B.D_A, modals="avl",
    att_type="concat_subsequently", Deep_GCN_nlayers=B.LAYERS, use_speaker=False, reason_flag=True, use_crn_speaker=True,
    speaker_weights=B.SPK_W)
model = model.to(dev).train()
loss_fn = mmdfn_b200.FocalLoss(gamma=B.GAMMA, alpha=B.class_weights().to(dev))
trainer = FlatAdamTrainer(model, loss_fn, lr=B.LR, weight_decay=B.L2)
batches = [tuple(x.to(dev) for x in b) for b in B.make_batches(2, seed0=1000)]
lengths = [B.UTT] * B.DIALOGUES_PER_GPU


def step(i):
    t, a, v, q, u, lab = batches[i % 2]
    return trainer.step(t, q, u, lengths, a, v, lab, sum(lengths))


GRAPH = "--graph" in sys.argv          # profile replays of the captured whole-step graph (the bench's launch mode)
for i in range(5):
    step(i)
torch.cuda.synchronize()
if GRAPH:
    t, a, v, q, u, lab = batches[0]
    trainer.capture(t, q, u, lengths, a, v, lab, sum(lengths), warmup=0)

    def step(i):
        t, a, v, q, u, lab = batches[i % 2]
        return trainer.replay(t, q, u, a, v, lab, lengths)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
NSTEP = 4
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(NSTEP):
        step(i)
        torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "step_trace.json")
prof.export_chrome_trace(out)
allev = json.load(open(out))["traceEvents"]
ev = [e for e in allev if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# keep the last step: kernels after the (NSTEP-1)th adam kernel
adam = [i for i, e in enumerate(ev) if "adam_kernel" in e["name"] or "adam_dev_kernel" in e["name"]]
last = ev[adam[-2] + 1: adam[-1] + 1]
t0, t1 = last[0]["ts"], last[-1]["ts"] + last[-1]["dur"]
print("last step: %d kernels, span %.1f us" % (len(last), t1 - t0))
tot = collections.Counter(); cnt = collections.Counter()
streams = collections.defaultdict(list)
for e in last:
    n = re.sub(r"[\(<].*", "", e["name"]).replace("void ", "").replace("mmdfn::", "")
    if "umma_gemm" in e["name"]:
        n += re.search(r"<\(?(?:int\))?(\d)", e["name"]).group(0)[-2:] if re.search(r"<\(?(?:int\))?(\d)", e["name"]) else ""
    tot[n] += e["dur"]; cnt[n] += 1
    streams[e["args"].get("stream")].append(e)
print("sum of kernel durations %.1f us" % sum(tot.values()))
for n, v in tot.most_common(16):
    print("  %-38s %3d  %8.1f us" % (n, cnt[n], v))
for s, es in streams.items():
    busy = sum(e["dur"] for e in es)
    print("stream %s: %d kernels, busy %.1f us, first @%.1f last end @%.1f" % (s, len(es), busy, es[0]["ts"] - t0, es[-1]["ts"] + es[-1]["dur"] - t0))
# idle time: union of busy intervals over all streams
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in last)
cover, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        cover += cur_e - cur_s; gaps.append((s - cur_e, cur_e - t0)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
cover += cur_e - cur_s
print("GPU busy (any stream) %.1f us of %.1f us span; %d gaps, total gap %.1f us, mean %.2f us" % (cover, t1 - t0, len(gaps), sum(g for g, _ in gaps), sum(g for g, _ in gaps) / max(1, len(gaps))))
print("largest gaps (us @offset):", sorted(gaps, reverse=True)[:8])
with open(os.path.join(ROOT, "gpurun_out", "step_timeline.txt"), "w") as f:
    for e in last:
        g = e["args"].get("grid", [0, 0, 0])
        f.write("%9.1f %8.1f  s%-3s g%-5d %s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), g[0] * g[1] * g[2], e["name"][:90]))

# what sits in the largest gap: every device-side trace event (any category) that overlaps it
if gaps:
    g, off = max(gaps)
    lo, hi = t0 + off - 1.0, t0 + off + g + 1.0
    print("events overlapping the largest gap (%.1f us @%.1f):" % (g, off))
    for e in allev:
        if e.get("ph") == "X" and e.get("cat") not in ("cpu_op", "python_function", "user_annotation", "cuda_runtime", "cuda_driver") \
                and e["ts"] < hi and e["ts"] + e.get("dur", 0) > lo:
            print("   %-14s %9.1f %7.1f  %s" % (e.get("cat"), e["ts"] - t0, e.get("dur", 0), e["name"][:80]))
