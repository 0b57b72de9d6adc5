"""Event trace of one chained layer (group 0's CTA): needs a library built with NVCC_EXTRA=-DOVIS_CHAIN_EVTRACE."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L, decoder as D
from openvis_b200.synthetic import decoder_param_shapes, seeded_params
kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=100, nheads=8,
          dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
m = D.VideoMultiScaleMaskedTransformerDecoder(**kw)
m.load_state_dict(seeded_params(decoder_param_shapes("video", Q=100), 0))
m = m.cuda().eval(); m.use_cuda_graph = False; m.use_chain = True
g = torch.Generator(device="cuda").manual_seed(1)
x = [torch.randn(2, 256, 4 * 2 ** l, 6 * 2 ** l, generator=g, device="cuda") for l in range(3)]
mf = torch.randn(2, 256, 32, 48, generator=g, device="cuda")
m(x, mf)
c = next(iter(m._ws.values()))["chain"]
ch = c["chain"]
first, count = c["first"][3], c["count"][3]
for _ in range(3): ch.run(first, count)
tr = ch.run_traced(first, count).cpu().tolist()
t0 = tr[0]
print("phase starts:", [t - t0 for t in tr[:count + 1]])
ev = []
for slot in range(6000):
    v = tr[66 + slot]
    if v:
        kind, rest = slot // 2000, slot % 2000
        ev.append((v - t0, kind, rest // 160, (rest % 160) // 16, rest % 16))
names = {0: "load", 1: "mma ", 2: "epi "}
for t, kind, p, nt, sub in sorted(ev):
    print(f"{t:8d} {names[kind]} phase {p:2d} tile {nt} sub {sub}")
