"""MSDeformAttn forward at the pixel decoder's shape: N frames of 736 x 1280 (levels 23x40, 46x80, 92x160), 8 heads x 32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shapes = torch.tensor([(23, 40), (46, 80), (92, 160)], dtype=torch.long)
S = int(shapes.prod(1).sum()); M, D, Lv, P = 8, 32, 3, 4
g = torch.Generator(device="cuda").manual_seed(0)
value = torch.rand(N, S, M, D, generator=g, device="cuda")
loc = torch.rand(N, S, M, Lv, P, 2, generator=g, device="cuda")
w = torch.rand(N, S, M, Lv, P, generator=g, device="cuda"); w = w / w.sum((-1, -2), keepdim=True)
start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1])).cuda(); shp = shapes.cuda()
for _ in range(2): out = L.ms_deform_attn_forward(value, shp, start, loc, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = L.ms_deform_attn_forward(value, shp, start, loc, w)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
hbm = (value.numel() + loc.numel() + w.numel() + out.numel()) * 4
gather = N * S * M * Lv * P * 4 * D * 4
print(f"N={N} S=Lq={S}: {ms*1e3:.0f} us per call; compulsory HBM {hbm/1e6:.0f} MB -> {hbm/ms/1e6:.0f} GB/s; gathered {gather/1e9:.1f} GB from L2 -> {gather/ms/1e9:.1f} TB/s")
