"""One pass of the CLIP visual tower (ClipVisualEncoder, 256 crops) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/clip_tower_launches.csv python tools/prof_clip_tower.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200.clip_adapter import ClipVisualEncoder
from openvis_b200.synthetic import seeded_clip_visual_params
v = ClipVisualEncoder().load_state_dict(seeded_clip_visual_params(3))
x = (torch.rand(256, 3, 224, 224, device="cuda") * 255).half()
for _ in range(2 if len(sys.argv) > 1 else 1):
    f = v(x)
torch.cuda.synchronize()
print(f.shape, float(f.abs().mean()))
