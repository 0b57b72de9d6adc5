"""Two pixel-decoder passes (12 frames of 736 x 1280) for an ncu launch list: the second pass is the one to read."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openvis_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
from openvis_b200.synthetic import seeded_pixel_decoder_params

Nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ch, Hp, Wp = (256, 512, 1024, 2048), 736, 1280
dev = torch.device("cuda:0")
pd = MSDeformAttnPixelDecoder({f"res{i + 2}": ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(ch)})
pd.load_state_dict(seeded_pixel_decoder_params(2, in_channels=ch))
pd = pd.to(dev)
g = torch.Generator(device=dev).manual_seed(13)
feats = {f"res{i + 2}": torch.randn(Nf, c, Hp // (4 << i), Wp // (4 << i), generator=g, device=dev) for i, c in enumerate(ch)}
pd.forward_features(feats)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("pass2")
pd.forward_features(feats)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
