import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_decoder_gpu as TD
from oracle import decoder_ref as O
T, Hp, Wp = 2, 64, 96
m, ref, out = TD.run_case("frame", T, Hp, Wp)
for i in range(9):
    a, b = out["aux_outputs"][i], ref["aux_outputs"][i]
    print(i, TD.frac_within(a["pred_masks"].cpu(), b["pred_masks"], 0.25), (a["pred_logits"].cpu() - b["pred_logits"]).abs().max().item())
print("final", TD.frac_within(out["pred_masks"].cpu(), ref["pred_masks"], 0.25))
