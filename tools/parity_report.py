"""Prints a parity table (CUDA path vs CPU oracle) for several shapes; run on the GPU box.
python tools/parity_report.py [out.txt]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_decoder_gpu import build, unpack_bits, frac_within
from oracle import decoder_ref as O

CASES = [("frame", 5, 384, 640, 100), ("video", 5, 384, 640, 100), ("san_frame", 5, 384, 640, 100), ("san_video", 3, 384, 640, 100),
         ("frame", 2, 64, 64, 100), ("video", 3, 64, 96, 100), ("san_frame", 2, 96, 160, 200), ("frame", 1, 96, 64, 200),
         ("frame", 2, 736, 1280, 100), ("video", 2, 736, 1280, 100)]
lines = []
for kind, T, Hp, Wp, Q in CASES:
    m, P = build(kind, Q, 0)
    x, mf = O.seeded_inputs(T, Hp, Wp)
    t0 = time.time(); ref = O.decoder_forward(P, x, mf, kind=kind); tc = time.time() - t0
    m.debug_capture = []
    out = m([t.cuda() for t in x], mf.cuda()); torch.cuda.synchronize()
    sizes = [(Hp // 32 * 2 ** l) * (Wp // 32 * 2 ** l) for l in range(3)]
    agree = []
    for hidx, level, bits, flags in m.debug_capture:
        keys = sizes[level] * (T if kind.endswith("video") else 1)
        agree.append((unpack_bits(bits, keys).cpu() == ref["attn_masks"][hidx]).float().mean().item())
    pm, rm = out["pred_masks"].cpu(), ref["pred_masks"]
    err = (pm - rm).abs()
    s = f"{kind:9s} T={T} {Hp}x{Wp} Q={Q} cpu={tc:.2f}s | mask agree min {min(agree):.5f} last {agree[-1]:.5f} | pred_masks maxerr {err.max():.3f} p99.9 {err.flatten().kthvalue(int(err.numel()*0.999)).values:.3f} within0.25 {frac_within(pm, rm, 0.25):.5f} sign {((pm>0)==(rm>0)).float().mean():.5f}"
    if "pred_logits" in ref:
        s += f" | logits maxerr {(out['pred_logits'].cpu()-ref['pred_logits']).abs().max():.4f} top1 {(out['pred_logits'].cpu().argmax(-1)==ref['pred_logits'].argmax(-1)).float().mean():.4f}"
    if "class_attn_biases" in ref:
        e = (out['class_attn_biases'].cpu()-ref['class_attn_biases']).abs()
        s += f" | biases maxerr {e.max():.4f} within3e-2 {(e<=3e-2).float().mean():.5f}"
    if "pred_embeds" in ref:
        e = (out['pred_embeds'].cpu()-ref['pred_embeds']).abs()
        s += f" | embeds maxerr {e.max():.4f} within3e-2 {(e<=3e-2).float().mean():.5f}"
    print(s, flush=True); lines.append(s)
    del m, out; torch.cuda.empty_cache()
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write("\n".join(lines) + "\n")
