"""Where does the fp16 mode's score error come from?  (test infrastructure: runs the oracle, CPU is enough)

The oracle tower in fp32 with ONE class of tensors rounded to fp16 at a time -- the sites where the engine's `fp16` mode
rounds: the GEMM weights, the patch pixels, the A operands of q/k/v and fc1, the q/k/v, attention-output and `hid`
activations, the attention probabilities, and the residual stream as an fp16 (hi, lo) pair -- on 96 ID + 96 OOD images
of the K = 1000 ViT-B/16 prototype harness (tests/k1000_harness.py).  Prints the rms / max score error in units of the
score spread.  Result of the round-2 run: weights 1.15e-3, q/k/v 0.69e-3, A of q/k/v 0.65e-3, A of
fc1 0.53e-3, hid 0.53e-3, pixels 0.52e-3, attention output 0.36e-3, probabilities 0.14e-3, residual pair < 1e-5; all
together 1.9e-3 rms.  No single site dominates, so no partial split gets fp32-class scores.

    python tests/precision_sensitivity_report.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from k1000_harness import K1000Harness
from oracle import clip_mcm_oracle as O
torch.set_num_threads(8)
t0 = time.time()
h = K1000Harness('ViT-B/16', K=1000, noise=0.8, device='cpu')
print('harness', time.time() - t0, flush=True)
cfg, sd = h.cfg, h.sd
N = 96
xid = next(h.slabs('id', N)); xood = next(h.slabs('ood', N))
x_all = torch.cat([xid, xood])
q16 = lambda t: t.half().float()
ident = lambda t: t

def forward(pix, Q):
    """Q: dict site -> quantizer"""
    g = lambda s: Q.get(s, ident)
    sdq = {k: (g('W')(v) if (k.endswith('proj.weight') or 'mlp.fc' in k and k.endswith('weight') or 'patch_embedding' in k) else v) for k, v in sd.items()}
    x = O.vision_embeddings(g('pix')(pix), sdq, cfg)
    x = O._ln(x, sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"], cfg.eps)
    B, S, D = x.shape; H = cfg.heads; dh = D // H
    for i in range(cfg.layers):
        pre = f"vision_model.encoder.layers.{i}."
        x = g('resid')(x)
        hh = g('a_qkv')(O._ln(x, sd[pre + "layer_norm1.weight"], sd[pre + "layer_norm1.bias"], cfg.eps))
        q = g('qkv')(O._lin(hh, sdq[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.q_proj.bias"]))
        k = g('qkv')(O._lin(hh, sdq[pre + "self_attn.k_proj.weight"], sd[pre + "self_attn.k_proj.bias"]))
        v = g('qkv')(O._lin(hh, sdq[pre + "self_attn.v_proj.weight"], sd[pre + "self_attn.v_proj.bias"]))
        q = q.view(B, S, H, dh).transpose(1, 2); k = k.view(B, S, H, dh).transpose(1, 2); v = v.view(B, S, H, dh).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * (dh ** -0.5)
        m = s.max(dim=-1, keepdim=True).values
        p = torch.exp(s - m)
        den = p.sum(dim=-1, keepdim=True)
        o = (g('P')(p) @ v) / den
        o = g('attn')(o.transpose(1, 2).reshape(B, S, D))
        x = x + O._lin(o, sdq[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
        x = g('resid')(x)
        hh = g('a_fc1')(O._ln(x, sd[pre + "layer_norm2.weight"], sd[pre + "layer_norm2.bias"], cfg.eps))
        hh = O._lin(hh, sdq[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
        hh = g('hid')(hh * torch.sigmoid(1.702 * hh))
        x = x + O._lin(hh, sdq[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    pooled = O._ln(x[:, 0], sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"], cfg.eps)
    return pooled @ sd["visual_projection.weight"].t()

bank = torch.from_numpy(h.bank); bank = bank / bank.norm(dim=-1, keepdim=True)
def scores(f):
    return O.scores_from_features(f, bank, T=1, score="MCM")
with torch.no_grad():
    base = np.concatenate([scores(forward(x_all[i:i + 32], {})) for i in range(0, 2 * N, 32)])
    print('base std', base.std(), 'mean', base.mean(), time.time() - t0, flush=True)
    # hi/lo pair for the residual (what the engine stores): 22 bits
    pair = lambda t: (t.half().float() + (t - t.half().float()).half().float())
    for name, Q in [('W', {'W': q16}), ('pix', {'pix': q16}), ('a_qkv', {'a_qkv': q16}), ('qkv', {'qkv': q16}), ('P', {'P': q16}), ('attn', {'attn': q16}),
                    ('a_fc1', {'a_fc1': q16}), ('hid', {'hid': q16}), ('resid_pair', {'resid': pair}),
                    ('all', {k: q16 for k in ('W', 'pix', 'a_qkv', 'qkv', 'P', 'attn', 'a_fc1', 'hid')})]:
        got = np.concatenate([scores(forward(x_all[i:i + 32], Q)) for i in range(0, 2 * N, 32)])
        d = got - base
        print(f"{name:10s} rms err / std = {np.sqrt((d**2).mean())/base.std():.5f}  max = {np.abs(d).max()/base.std():.5f}  bias = {d.mean()/base.std():+.5f}", time.time() - t0, flush=True)
