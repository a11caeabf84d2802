"""Synthetic evaluation harness at the sizes the headline metric is quoted on (test infrastructure).

BASELINE configs[2..4]: ViT-B/16 or ViT-L/14, K = 1000 prompts, ID stream vs an OOD stream of iNaturalist shape
(10 000 images).  There are no real weights, prompts or datasets here, so the streams are the
prototype harness of SURVEY.md section 8d: the bank rows are the centred oracle features of K seeded prototype
images, an ID image is prototype (i mod K) + noise * N(0, 1), an OOD image is N(0, 1 + noise^2) (the same
per-pixel variance, no prototype).  Everything is generated ON THE DEVICE from seeds (torch.Generator), slab by
slab, so 20 000 images never sit in host memory; the engine and the oracle see bit-identical tensors because the
slabs are regenerated from the same seeds for each of them.

The checker is oracle/clip_mcm_oracle.py (the restatement pinned to the unmodified reference by the golden
fixtures) run in fp32 on the GPU with TF32 off: the CPU needs ~20 min per 1 000 ViT-L/14 images.
"""
import numpy as np
import torch

SLAB = 500


class K1000Harness:
    def __init__(self, cfg_name, K=1000, noise=0.8, wseed=5, device="cuda"):
        from mcm_b200 import synth
        from oracle import clip_mcm_oracle as O
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.O = O
        self.cfg_name, self.K, self.noise = cfg_name, int(K), float(noise)
        self.dev = torch.device(device)
        self.cfg = synth.CFGS[cfg_name]
        self.sd = synth.synth_vision_state_dict(self.cfg, wseed)
        self.sd_gpu = {k: v.to(self.dev) for k, v in self.sd.items()}
        g = torch.Generator(device=self.dev)
        g.manual_seed(1000 + wseed)
        s = self.cfg.image_size
        self.protos = torch.randn((self.K, 3, s, s), generator=g, device=self.dev, dtype=torch.float32)
        with torch.no_grad():
            pf = torch.cat([O.image_features(self.protos[i:i + 50], self.sd_gpu, self.cfg) for i in range(0, self.K, 50)])
        self.bank = synth.centred_prototype_bank(pf.cpu().numpy())          # [K, P] fp32 unit rows
        bt = torch.from_numpy(self.bank).to(self.dev)
        self.bank_gpu = bt / bt.norm(dim=-1, keepdim=True)                  # utils/detection_util.py:231

    def slabs(self, name, n):
        """Device tensors [<= SLAB, 3, H, W] of stream `name` ("id" / "ood"), regenerated identically on every call."""
        s = self.cfg.image_size
        for k, s0 in enumerate(range(0, n, SLAB)):
            m = min(SLAB, n - s0)
            g = torch.Generator(device=self.dev)
            g.manual_seed((7 if name == "id" else 8) * 100003 + k)
            x = torch.randn((m, 3, s, s), generator=g, device=self.dev, dtype=torch.float32)
            if name == "id":
                x.mul_(self.noise)
                x.add_(self.protos[(torch.arange(s0, s0 + m, device=self.dev)) % self.K])
            else:
                x.mul_(float(np.sqrt(1.0 + self.noise ** 2)))
            yield x

    @torch.no_grad()
    def oracle_scores(self, n_id, n_ood, T=1, score="MCM", batch=125):
        out = []
        for name, n in (("id", n_id), ("ood", n_ood)):
            out.append(np.concatenate([self.O.ood_scores(x, self.sd_gpu, self.cfg, self.bank_gpu, T=T, score=score, batch=batch)
                                       for x in self.slabs(name, n)]))
        return out

    def engine_scores(self, eng, n_id, n_ood, batch, T=1.0, score="MCM"):
        out = []
        for name, n in (("id", n_id), ("ood", n_ood)):
            parts = []
            for x in self.slabs(name, n):
                for s0 in range(0, x.shape[0], batch):
                    parts.append(eng.score(x[s0:s0 + batch], T=T, score=score).clone())
                torch.cuda.synchronize()
            out.append(torch.cat(parts).cpu().numpy().astype(np.float32))
        return out
