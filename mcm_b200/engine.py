"""Host side of the B200 MCM engine: a thin PyTorch-facing wrapper over the C ABI.

PyTorch is plumbing here (device buffers, streams); every FLOP of the path runs in the
hand-written sm_100a kernels behind ``include/mcm_b200.h``.

Two seams of the reference are re-created on top of :class:`McmEngine`:

* seam 2, the model object returned by ``set_model_clip`` (``utils/train_eval_util.py:15-36``):
  :class:`B200ClipNet` -- ``.eval()``, ``.get_image_features(pixel_values=)``,
  ``.get_text_features(input_ids=, attention_mask=)``;
* seam 1, ``get_ood_scores_clip`` (``utils/detection_util.py:209-249``): in
  ``mcm_b200/detection_util.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Mapping, Optional

import numpy as np
import torch

from . import _lib
from .synth import CFGS, CKPT_TO_CFG, VisionCfg

__all__ = ["McmEngine", "B200ClipNet", "VisionCfg", "CFGS"]


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _score_kind(score: str) -> int:
    try:
        return _lib.SCORE_KINDS[score]
    except KeyError:
        raise ValueError(f"score {score!r} is not part of the B200 path (choices: {sorted(_lib.SCORE_KINDS)}); "
                         "'maha' goes through get_Mahalanobis_score / McmEngine.maha_score (utils/detection_util.py:148-207)") from None


class McmEngine:
    """One CLIP vision tower + prompt bank resident on one B200.

    Replaces, for the MCM path, ``CLIPModel.from_pretrained(...).cuda()``
    (``utils/train_eval_util.py:23-26``) and the arithmetic of
    ``utils/detection_util.py:225-248``.
    """

    def __init__(self, cfg: VisionCfg, max_batch: int = 256, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("mcm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = _lib.load()
        self.cfg = cfg
        self.max_batch = int(max_batch)
        self.device = torch.device("cuda", int(device))
        self._ccfg = _lib.McmConfig(cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.heads, cfg.mlp, cfg.proj,
                                    cfg.eps, self.max_batch, int(device))
        self._h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            _lib.check(self._lib.mcm_create(C.byref(self._ccfg), C.byref(self._h)), None)
        self.K = 0
        self.precision = "fp16"

    # ------------------------------------------------------------------ lifecycle --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mcm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.check(rc, self._h)

    # -------------------------------------------------------------------- weights --
    def load_state_dict(self, sd: Mapping[str, torch.Tensor]) -> list:
        """Feed a HuggingFace ``CLIPModel.state_dict()`` (fp32).  Returns the keys consumed; text
        tower / logit_scale entries are ignored.  Raises if a vision tensor is missing."""
        used_keys = []
        used = C.c_int32(0)
        for k, v in sd.items():
            if not (k.startswith("vision_model.") or k == "visual_projection.weight"):
                continue
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to(torch.float32).contiguous()
            self._check(self._lib.mcm_load_weight(self._h, k.encode(), _ptr(t), t.numel(), C.byref(used)))
            if used.value:
                used_keys.append(k)
        self._check(self._lib.mcm_finalize_weights(self._h))
        return used_keys

    @classmethod
    def from_state_dict(cls, sd, cfg: VisionCfg, max_batch: int = 256, device: int = 0) -> "McmEngine":
        e = cls(cfg, max_batch, device)
        e.load_state_dict(sd)
        return e

    def set_text_bank(self, bank, already_unit: bool = False) -> None:
        """Install the pre-encoded prompt bank ``[K, P]`` (rows are L2-normalised on the device like
        ``utils/detection_util.py:231`` unless ``already_unit``)."""
        t = torch.as_tensor(bank).detach().to(torch.float32).contiguous()
        if t.dim() != 2 or t.shape[1] != self.cfg.proj:
            raise ValueError(f"text bank must be [K, {self.cfg.proj}], got {tuple(t.shape)}")
        self._check(self._lib.mcm_set_text_bank(self._h, _ptr(t), t.shape[0], 1 if already_unit else 0))
        self.K = int(t.shape[0])
        self.__dict__.pop("_mcm_bank_key", None)      # detection_util's per-label-set cache no longer describes the bank

    # -------------------------------------------------------------------- compute --
    def _check_images(self, images: torch.Tensor) -> int:
        c = self.cfg
        if not torch.is_tensor(images) or images.dim() != 4 or images.shape[1] != 3:
            raise ValueError("pixel_values must be a [b, 3, H, W] tensor")
        if images.shape[2] != c.image_size or images.shape[3] != c.image_size:
            # same condition and wording as HF CLIPVisionEmbeddings.forward (HF:204-207)
            raise ValueError(f"Input image size ({images.shape[2]}*{images.shape[3]}) doesn't match model "
                             f"({c.image_size}*{c.image_size}).")
        if images.device != self.device:
            raise ValueError(f"pixel_values must live on {self.device}, got {images.device}")
        if images.dtype != torch.float32:
            raise ValueError(f"pixel_values must be float32 (the reference preprocess yields fp32), got {images.dtype}")
        if images.shape[0] > self.max_batch:
            raise ValueError(f"batch {images.shape[0]} exceeds max_batch {self.max_batch}")
        return int(images.shape[0])

    def _check_images_u8(self, images: torch.Tensor, device) -> int:
        c = self.cfg
        if not torch.is_tensor(images) or images.dim() != 4 or images.shape[3] != 3:
            raise ValueError("uint8 images must be a [b, H, W, 3] tensor (HWC, as decoded)")
        if images.shape[1] != c.image_size or images.shape[2] != c.image_size:
            raise ValueError(f"Input image size ({images.shape[1]}*{images.shape[2]}) doesn't match model "
                             f"({c.image_size}*{c.image_size}).")
        if images.dtype != torch.uint8:
            raise ValueError(f"uint8 ingest needs a torch.uint8 tensor, got {images.dtype}")
        if device is not None and images.device != device:
            raise ValueError(f"images must live on {device}, got {images.device}")
        return int(images.shape[0])

    def set_normalization(self, mean, std) -> None:
        """Constants of the fused ``ToTensor -> Normalize`` of the uint8 ingest (default: the CLIP constants of
        ``utils/train_eval_util.py:27-28``)."""
        m = (C.c_float * 3)(*[float(v) for v in mean])
        s = (C.c_float * 3)(*[float(v) for v in std])
        self._check(self._lib.mcm_set_normalization(self._h, m, s))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """``net.get_image_features(pixel_values=...)`` -> un-normalised ``[b, P]`` fp32 features."""
        b = self._check_images(pixel_values)
        x = pixel_values.contiguous()
        out = torch.empty((b, self.cfg.proj), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_image_features(self._h, _ptr(x), b, _ptr(out), self._stream()))
        return out

    def score(self, images: torch.Tensor, T: float = 1.0, score: str = "MCM",
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One batch of ``utils/detection_util.py:225-248``: device tensor ``[b]`` of scores
        (asynchronous on the current stream)."""
        b = self._check_images(images)
        x = images.contiguous()
        if out is None:
            out = torch.empty((b,), dtype=torch.float32, device=self.device)
        elif out.numel() < b or out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 device tensor with at least b elements")
        self._check(self._lib.mcm_score(self._h, _ptr(x), b, float(T), _score_kind(score), _ptr(out), self._stream()))
        return out[:b]

    def image_features_u8(self, images_u8: torch.Tensor) -> torch.Tensor:
        """``get_image_features`` from decoded uint8 ``[b, H, W, 3]`` pixels (ToTensor + Normalize fused on the device)."""
        b = self._check_images_u8(images_u8, self.device)
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        out = torch.empty((b, self.cfg.proj), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_image_features_u8(self._h, _ptr(images_u8.contiguous()), b, _ptr(out), self._stream()))
        return out

    def score_u8(self, images_u8: torch.Tensor, T: float = 1.0, score: str = "MCM",
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:meth:`score` from decoded uint8 ``[b, H, W, 3]`` device pixels."""
        b = self._check_images_u8(images_u8, self.device)
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        if out is None:
            out = torch.empty((b,), dtype=torch.float32, device=self.device)
        elif out.numel() < b or out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 device tensor with at least b elements")
        self._check(self._lib.mcm_score_u8(self._h, _ptr(images_u8.contiguous()), b, float(T), _score_kind(score), _ptr(out),
                                           self._stream()))
        return out[:b]

    def score_stream_host_u8(self, images_host_u8, batch: Optional[int] = None, T: float = 1.0,
                             score: str = "MCM") -> np.ndarray:
        """:meth:`score_stream_host` from uint8 ``[n, H, W, 3]`` HOST pixels: 4x fewer PCIe bytes."""
        t = torch.as_tensor(images_host_u8)
        if t.device.type != "cpu":
            raise ValueError("images_host_u8 must be a CPU tensor / ndarray")
        n = self._check_images_u8(t, None)
        t = t.contiguous()
        batch = int(batch or self.max_batch)
        out = np.empty((n,), dtype=np.float32)
        if n:
            self._check(self._lib.mcm_score_stream_host_u8(self._h, _ptr(t), n, batch, float(T), _score_kind(score),
                                                           C.c_void_p(out.ctypes.data)))
        return out

    def resize_crop_u8(self, images) -> torch.Tensor:
        """``CenterCrop(S)(Resize(S)(img))`` of the reference preprocess (``utils/train_eval_util.py:29-31``) for a list
        of decoded RGB images of arbitrary sizes (uint8 ``[h, w, 3]`` ndarrays / CPU tensors), on the device:
        returns a uint8 ``[n, S, S, 3]`` device tensor, bit-identical to torchvision on PIL images, ready for
        :meth:`score_u8`.  The images are packed into one pinned buffer and cross PCIe once, undecimated."""
        arrs = [np.ascontiguousarray(np.asarray(im)) for im in images]
        for a in arrs:
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("resize_crop_u8 needs uint8 [h, w, 3] (RGB, HWC) images")
        n = len(arrs)
        S = self.cfg.image_size
        out = torch.empty((n, S, S, 3), dtype=torch.uint8, device=self.device)
        if n == 0:
            return out
        sizes = np.array([a.size for a in arrs], dtype=np.int64)
        offsets = np.zeros(n, dtype=np.int64)
        offsets[1:] = np.cumsum((sizes[:-1] + 15) // 16 * 16)          # 16-byte aligned starts
        total = int(offsets[-1] + sizes[-1])
        packed = torch.empty((total,), dtype=torch.uint8).pin_memory()
        pk = packed.numpy()
        for a, o in zip(arrs, offsets):
            pk[o:o + a.size] = a.reshape(-1)
        src = packed.to(self.device, non_blocking=True)
        hs = np.array([a.shape[0] for a in arrs], dtype=np.int32)
        ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
        self.resize_crop_u8_packed(src, offsets, hs, ws, out=out)
        src.record_stream(torch.cuda.current_stream(self.device))
        return out

    def resize_crop_u8_packed(self, src: torch.Tensor, offsets, hs, ws, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:meth:`resize_crop_u8` for images already packed in one uint8 DEVICE buffer: image ``i`` is
        ``[hs[i], ws[i], 3]`` at byte ``offsets[i]`` of ``src`` (``offsets`` / ``hs`` / ``ws``: host integer sequences)."""
        if not torch.is_tensor(src) or src.dtype != torch.uint8 or src.device != self.device or not src.is_contiguous():
            raise ValueError(f"src must be a contiguous uint8 tensor on {self.device}")
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        hs = np.ascontiguousarray(hs, dtype=np.int32)
        ws = np.ascontiguousarray(ws, dtype=np.int32)
        n = int(offsets.shape[0])
        if hs.shape != (n,) or ws.shape != (n,):
            raise ValueError("offsets, hs and ws must have one entry per image")
        if n and (offsets.min() < 0 or int((offsets + hs.astype(np.int64) * ws * 3).max()) > src.numel()):
            raise ValueError("an image lies outside the packed source buffer")
        S = self.cfg.image_size
        if out is None:
            out = torch.empty((n, S, S, 3), dtype=torch.uint8, device=self.device)
        elif out.dtype != torch.uint8 or out.device != self.device or not out.is_contiguous() or out.numel() < n * S * S * 3:
            raise ValueError("out must be a contiguous uint8 device tensor of at least n * S * S * 3 elements")
        if n:
            self._check(self._lib.mcm_resize_crop_u8(
                self._h, _ptr(src), offsets.ctypes.data_as(C.POINTER(C.c_int64)), hs.ctypes.data_as(C.POINTER(C.c_int32)),
                ws.ctypes.data_as(C.POINTER(C.c_int32)), n, _ptr(out), self._stream()))
        return out

    def score_stream_host_images(self, packed, offsets, hs, ws, batch: Optional[int] = None, T: float = 1.0,
                                 score: str = "MCM") -> np.ndarray:
        """The whole evaluation stream from decoded images packed in one HOST uint8 buffer (image ``i`` is
        ``[hs[i], ws[i], 3]`` at byte ``offsets[i]``): pipelined H2D, device preprocess, scoring; float32 numpy ``[n]``."""
        t = torch.as_tensor(packed)
        if t.device.type != "cpu" or t.dtype != torch.uint8 or t.dim() != 1 or not t.is_contiguous():
            raise ValueError("packed must be a contiguous 1-D uint8 CPU tensor / ndarray")
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        hs = np.ascontiguousarray(hs, dtype=np.int32)
        ws = np.ascontiguousarray(ws, dtype=np.int32)
        n = int(offsets.shape[0])
        if hs.shape != (n,) or ws.shape != (n,):
            raise ValueError("offsets, hs and ws must have one entry per image")
        if n and (offsets.min() < 0 or int((offsets + hs.astype(np.int64) * ws * 3).max()) > t.numel()):
            raise ValueError("an image lies outside the packed buffer")
        out = np.empty((n,), dtype=np.float32)
        if n:
            self._check(self._lib.mcm_score_stream_host_images(
                self._h, _ptr(t), offsets.ctypes.data_as(C.POINTER(C.c_int64)), hs.ctypes.data_as(C.POINTER(C.c_int32)),
                ws.ctypes.data_as(C.POINTER(C.c_int32)), n, int(batch or self.max_batch), float(T), _score_kind(score),
                C.c_void_p(out.ctypes.data)))
        return out

    def score_images(self, images, T: float = 1.0, score: str = "MCM") -> torch.Tensor:
        """Decoded RGB images of arbitrary sizes -> scores: the WHOLE reference preprocess (Resize, CenterCrop, ToTensor,
        Normalize) and the scoring path on the device.  Returns a device tensor ``[n]``."""
        images = list(images)
        parts = []
        for s0 in range(0, len(images), self.max_batch):
            parts.append(self.score_u8(self.resize_crop_u8(images[s0:s0 + self.max_batch]), T=T, score=score))
        if not parts:
            return torch.empty((0,), dtype=torch.float32, device=self.device)
        return torch.cat(parts)

    def score_stream_host(self, images_host, batch: Optional[int] = None, T: float = 1.0,
                          score: str = "MCM") -> np.ndarray:
        """Whole evaluation stream from HOST memory (the loop of ``utils/detection_util.py:220-249``):
        H2D copies overlap the scoring of the previous batch; returns float32 numpy ``[n]``."""
        t = torch.as_tensor(images_host)
        if t.device.type != "cpu" or t.dtype != torch.float32:
            raise ValueError("images_host must be a float32 CPU tensor / ndarray")
        t = t.contiguous()
        c = self.cfg
        if t.dim() != 4 or tuple(t.shape[1:]) != (3, c.image_size, c.image_size):
            raise ValueError(f"images_host must be [n, 3, {c.image_size}, {c.image_size}], got {tuple(t.shape)}")
        n = int(t.shape[0])
        batch = int(batch or self.max_batch)
        out = np.empty((n,), dtype=np.float32)
        if n:
            self._check(self._lib.mcm_score_stream_host(self._h, _ptr(t), n, batch, float(T), _score_kind(score),
                                                        C.c_void_p(out.ctypes.data)))
        return out

    # ------------------------------------------------------------ Mahalanobis baseline --
    def set_maha(self, classwise_mean, precision, normalize: bool = False) -> None:
        """Install the statistics of the reference's ``--score maha`` (``get_mean_prec``,
        ``utils/detection_util.py:148-180``): class means ``[K, P]`` and the shared precision matrix ``[P, P]``.
        The quadratic form only sees the symmetric part of ``precision``; it is factored on the host in fp64
        (``P = L L^T``) so that the device computes ``0.5 * min_k |f L - mu_k L|^2``."""
        mean = torch.as_tensor(classwise_mean).detach().cpu().double()
        prec = torch.as_tensor(precision).detach().cpu().double()
        P = self.cfg.proj
        if mean.dim() != 2 or mean.shape[1] != P or tuple(prec.shape) != (P, P):
            raise ValueError(f"classwise_mean must be [K, {P}] and precision [{P}, {P}]")
        sym = 0.5 * (prec + prec.T)
        L, info = torch.linalg.cholesky_ex(sym)
        if int(info) != 0:
            raise ValueError("precision matrix is not positive definite (Cholesky failed); Mahalanobis distances are undefined")
        lt = L.T.contiguous().float()
        centres = (mean @ L).contiguous().float()
        self._check(self._lib.mcm_set_maha(self._h, _ptr(lt), _ptr(centres), int(mean.shape[0]), 1 if normalize else 0))
        self.maha_K = int(mean.shape[0])

    def maha_score(self, images: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One batch of ``get_Mahalanobis_score`` (``utils/detection_util.py:193-204``): device tensor ``[b]``."""
        b = self._check_images(images)
        if out is None:
            out = torch.empty((b,), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_maha_score(self._h, _ptr(images.contiguous()), b, _ptr(out), self._stream()))
        return out[:b]

    def dbg_maha_from_features(self, feats: torch.Tensor) -> torch.Tensor:
        b = int(feats.shape[0])
        out = torch.empty((b,), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_dbg_maha_from_features(self._h, _ptr(feats.contiguous().float()), b, _ptr(out), self._stream()))
        return out

    # ------------------------------------------------- per-kernel entry points (tests, bench) --
    def dbg_gemm(self, a, w, bias, resid=None, epi: int = 0):
        """epilogue(A[M,K] @ W[N,K]^T) through the tcgen05 GEMM; a, w fp16.  epi 0/1 -> fp16, 2 -> fp32."""
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty((M, N), dtype=torch.float32 if epi == 2 else torch.float16, device=self.device)
        self._check(self._lib.mcm_dbg_gemm(self._h, _ptr(a.contiguous()), _ptr(w.contiguous()), _ptr(bias),
                                           _ptr(resid), _ptr(out), M, N, K, int(epi), self._stream()))
        return out

    def dbg_fold_ln(self, w, gamma, beta, bias):
        """LayerNorm fold of one projection: (w16 = fp16(gamma o W), c, d); see csrc/gemm_tcgen05.cuh."""
        N, K = w.shape
        w16 = torch.empty((N, K), dtype=torch.float16, device=self.device)
        c = torch.empty((N,), dtype=torch.float32, device=self.device)
        d = torch.empty((N,), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_dbg_fold_ln(self._h, _ptr(w.contiguous()), _ptr(gamma), _ptr(beta), _ptr(bias), _ptr(w16),
                                              _ptr(c), _ptr(d), N, K, self._stream()))
        return w16, c, d

    def dbg_gemm_resid_ln(self, a, w, bias, resid, in_place: bool = False, mutate: bool = False):
        """out = resid + A @ W^T + bias (fp32), its fp16 copy and the partial row statistics [parts, M, 2].
        ``in_place`` aliases resid and out like the forward does (K <= 1024 then takes the TMA epilogue)."""
        M, K = a.shape
        N = w.shape[0]
        if mutate:
            out = resid          # timing runs: update the caller's tensor itself
        else:
            out = resid.clone().contiguous() if in_place else torch.empty((M, N), dtype=torch.float32, device=self.device)
        if in_place:
            resid = out
        out16 = torch.empty((M, N), dtype=torch.float16, device=self.device)
        stats = torch.zeros((N // 64, M, 2), dtype=torch.float32, device=self.device)
        parts = C.c_int32(0)
        self._check(self._lib.mcm_dbg_gemm_resid_ln(self._h, _ptr(a.contiguous()), _ptr(w.contiguous()), _ptr(bias),
                                                    _ptr(resid.contiguous()), _ptr(out), _ptr(out16), _ptr(stats), M, N, K,
                                                    C.byref(parts), self._stream()))
        return out, out16, stats[:parts.value]

    def dbg_gemm_resid_h2(self, a, w, bias, x_hi, x_lo):
        """(x_hi, x_lo) <- split(x_hi + x_lo + A @ W^T + bias) in place (the residual epilogue of the forward);
        returns the partial row statistics [parts, M, 2]."""
        M, K = a.shape
        N = w.shape[0]
        stats = torch.zeros((N // 64, M, 2), dtype=torch.float32, device=self.device)
        parts = C.c_int32(0)
        self._check(self._lib.mcm_dbg_gemm_resid_h2(self._h, _ptr(a.contiguous()), _ptr(w.contiguous()), _ptr(bias), _ptr(x_hi),
                                                    _ptr(x_lo), _ptr(stats), M, N, K, C.byref(parts), self._stream()))
        return stats[:parts.value]

    def dbg_gemm_ln(self, a, w16, d, c, stats, row_len: int, gelu: bool = False):
        """[quick_gelu](LayerNorm(rows) @ W^T + b) through the folded projection; stats: [parts, M, 2] fp32."""
        M, K = a.shape
        N = w16.shape[0]
        out = torch.empty((M, N), dtype=torch.float16, device=self.device)
        stats = stats.contiguous()
        self._check(self._lib.mcm_dbg_gemm_ln(self._h, _ptr(a.contiguous()), _ptr(w16.contiguous()), _ptr(d), _ptr(c), _ptr(stats),
                                              stats.shape[0], int(row_len), _ptr(out), M, N, K, 1 if gelu else 0, self._stream()))
        return out

    def dbg_layernorm(self, x, gamma, beta, eps: float = 1e-5, out_f16: bool = True):
        M, D = x.shape
        out = torch.empty((M, D), dtype=torch.float16 if out_f16 else torch.float32, device=self.device)
        self._check(self._lib.mcm_dbg_layernorm(self._h, _ptr(x.contiguous()), _ptr(gamma), _ptr(beta), _ptr(out), M, D,
                                                float(eps), 1 if out_f16 else 0, self._stream()))
        return out

    def dbg_attention(self, qkv, b: int, S: int, H: int):
        out = torch.empty((b * S, H * 64), dtype=torch.float16, device=self.device)
        self._check(self._lib.mcm_dbg_attention(self._h, _ptr(qkv.contiguous()), _ptr(out), b, S, H, self._stream()))
        return out

    def dbg_tail(self, x, b: int, T: float = 1.0, score: str = "MCM", want_feats=True, want_scores=True):
        feats = torch.empty((b, self.cfg.proj), dtype=torch.float32, device=self.device) if want_feats else None
        scores = torch.empty((b,), dtype=torch.float32, device=self.device) if want_scores else None
        self._check(self._lib.mcm_dbg_tail(self._h, _ptr(x.contiguous()), b, float(T), _score_kind(score), _ptr(feats),
                                           _ptr(scores), self._stream()))
        return feats, scores

    def dbg_embed(self, images):
        b = self._check_images(images)
        x = torch.empty((b * self.cfg.seq, self.cfg.width), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_dbg_embed(self._h, _ptr(images.contiguous()), b, _ptr(x), self._stream()))
        return x

    # ---------------------------------------------------------------- bookkeeping --
    @property
    def launch_count(self) -> int:
        return int(self._lib.mcm_launch_count(self._h))

    def reset_launch_count(self) -> None:
        self._lib.mcm_reset_launch_count(self._h)

    def set_precision(self, mode) -> None:
        """``"fp16"`` / 0 (default): one fp16 value per tensor-core operand element.  ``"split"`` / 1: fp16 (hi, lo)
        operand pairs and three-term products -- fp32-class results (AUROC / FPR95 identical to the fp32 reference to
        the parity bar) at 3x the tensor work.  See ``MCM_OPT_PRECISION`` in include/mcm_b200.h."""
        if mode not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of 'fp16', 'split' (got {mode!r})")
        self._check(self._lib.mcm_set_option(self._h, _lib.OPT_PRECISION, _lib.PRECISIONS[mode]))
        self.precision = "split" if _lib.PRECISIONS[mode] == _lib.PRECISION_SPLIT else "fp16"

    def set_cuda_graph(self, on: bool) -> None:
        """Replay the forward from a CUDA graph captured per (input buffer, batch size, options): small batches."""
        self._check(self._lib.mcm_set_option(self._h, _lib.OPT_CUDA_GRAPH, 1 if on else 0))

    def allgather_scores(self, nccl_comm: int, local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``mcm_allgather_scores``: one NCCL all-gather of this rank's (padded) score vector on the current stream;
        ``nccl_comm`` is a raw ``ncclComm_t`` (see :class:`mcm_b200.parallel.NcclComm`)."""
        if local.dtype != torch.float32 or out.dtype != torch.float32 or not local.is_contiguous() or not out.is_contiguous():
            raise ValueError("local / out must be contiguous float32 device tensors")
        self._check(self._lib.mcm_allgather_scores(self._h, C.c_void_p(int(nccl_comm)), _ptr(local), int(local.numel()), _ptr(out),
                                                   self._stream()))
        return out

    def dbg_gemm_split(self, a_hi, a_lo, w_hi, w_lo, bias, resid):
        """resid + (a_hi + a_lo) @ (w_hi + w_lo)^T + bias through the three-term split GEMM (fp32 out)."""
        M, K = a_hi.shape
        N = w_hi.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        self._check(self._lib.mcm_dbg_gemm_split(self._h, _ptr(a_hi.contiguous()), _ptr(a_lo.contiguous()), _ptr(w_hi.contiguous()),
                                                 _ptr(w_lo.contiguous()), _ptr(bias), _ptr(resid.contiguous()), _ptr(out), M, N, K,
                                                 self._stream()))
        return out

    def dbg_attention_split(self, qkv_hi, qkv_lo, b: int, S: int, H: int):
        o_hi = torch.empty((b * S, H * 64), dtype=torch.float16, device=self.device)
        o_lo = torch.empty_like(o_hi)
        self._check(self._lib.mcm_dbg_attention_split(self._h, _ptr(qkv_hi.contiguous()), _ptr(qkv_lo.contiguous()), _ptr(o_hi),
                                                      _ptr(o_lo), b, S, H, self._stream()))
        return o_hi, o_lo

    def set_cls_shortcut(self, on: bool) -> None:
        """Last-layer CLS-only shortcut (default on; identical results, ~6 % fewer executed FLOPs)."""
        self._check(self._lib.mcm_set_option(self._h, _lib.OPT_CLS_SHORTCUT, 1 if on else 0))

    def profile(self, on: bool) -> None:
        """Bracket every launch with CUDA events (bench.py's per-kernel roofline)."""
        self._check(self._lib.mcm_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        """{kind: (total_ms, launches)} accumulated since the last reset (synchronises)."""
        n = len(_lib.PROF_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._check(self._lib.mcm_profile_read(self._h, ms, cnt, 1 if reset else 0))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.PROF_KINDS)}

    def flops_per_image(self, K: Optional[int] = None) -> float:
        return float(self._lib.mcm_flops_per_image(C.byref(self._ccfg), int(self.K if K is None else K)))


class B200ClipNet:
    """Duck-typed stand-in for the ``net`` the reference threads through its functions
    (``eval_ood_detection.py:60-61``): image side on the B200 engine, text side (run once per label
    set, not a kernel target -- SURVEY.md section 2.1 row 4) delegated to ``text_model`` (any object
    with HF's ``get_text_features``) or replaced by a pre-encoded ``text_bank``."""

    def __init__(self, engine: McmEngine, text_model=None, text_bank=None):
        self.engine = engine
        self.text_model = text_model
        self.text_bank = None if text_bank is None else torch.as_tensor(text_bank, dtype=torch.float32)

    def eval(self):
        if self.text_model is not None and hasattr(self.text_model, "eval"):
            self.text_model.eval()
        return self

    def cuda(self, *a, **k):
        return self

    def get_image_features(self, pixel_values=None, **kw):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        return self.engine.image_features(pixel_values)

    @torch.no_grad()
    def get_text_features(self, input_ids=None, attention_mask=None, **kw):
        if self.text_bank is not None:
            return self.text_bank.clone()
        if self.text_model is None:
            raise RuntimeError("B200ClipNet has neither a text_model nor a pre-encoded text_bank")
        dev = next(self.text_model.parameters()).device
        out = self.text_model.get_text_features(input_ids=input_ids.to(dev),
                                                attention_mask=None if attention_mask is None else attention_mask.to(dev),
                                                **kw)
        # transformers >= 5 returns BaseModelOutputWithPooling, 4.x the projected tensor
        return out.pooler_output if hasattr(out, "pooler_output") else out


def engine_for_ckpt(ckpt: str, state_dict, max_batch: int = 256, device: int = 0) -> McmEngine:
    """``--CLIP_ckpt`` / ``args.ckpt`` -> engine (mapping of ``utils/train_eval_util.py:19-22``)."""
    names = {"ViT-B/32": "ViT-B/32", "ViT-B/16": "ViT-B/16", "ViT-L/14": "ViT-L/14", **CKPT_TO_CFG}
    if ckpt not in names:
        raise ValueError(f"unknown CLIP checkpoint {ckpt!r}")
    return McmEngine.from_state_dict(state_dict, CFGS[names[ckpt]], max_batch=max_batch, device=device)
