"""Drop-in for the scoring API of the reference's ``utils/detection_util.py`` on the B200 engine.

``get_ood_scores_clip(args, net, loader, test_labels, in_dist=False)`` keeps the reference
signature, argument meaning and return contract (``utils/detection_util.py:209-249``: float32
``np.ndarray`` of shape ``(len(loader.dataset),)`` holding ``-max softmax(cos/T)`` for
``args.score == 'MCM'``; the other ``--score`` reductions follow ``:233-248``), so
``eval_ood_detection.py:81,90`` can import it from here unchanged.

What differs from the reference is only *where* the work happens:

* the prompt bank is encoded ONCE per ``(args.ckpt, test_labels)`` instead of once per image batch
  (``:228-231``) -- identical values, the text tower is deterministic under ``eval()`` + ``no_grad``;
* image batches are scored entirely on the device by the sm_100a engine and only the ``[b]`` scores
  ever cross PCIe (the reference ships the ``[b, K]`` softmax, ``:236``), with one synchronisation
  at the end of the stream instead of one per batch.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
import torch

from . import metrics
from .engine import B200ClipNet, McmEngine
from .metrics import fpr_and_fdr_at_recall, get_measures, stable_cumsum  # noqa: F401  (re-exported API)

try:  # the reference builds its tokenizer from this module-level name (utils/detection_util.py:8,216)
    from transformers import CLIPTokenizer
except Exception:  # pragma: no cover - transformers is optional when the bank is pre-encoded
    CLIPTokenizer = None

to_np = lambda x: x.data.cpu().numpy()  # noqa: E731  (same helper name as the reference, :12)

_BANK_CACHE_ATTR = "_mcm_bank_key"


def prompt_texts(test_labels: Iterable[str]):
    """The single prompt template the reference uses (no trailing period), ``:228``."""
    return [f"a photo of a {c}" for c in test_labels]


def encode_text_bank(args, net, test_labels) -> torch.Tensor:
    """``[K, P]`` un-normalised text features of the prompts, exactly the tensors ``:228-230`` build."""
    if getattr(net, "text_bank", None) is not None:
        return net.get_text_features()
    if CLIPTokenizer is None:
        raise RuntimeError("transformers is not importable and the net carries no pre-encoded text_bank")
    tokenizer = CLIPTokenizer.from_pretrained(args.ckpt)
    text_inputs = tokenizer(prompt_texts(test_labels), padding=True, return_tensors="pt")
    with torch.no_grad():
        feats = net.get_text_features(input_ids=text_inputs["input_ids"], attention_mask=text_inputs["attention_mask"])
    return feats.float()


def _ensure_bank(args, net: B200ClipNet, test_labels) -> None:
    labels = [str(c) for c in test_labels]
    key = (getattr(args, "ckpt", None), tuple(labels), id(getattr(net, "text_model", None)),
           None if getattr(net, "text_bank", None) is None else net.text_bank.data_ptr())
    eng = net.engine
    if getattr(eng, _BANK_CACHE_ATTR, None) == key:
        return
    bank = encode_text_bank(args, net, labels)
    if bank.shape[0] != len(labels):
        raise ValueError(f"text bank has {bank.shape[0]} rows for {len(labels)} labels")
    eng.set_text_bank(bank, already_unit=False)   # rows normalised on the device, :231
    setattr(eng, _BANK_CACHE_ATTR, key)


def get_ood_scores_clip(args, net, loader, test_labels, in_dist=False):
    """MCM (or ``args.score``) scores of every image of ``loader`` -- reference signature, ``:209``.

    ``net`` must be a :class:`mcm_b200.engine.B200ClipNet`; there is no CPU / PyTorch fallback.
    ``in_dist`` is accepted and unused, as in the reference.
    """
    if not isinstance(net, B200ClipNet):
        raise TypeError("mcm_b200.get_ood_scores_clip needs a B200ClipNet (see mcm_b200.set_model_clip); "
                        "running an arbitrary torch module here would be a silent fallback")
    if getattr(args, "model", "CLIP") != "CLIP":
        raise ValueError(f"model {args.model!r} is not supported (the reference only scores 'CLIP', :227)")
    score = getattr(args, "score", "MCM")
    eng: McmEngine = net.engine
    _ensure_bank(args, net, test_labels)
    T = float(args.T)
    n_total = len(loader.dataset)
    parts = []
    for batch in loader:
        images = batch[0] if isinstance(batch, (tuple, list)) else batch
        images = images.to(eng.device, non_blocking=True)
        for s in range(0, images.shape[0], eng.max_batch):   # a loader may use a larger batch than the engine
            parts.append(eng.score(images[s:s + eng.max_batch], T=T, score=score))
    if not parts:
        return np.zeros((0,), dtype=np.float32)
    # one device->host copy + synchronisation for the whole stream; trim like [:len(loader.dataset)], :249
    return torch.cat(parts).cpu().numpy().astype(np.float32, copy=False)[:n_total].copy()


def get_mean_prec(args, net, train_loader):
    """Class-wise means and the shared precision matrix for the Mahalanobis score -- reference signature and results
    (``utils/detection_util.py:148-180``): features come from the B200 engine, the statistics are the reference's own
    torch code (fp64 covariance, ``linalg.inv``), INCLUDING that ``classwise_idx`` collects the batch index once per
    sample (``:166-167``, sic) -- the files it writes under ``args.template_dir`` are interchangeable with the reference's."""
    import os
    from collections import defaultdict
    if not isinstance(net, B200ClipNet):
        raise TypeError("mcm_b200.get_mean_prec needs a B200ClipNet")
    eng: McmEngine = net.engine
    classwise_mean = torch.empty(args.n_cls, args.feat_dim)
    all_features = []
    classwise_idx = defaultdict(list)
    with torch.no_grad():
        for idx, (images, labels) in enumerate(train_loader):
            images = images.to(eng.device, non_blocking=True)
            features = torch.cat([eng.image_features(images[s:s + eng.max_batch]) for s in range(0, images.shape[0], eng.max_batch)]).float()
            if args.normalize:
                features /= features.norm(dim=-1, keepdim=True)
            for label in labels:
                classwise_idx[label.item()].append(idx)
            all_features.append(features.cpu())
    all_features = torch.cat(all_features)
    for cls in range(args.n_cls):
        classwise_mean[cls] = torch.mean(all_features[classwise_idx[cls]].float(), dim=0)
        if args.normalize:
            classwise_mean[cls] /= classwise_mean[cls].norm(dim=-1, keepdim=True)
    cov = torch.cov(all_features.T.double())
    precision = torch.linalg.inv(cov).float()
    print(f"cond number: {torch.linalg.cond(precision)}")
    tdir = getattr(args, "template_dir", None)
    if tdir:
        os.makedirs(tdir, exist_ok=True)
        torch.save(classwise_mean, os.path.join(tdir, f"{args.model}_classwise_mean_{args.in_dataset}_{args.max_count}_{args.normalize}.pt"))
        torch.save(precision, os.path.join(tdir, f"{args.model}_precision_{args.in_dataset}_{args.max_count}_{args.normalize}.pt"))
    return classwise_mean, precision


def get_Mahalanobis_score(args, net, test_loader, classwise_mean, precision, in_dist=True):
    """Mahalanobis confidence score of every image of ``test_loader`` -- reference signature and return contract
    (``utils/detection_util.py:182-207``), including its batch rule: with ``in_dist=False`` the loop stops at batch
    ``len(dataset) // args.batch_size`` (``:191-192``), so a trailing partial batch of an OOD set is not scored."""
    if not isinstance(net, B200ClipNet):
        raise TypeError("mcm_b200.get_Mahalanobis_score needs a B200ClipNet")
    eng: McmEngine = net.engine
    key = (id(classwise_mean), id(precision), bool(args.normalize))
    if getattr(eng, "_mcm_maha_key", None) != key:
        eng.set_maha(classwise_mean, precision, bool(args.normalize))
        eng._mcm_maha_key = key
        eng._mcm_maha_refs = (classwise_mean, precision)      # keep the ids alive
    total_len = len(test_loader.dataset)
    parts = []
    for batch_idx, batch in enumerate(test_loader):
        if (batch_idx >= total_len // args.batch_size) and in_dist is False:
            break
        images = (batch[0] if isinstance(batch, (tuple, list)) else batch).to(eng.device, non_blocking=True)
        for s in range(0, images.shape[0], eng.max_batch):
            parts.append(eng.maha_score(images[s:s + eng.max_batch]))
    if not parts:
        return np.zeros((0,), dtype=np.float32)
    return torch.cat(parts).cpu().numpy().astype(np.float32, copy=True)


def print_measures(log, auroc, aupr, fpr, method_name="Ours", recall_level=0.95):
    """Same report lines as the reference (``:37-45``)."""
    r = int(100 * recall_level)
    if log is None:
        print(f"FPR{r:d}:\t\t\t{100 * fpr:.2f}")
        print(f"AUROC: \t\t\t{100 * auroc:.2f}")
        print(f"AUPR:  \t\t\t{100 * aupr:.2f}")
    else:
        log.debug("\t\t\t\t" + method_name)
        log.debug(f"  FPR{r:d} AUROC AUPR")
        log.debug(f"& {100 * fpr:.2f} & {100 * auroc:.2f} & {100 * aupr:.2f}")


def get_and_print_results(args, log, in_score, out_score, auroc_list, aupr_list, fpr_list):
    """Metrics of one OOD set, appended to the running lists (``:253-265``)."""
    auroc, aupr, fpr = metrics.get_measures(-np.asarray(in_score), -np.asarray(out_score))
    print(f"in score samples (random sampled): {in_score[:3]}, out score samples: {out_score[:3]}")
    auroc_list.append(auroc)
    aupr_list.append(aupr)
    fpr_list.append(fpr)
    print_measures(log, auroc, aupr, fpr, args.score)
