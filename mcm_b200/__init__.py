"""mcm_b200 -- B200-native implementation of the MCM (Maximum Concept Matching) scoring hot path.

Public surface (mirrors the reference's names for this path):

* ``mcm_b200.detection_util.get_ood_scores_clip`` -- drop-in for ``utils.detection_util`` (seam 1)
* ``mcm_b200.train_eval_util.set_model_clip`` / ``wrap_clip_model`` -- the ``net`` object (seam 2)
* ``mcm_b200.engine.McmEngine`` -- the engine behind both, over the C ABI in ``include/mcm_b200.h``
* ``mcm_b200.metrics`` -- AUROC / AUPR / FPR95
* ``mcm_b200.parallel`` -- stream sharding + score all-gather

Importing the package does not need a GPU; any compute call does (there is no CPU fallback).
"""
__version__ = "0.1.0"
