"""openglue_b200: the OpenGlue matching core (SuperGlue-style GNN + Sinkhorn) on B200.

Public surface mirrors the reference for this one path:
    SuperGlue(config).forward(data)  -> {'context_descriptors0', 'context_descriptors1', 'scores'}
    MatchingCore(superglue)(data)    -> {'matches0', 'matching_scores0', 'matches1', 'matching_scores1'}
and, for the step right before it in the reference's training / validation step:
    generate_gt_matches(data, features0, features1, positive_threshold, negative_threshold) -> (data, y_true)
and the steps either side of those:
    collate_features(batch, target_num_keypoints, random)     (the cached-feature DataLoader's collate_fn, on the GPU)
    criterion(y_true, y_pred, margin=None) -> {'loss', 'metric_loss'}
    matching_log_probs(S, dustbin_score, num_iters, reg)      (differentiable Sinkhorn: forward + backward kernels)
    SuperGlue(config).train()(data)                           (training mode: batch-statistics BatchNorm, explicit backward pass)
    SuperPointNet(max_keypoints, ...)(image) -> (lafs, scores, descriptors)   (the detector / descriptor front-end; SuperPointNetBn: its BatchNorm variant)
"""
from .gt_matches import generate_gt_matches  # noqa: F401
from .feature_cache import FeatureStore, collate_features  # noqa: F401
from .losses import criterion  # noqa: F401
from .sinkhorn import matching_log_probs  # noqa: F401
from .superglue import MatchingCore, PendingMatches, SuperGlue  # noqa: F401
from .superpoint import SuperPointNet, SuperPointNetBn  # noqa: F401

__version__ = '0.1.0'
