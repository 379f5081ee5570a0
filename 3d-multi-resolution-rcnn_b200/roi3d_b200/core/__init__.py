"""Host mirrors of mmdet.core pieces on the hot path: anchors, boxes, post-processing, evaluation-time NMS."""
