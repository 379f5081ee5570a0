from .gather import gather_detections, shard_indices

__all__ = ['gather_detections', 'shard_indices']
