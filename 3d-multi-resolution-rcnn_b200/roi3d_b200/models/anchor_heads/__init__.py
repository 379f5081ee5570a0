from .rpn_head_3d import RPNProposal3D, decode_proposals, topk_segmented

__all__ = ['RPNProposal3D', 'decode_proposals', 'topk_segmented']
