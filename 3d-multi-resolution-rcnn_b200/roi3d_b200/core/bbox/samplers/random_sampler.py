"""RandomSampler of the RoI-stage training glue.

Reference: mmdet/core/bbox/samplers/random_sampler.py:8-58 on top of base_sampler.py:8-103.  The draw is host-side
numpy RNG in the reference (`np.random.randint(0, len(gallery), num)` -- WITH replacement, then `.unique()`), once for
the positives and once for the negatives, and it stays that way here so that a seeded run selects the same boxes:
the candidates are found on the device, only their count comes to the host (the reference reads the whole index
tensors), the drawn positions go up as one small tensor.
"""
import numpy as np
import torch

from .sampling_result import SamplingResult


class RandomSampler(object):

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, **kwargs):
        self.num, self.pos_fraction = num, pos_fraction
        self.neg_pos_ub, self.add_gt_as_proposals = neg_pos_ub, add_gt_as_proposals
        self.pos_sampler = self
        self.neg_sampler = self

    @staticmethod
    def random_choice(gallery, num):
        """random_sampler.py:19-40: `num` positions drawn with np.random.randint (duplicates possible)."""
        assert len(gallery) >= num
        if isinstance(gallery, list):
            gallery = np.array(gallery)
        rand_inds = np.random.randint(low=0, high=len(gallery), size=num)
        if not isinstance(gallery, np.ndarray):
            rand_inds = torch.from_numpy(rand_inds).long().to(gallery.device)
        return gallery[rand_inds]

    def _sample(self, flags, num_expected):
        inds = torch.nonzero(flags)
        if inds.numel() != 0:
            inds = inds.squeeze(1)
        if inds.numel() <= num_expected:
            return inds
        return self.random_choice(inds, num_expected)

    def _sample_pos(self, assign_result, num_expected, **kwargs):
        return self._sample(assign_result.gt_inds > 0, num_expected)

    def _sample_neg(self, assign_result, num_expected, **kwargs):
        return self._sample(assign_result.gt_inds == 0, num_expected)

    def sample(self, assign_result, bboxes, gt_bboxes, gt_labels=None, **kwargs):
        """base_sampler.py:31-103."""
        if isinstance(gt_bboxes, list) and len(gt_bboxes) == 1:
            gt_bboxes = gt_bboxes[0]
        if isinstance(gt_labels, list) and len(gt_labels) == 1:
            gt_labels = gt_labels[0]
        if bboxes.shape[1] >= 6:
            bboxes = bboxes[:, :6]
        elif bboxes.shape[1] >= 4:
            bboxes = bboxes[:, :4]
        gt_flags = bboxes.new_zeros((bboxes.shape[0],), dtype=torch.uint8)
        if self.add_gt_as_proposals:
            bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
            assign_result.add_gt_(gt_labels)
            gt_ones = bboxes.new_ones(gt_bboxes.shape[0], dtype=torch.uint8)
            gt_flags = torch.cat([gt_ones, gt_flags])
        num_expected_pos = int(self.num * self.pos_fraction)
        pos_inds = self.pos_sampler._sample_pos(assign_result, num_expected_pos, bboxes=bboxes, **kwargs)
        pos_inds = pos_inds.unique()
        num_sampled_pos = pos_inds.numel()
        num_expected_neg = self.num - num_sampled_pos
        if self.neg_pos_ub >= 0:
            _pos = max(1, num_sampled_pos)
            neg_upper_bound = int(self.neg_pos_ub * _pos)
            if num_expected_neg > neg_upper_bound:
                num_expected_neg = neg_upper_bound
        neg_inds = self.neg_sampler._sample_neg(assign_result, num_expected_neg, bboxes=bboxes, **kwargs)
        neg_inds = neg_inds.unique()
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)
