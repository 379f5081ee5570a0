"""Result record of the assigner (reference: mmdet/core/bbox/assigners/assign_result.py:4-25)."""
import torch


class AssignResult(object):

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None, bregions=None):
        self.num_gts, self.gt_inds, self.max_overlaps = num_gts, gt_inds, max_overlaps
        self.labels, self.bregions = labels, bregions

    def add_gt_(self, gt_labels, gt_bregions=None):
        """Prepend the gts themselves as positives of themselves (sampler option add_gt_as_proposals)."""
        own = torch.arange(1, len(gt_labels) + 1, dtype=torch.long, device=gt_labels.device)
        self.gt_inds = torch.cat([own, self.gt_inds])
        self.max_overlaps = torch.cat([self.max_overlaps.new_ones(self.num_gts), self.max_overlaps])
        if self.labels is not None:
            self.labels = torch.cat([gt_labels, self.labels])
        if self.bregions is not None:
            self.bregions = torch.cat([gt_bregions, self.bregions])
