"""Hungarian matcher — mirror of the reference's models/matcher.py (HungarianMatcher :66-201).

The cost matrix C[q, t] = cost_mask * BCE + cost_class * (-p[q, label_t]) + cost_dice * dice of one scene is ONE
fused pass over the [S, Q] mask logits (libus3d us3d_matcher_cost) instead of the reference's chain of
softplus / einsum kernels (:35-59, 12-27, 107-160); the assignment itself stays
scipy.optimize.linear_sum_assignment on the host, exactly as the reference (:161-163).
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, cost_noise_robust: float = 1.0,
                 num_points: int = 0):
        super().__init__()
        self.cost_class, self.cost_mask, self.cost_dice, self.cost_noise_robust = cost_class, cost_mask, cost_dice, cost_noise_robust
        if self.cost_class == 0 and self.cost_mask == 0 and self.cost_dice == 0:
            self.cost_mask = 1
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points

    @torch.no_grad()
    def cost_matrices(self, outputs, targets, mask_type):
        """Per-scene cost matrices [Q, T] on the device (no host sync)."""
        from unscene3d_b200.engine import functional as Fn  # CUDA only: there is no CPU path

        bs = outputs["pred_logits"].shape[0]
        costs = []
        for b in range(bs):
            prob = outputs["pred_logits"][b].float().softmax(-1)
            labels = targets[b]["labels"]
            logits_sq = outputs["pred_masks"][b]          # [S, Q]
            tgt_ts = targets[b][mask_type].to(logits_sq)  # [T, S]
            if self.num_points != -1:  # sub-sample the points shared by all masks (models/matcher.py:122-127)
                idx = torch.randperm(tgt_ts.shape[1], device=tgt_ts.device)[:int(self.num_points * tgt_ts.shape[1])]
                logits_sq, tgt_ts = logits_sq[idx], tgt_ts[:, idx]
            costs.append(Fn.matcher_cost(logits_sq, tgt_ts, prob, labels, self.cost_class, self.cost_mask, self.cost_dice))
        return costs

    @torch.no_grad()
    def memory_efficient_forward(self, outputs, targets, mask_type):
        costs = self.cost_matrices(outputs, targets, mask_type)
        indices = [linear_sum_assignment(c.cpu()) for c in costs]
        return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in indices]

    @torch.no_grad()
    def forward(self, outputs, targets, mask_type):
        return self.memory_efficient_forward(outputs, targets, mask_type)

    def __repr__(self, _repr_indent=4):
        body = [f"cost_class: {self.cost_class}", f"cost_mask: {self.cost_mask}", f"cost_dice: {self.cost_dice}"]
        return "\n".join(["Matcher " + self.__class__.__name__] + [" " * _repr_indent + line for line in body])
