"""Set criterion of the self-training step — mirror of the reference's models/criterion.py
(dice_loss :22-43, sigmoid_ce_loss :51-70, SetCriterion :90-292) with the same loss keys and weights
semantics: matcher -> weighted cross-entropy over (num_classes + no-object) -> per-scene sigmoid-CE + dice on the
matched masks (optional DropLoss IoU gating) -> the same again for every auxiliary decoder output.

The tri-plane noise-robust term (models/noise_robust_loss.py over libus3d's projection kernels) is only evaluated when its
weight is non-zero (models/criterion.py:170); the self-training configuration keeps it at 0
(conf/matcher/hungarian_matcher.yaml:6).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn


def _world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _cuda_mask_losses(logits_sq, targets_ts, qidx, tidx, weights, n):
    from unscene3d_b200.engine import functional as Fn  # CUDA only: there is no CPU path

    return Fn.mask_losses(logits_sq, targets_ts, qidx, tidx, weights, n)


class SetCriterion(nn.Module):
    # (loss_mask, loss_dice) of one scene's matched pairs: the libus3d kernels (one pass over the matched columns).  Like the
    # sparse operators these model files take from `MinkowskiEngine`, the core can be swapped: the CPU tests that run the
    # definitions over the oracle install the oracle's restatement of dice_loss / sigmoid_ce_loss here (tests/helpers.py).
    mask_loss_core = None

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses, num_points, oversample_ratio,
                 importance_sample_ratio, class_weights, directions="xyz", use_droploss=False, droploss_iou_thresh=0.1):
        super().__init__()
        self.num_classes = num_classes - 1
        self.class_weights = class_weights
        self.matcher, self.weight_dict, self.eos_coef, self.losses = matcher, weight_dict, eos_coef, losses
        self.use_droploss, self.droploss_iou_thresh = use_droploss, droploss_iou_thresh
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        if self.class_weights != -1:
            assert len(self.class_weights) == self.num_classes, "CLASS WEIGHTS DO NOT MATCH"
            empty_weight[:-1] = torch.tensor(self.class_weights)
        self.register_buffer("empty_weight", empty_weight)
        self.num_points, self.oversample_ratio, self.importance_sample_ratio = num_points, oversample_ratio, importance_sample_ratio
        self.directions = directions
        from .noise_robust_loss import ProjectionMaskLoss  # binds `custom_cuda_utils` like the reference (models/criterion.py:19, 136)

        self.noise_robust_projection_loss = ProjectionMaskLoss(directions=directions)

    def loss_labels(self, outputs, targets, indices, num_masks, mask_type, coords=None):
        logits = outputs["pred_logits"].float()
        idx = self._get_src_permutation_idx(indices)
        matched = torch.cat([t["labels"][J] for t, (_, J) in zip(targets, indices)])
        target_classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        target_classes[idx] = matched
        return {"loss_ce": F.cross_entropy(logits.transpose(1, 2), target_classes, self.empty_weight, ignore_index=253)}

    def loss_masks(self, outputs, targets, indices, num_masks, mask_type="masks", coords=None):
        use_robust = self.weight_dict.get("loss_noise_robust", 0) != 0
        ce, dice, robust = [], [], []
        core = type(self).mask_loss_core or _cuda_mask_losses
        for b, (map_id, target_id) in enumerate(indices):
            logits = outputs["pred_masks"][b]                 # [S, Q]
            tgt_all = targets[b][mask_type]                   # [T_all, S]
            if use_robust:  # models/criterion.py:170-179: the matched masks at point resolution, projected onto the three planes
                pred = logits[:, map_id].T
                if coords.shape[0] != pred.shape[1]:
                    pred = pred[:, targets[b]["point2segment"]]
                scene = coords[:, 0] == b
                batch_loss, all_shape = self.noise_robust_projection_loss(pred, targets[b]["masks"][target_id].float(), coords[scene])
                robust.append(batch_loss / all_shape)
            else:
                robust.append(torch.as_tensor(0.0, dtype=torch.float32, device=logits.device))
            if self.num_points != -1:  # sub-sample the points shared by all masks (models/criterion.py:184-191)
                pidx = torch.randperm(tgt_all.shape[1], device=tgt_all.device)[:int(self.num_points * tgt_all.shape[1])]
                logits, tgt_all = logits[pidx], tgt_all[:, pidx]
            n_scene = len(target_id)
            weights = None
            if self.use_droploss:
                pred = logits[:, map_id].T
                tgt = tgt_all[target_id]
                fg = pred > 0.0
                iou = (fg * tgt).sum(dim=1) / (fg + tgt).sum(dim=1)
                weights = (iou >= self.droploss_iou_thresh).float()
            l_ce, l_dice = core(logits, tgt_all, map_id, target_id, weights, n_scene)
            ce.append(l_ce)
            dice.append(l_dice)
        return {"loss_mask": torch.sum(torch.stack(ce)), "loss_dice": torch.sum(torch.stack(dice)),
                "loss_noise_robust": torch.sum(torch.stack(robust))}

    @staticmethod
    def _get_src_permutation_idx(indices):
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        return batch_idx, torch.cat([src for (src, _) in indices])

    @staticmethod
    def _get_tgt_permutation_idx(indices):
        batch_idx = torch.cat([torch.full_like(tgt, i) for i, (_, tgt) in enumerate(indices)])
        return batch_idx, torch.cat([tgt for (_, tgt) in indices])

    def get_loss(self, loss, outputs, targets, indices, num_masks, mask_type, coords=None):
        table = {"labels": self.loss_labels, "masks": self.loss_masks}
        assert loss in table, f"do you really want to compute {loss} loss?"
        return table[loss](outputs, targets, indices, num_masks, mask_type, coords)

    def _all_assignments(self, outs, targets, mask_type):
        """Hungarian assignments of every decoder output at once, or None when the matcher is not the device-side one.

        The reference matches output by output (models/criterion.py:255-276) and every match reads its B cost matrices back to the
        host — 13 x B stream synchronisations per step, each draining the device queue in the middle of the step.  The cost
        matrices do not depend on earlier assignments, so all of them are computed first and cross the bus in ONE copy; the
        assignments are the same.  Only without point sub-sampling: with it the order of the random permutations is part of the
        result (matcher and losses alternate in the reference)."""
        m = self.matcher
        if not hasattr(m, "cost_matrices") or getattr(m, "num_points", 0) != -1 or self.num_points != -1:
            return None
        from scipy.optimize import linear_sum_assignment

        costs = [m.cost_matrices(o, targets, mask_type) for o in outs]
        flat = torch.cat([c.reshape(-1).float() for cs in costs for c in cs]).cpu()
        res, off = [], 0
        for cs in costs:
            per_scene = []
            for c in cs:
                n = c.numel()
                i, j = linear_sum_assignment(flat[off:off + n].view(c.shape))
                off += n
                per_scene.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
            res.append(per_scene)
        return res

    def forward(self, outputs, targets, mask_type, coords=None):
        main = {k: v for k, v in outputs.items() if k != "aux_outputs"}
        aux_outputs = outputs.get("aux_outputs", [])
        assigned = self._all_assignments([main] + list(aux_outputs), targets, mask_type)
        indices = assigned[0] if assigned is not None else self.matcher(main, targets, mask_type)
        n_targets = sum(len(t["labels"]) for t in targets)
        if dist.is_available() and dist.is_initialized():
            num_masks = torch.as_tensor([n_targets], dtype=torch.float, device=next(iter(outputs.values())).device)
            dist.all_reduce(num_masks)
            num_masks = torch.clamp(num_masks / _world_size(), min=1).item()
        else:  # single process: the count is known on the host (no device round trip)
            num_masks = float(max(n_targets, 1))
        losses = {}
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, indices, num_masks, mask_type, coords))
        for i, aux in enumerate(aux_outputs):
            indices = assigned[i + 1] if assigned is not None else self.matcher(aux, targets, mask_type)
            for loss in self.losses:
                losses.update({f"{k}_{i}": v for k, v in
                               self.get_loss(loss, aux, targets, indices, num_masks, mask_type, coords).items()})
        return losses
