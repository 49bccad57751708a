"""SDFLoss / SDFLossFunction with the reference's names and call signature
(layers/sdf_matching_loss.py:8-66), backed by libomgb200.so."""
import torch
from torch import nn
from torch.autograd import Function

from .engine import sdf_loss_forward


class SDFLossFunction(Function):
    @staticmethod
    def forward(ctx, pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables):
        potentials, potential_grads, collides = sdf_loss_forward(
            pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables)
        return potentials, potential_grads, collides

    @staticmethod
    def backward(ctx, *grads):  # the reference returns no gradients either (sdf_matching_loss.py:37-39)
        return None, None, None, None, None, None, None, None


class SDFLoss(nn.Module):
    def forward(self, pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables):
        return SDFLossFunction.apply(pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales,
                                     clearances, disables)
