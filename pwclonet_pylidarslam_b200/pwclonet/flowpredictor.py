"""reference PW/flowpredictor.py:15-84"""
import torch
import torch.nn as nn

from .. import pytorch_utils as pt_utils


class FlowPredictor(nn.Module):
    def __init__(self, in_channel, mlp, bn_decay=None):
        super().__init__()
        self.in_channel = [in_channel]
        self.mlp_convs = pt_utils.SharedMLP([in_channel] + list(mlp), bn=True, init=torch.nn.init.xavier_uniform_)
        self.out_channel = mlp[-1]

    def forward(self, points_f1, cost_volume, upsampled_feat=None):
        """(B,C1,N), (B,C2,N)[, (B,C',N)] -> (B,mlp[-1],N)"""
        if points_f1 is None:
            x = cost_volume
        elif upsampled_feat is not None:
            x = torch.cat((points_f1, cost_volume, upsampled_feat), dim=1)
        else:
            x = torch.cat((points_f1, cost_volume), dim=1)
        return self.mlp_convs(x.unsqueeze(3)).squeeze(3)
