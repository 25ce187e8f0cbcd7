"""reference PW/pose_calculator.py:20-87"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..pytorch_utils import Conv1d


class PoseCalculator(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size=1, padding="valid", activation=None, pose=None,
                 squeeze=True, bn_decay=None):
        super().__init__()
        self.pose, self.squeeze = pose, squeeze
        xu = torch.nn.init.xavier_uniform_
        self.conv1d_q_t = Conv1d(in_channel, out_channel, kernel_size=kernel_size, activation=activation, init=xu)
        self.conv1d_q = Conv1d(out_channel, 4, kernel_size=kernel_size, activation=activation, init=xu)
        self.conv1d_t = Conv1d(out_channel, 3, kernel_size=kernel_size, activation=activation, init=xu)

    def forward(self, embedding_features, mask):
        """(B,C,N), (B,C,N) -> q (B,4[,1]), t (B,3[,1])"""
        s = torch.sum(embedding_features * mask, dim=2, keepdim=True)
        big = self.conv1d_q_t(s)
        q = self.conv1d_q(F.dropout(big, p=0.5, training=self.training))
        q = q / (torch.sqrt(torch.sum(q * q, dim=1, keepdim=True) + 1e-10) + 1e-10)
        t = self.conv1d_t(F.dropout(big, p=0.5, training=self.training))
        if self.squeeze:
            q, t = q.squeeze(2), t.squeeze(2)
        return q, t
