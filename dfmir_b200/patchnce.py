"""PatchNCE contrastive loss (reference: models/patchnce.py:7-55), backed by the fused kernels of
libdfmir_b200.so (per-image Q K^T product + masked log-softmax; backward dS K)."""
import torch
from torch import nn

from . import functional as Fn


class PatchNCELoss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.mask_dtype = torch.bool

    def forward(self, feat_q, feat_k):
        feat_k = feat_k.detach()
        # negatives come from the same image unless --nce_includes_all_negatives_from_minibatch
        batch = 1 if self.opt.nce_includes_all_negatives_from_minibatch else self.opt.batch_size
        return Fn.patchnce(feat_q, feat_k, batch, self.opt.nce_T)
