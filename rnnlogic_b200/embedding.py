"""RotatE entity-feature scorer with the reference's interface (src/embedding.py:6-70)."""
import json
import os

import numpy as np
import torch

from . import _lib


class RotatE(torch.nn.Module):
    def __init__(self, path):
        super(RotatE, self).__init__()
        self.path = path
        with open(os.path.join(path, 'config.json'), 'r') as fi:
            cfg = json.load(fi)
        self.emb_dim = cfg['hidden_dim']
        self.gamma = cfg['gamma']
        self.range = (self.gamma + 2.0) / self.emb_dim
        self.num_entities = cfg['nentity']
        eemb = np.load(os.path.join(path, 'entity_embedding.npy'))
        self.eemb = torch.nn.parameter.Parameter(torch.tensor(eemb))
        remb = torch.tensor(np.load(os.path.join(path, 'relation_embedding.npy')))
        self.remb = torch.nn.parameter.Parameter(torch.cat([remb, -remb], dim=0))     # embedding.py:23-26

    def slot_scores(self, sk, sl):
        """Entity-major logits [S][N][32] of gamma - sum_d |h o rot(r) - e|_d for every entity."""
        from .rotate import rotate_slot_scores
        return rotate_slot_scores(self, sk, sl)

    def forward(self, all_h, all_r):
        """fp32[B,N] (embedding.py:64-70) through the same kernel."""
        from .rotate import rotate_dense_scores
        return rotate_dense_scores(self, all_h, all_r)
