"""Dense tail of PredictorPlus with the reference's module / state_dict names (src/layers.py:9-126).

The reference materialises a [C, R_q, H] broadcast to aggregate rule embeddings onto candidate
entities; here the aggregation statistics come from the CUDA kernel rl_plus_features
(rnnlogic_b200/csrc/rl_plus.cu) and these modules apply the remaining small dense layers."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MLP(nn.Module):
    """Stack of Linear layers with an activation between them (layers.py:9-51)."""

    def __init__(self, input_dim, hidden_dims, short_cut=False, batch_norm=False, activation="relu", dropout=0):
        super(MLP, self).__init__()
        self.dims = [input_dim] + list(hidden_dims)
        self.short_cut = short_cut
        self.activation = getattr(F, activation) if isinstance(activation, str) else activation
        self.dropout = nn.Dropout(dropout) if dropout else None
        self.layers = nn.ModuleList([nn.Linear(self.dims[i], self.dims[i + 1]) for i in range(len(self.dims) - 1)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(d) for d in self.dims[1:-1]]) if batch_norm else None

    def forward(self, input):
        x = input
        for i, layer in enumerate(self.layers):
            hidden = layer(x)
            if i < len(self.layers) - 1:
                if self.batch_norms:
                    hidden = self.batch_norms[i](hidden.flatten(0, -2)).view_as(hidden)
                hidden = self.activation(hidden)
                if self.dropout:
                    hidden = self.dropout(hidden)
            if self.short_cut and hidden.shape == x.shape:
                hidden = hidden + x
            x = hidden
        return hidden


class FuncToNodeSum(nn.Module):
    """sum aggregator (layers.py:53-77): Linear(H,H) -> LayerNorm -> ReLU of counts^T . emb."""

    def __init__(self, vector_dim):
        super(FuncToNodeSum, self).__init__()
        self.vector_dim = vector_dim
        self.layer_norm = nn.LayerNorm(self.vector_dim)
        self.add_model = MLP(self.vector_dim, [self.vector_dim])
        self.eps = 1e-6

    def post(self, features):
        return torch.relu(self.layer_norm(self.add_model(features)))

    def forward(self, A_fn, x_f, b_n):
        """Reference signature: A_fn fp32[R_q,C] counts, x_f fp32[R_q,H] rule embeddings."""
        return self.post(A_fn.t() @ x_f)


class FuncToNode(nn.Module):
    """PNA aggregator (layers.py:79-126): mean/min/max/std x {1, s, 1/s} -> Linear(12H,H) -> LN -> ReLU."""

    def __init__(self, vector_dim):
        super(FuncToNode, self).__init__()
        self.vector_dim = vector_dim
        self.layer_norm = nn.LayerNorm(self.vector_dim)
        self.add_model = MLP(self.vector_dim * 12, [self.vector_dim])
        self.eps = 1e-6

    def post(self, s1, s2, mn, mx, degree, b_n, num_queries):
        """s1 = sum count*emb, s2 = sum count*emb^2, mn/mx over rules with count != 0 ([C,H] each),
        degree[C] = sum count + 1, b_n[C] = query of the candidate (layers.py:100-126)."""
        deg = degree.unsqueeze(-1)
        mean = s1 / deg.clamp(min=self.eps)
        sq_mean = s2 / deg.clamp(min=self.eps)
        std = (sq_mean - mean ** 2).clamp(min=self.eps).sqrt()
        features = torch.cat([mean, mn, mx, std], dim=-1)
        scale = deg.log()
        sum_scale = torch.zeros(num_queries, device=deg.device).scatter_add_(0, b_n, scale.squeeze(-1))
        cn_scale = torch.zeros(num_queries, device=deg.device).scatter_add_(0, b_n, torch.ones_like(scale.squeeze(-1)))
        mean_scale = sum_scale / cn_scale.clamp(min=self.eps)
        scale = scale / mean_scale[b_n].unsqueeze(-1).clamp(min=self.eps)
        scales = torch.cat([torch.ones_like(scale), scale, 1 / scale.clamp(min=self.eps)], dim=-1)
        update = (features.unsqueeze(-1) * scales.unsqueeze(-2)).flatten(-2)
        return torch.relu(self.layer_norm(self.add_model(update)))

    def forward(self, A_fn, x_f, b_n):
        """Reference signature (dense counts); the hot path calls post() with kernel statistics."""
        w = A_fn.t()
        nz = (w != 0).unsqueeze(-1)
        msg = x_f.unsqueeze(0).expand(w.size(0), -1, -1)
        mn = msg.masked_fill(~nz, float("inf")).min(1)[0]
        mx = msg.masked_fill(~nz, float("-inf")).max(1)[0]
        return self.post(w @ x_f, w @ (x_f ** 2), mn, mx, A_fn.sum(0) + 1, b_n, int(b_n.max().item()) + 1)
