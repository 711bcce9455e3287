"""GaLore low-rank gradient projector (twin of bitorch_engine/optim/galore_projector.py:17-119).  The SVD runs in
torch (cuSOLVER) exactly as in the reference -- it fires every `update_proj_gap` steps and is library work, not part of
the streaming hot path."""
import torch


class GaLoreProjector:
    def __init__(self, rank, verbose=False, update_proj_gap=200, scale=1.0, proj_type="std"):
        self.rank, self.verbose, self.update_proj_gap, self.scale, self.proj_type = rank, verbose, update_proj_gap, scale, proj_type
        self.ortho_matrix = None

    # which side(s) of the gradient the orthogonal factor multiplies
    def _side(self, shape):
        if self.proj_type == "std":
            return "right" if shape[0] >= shape[1] else "left"
        if self.proj_type == "reverse_std":
            return "left" if shape[0] >= shape[1] else "right"
        if self.proj_type in ("right", "left", "full"):
            return self.proj_type
        raise ValueError(f"unknown proj_type {self.proj_type}")

    def project(self, full_rank_grad, it):
        side = self._side(full_rank_grad.shape)
        if self.ortho_matrix is None or it % self.update_proj_gap == 0:
            self.ortho_matrix = self.get_orthogonal_matrix(full_rank_grad, self.rank, type=side)
        if side == "right":
            return torch.matmul(full_rank_grad, self.ortho_matrix.t())
        if side == "left":
            return torch.matmul(self.ortho_matrix.t(), full_rank_grad)
        return torch.matmul(self.ortho_matrix[0].t(), full_rank_grad) @ self.ortho_matrix[1].t()

    def project_back(self, low_rank_grad):
        """galore_projector.py:85-107: the side follows proj_type (and, for std / reverse_std, the shape of the LOW-rank
        gradient) -- not the shapes of the stored factor, which are ambiguous for square weights."""
        m, t = self.ortho_matrix, self.proj_type
        tall = low_rank_grad.shape[0] >= low_rank_grad.shape[1]
        if t == "std":
            out = torch.matmul(low_rank_grad, m) if tall else torch.matmul(m, low_rank_grad)
        elif t == "reverse_std":
            wide = low_rank_grad.shape[0] <= low_rank_grad.shape[1]
            out = torch.matmul(m, low_rank_grad) if wide else torch.matmul(low_rank_grad, m)
        elif t == "right":
            out = torch.matmul(low_rank_grad, m)
        elif t == "left":
            out = torch.matmul(m, low_rank_grad)
        elif t == "full":
            out = torch.matmul(m[0], low_rank_grad) @ m[1]
        else:
            raise ValueError(f"unknown proj_type {t}")
        return out * self.scale

    def get_orthogonal_matrix(self, weights, rank, type):
        data = weights.data
        orig_dtype, orig_device = data.dtype, data.device
        mat = data.float() if orig_dtype != torch.float else data
        U, s, Vh = torch.linalg.svd(mat, full_matrices=False)
        conv = (lambda t: t.to(orig_device).type(orig_dtype)) if orig_dtype != torch.float else (lambda t: t)
        if type == "right":
            return conv(Vh[:rank, :])
        if type == "left":
            return conv(U[:, :rank])
        if type == "full":
            return [conv(U[:, :rank]), conv(Vh[:rank, :])]
        raise ValueError("type should be left, right or full")
