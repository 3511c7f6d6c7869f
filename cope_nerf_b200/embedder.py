"""Positional encoding (reference: model/neus_embedder.py:6-51).  Inside the SDF / colour networks the encoding
is fused into the MLP kernels; `get_embedder` is kept for API parity and runs the standalone kernel."""
import torch

from . import _lib as L

__all__ = ["get_embedder", "embed_dim", "Embedder"]


def embed_dim(d, n_freqs):
    return d * (1 + 2 * n_freqs)


class Embedder:
    """[x | sin(2^k x) | cos(2^k x)]_{k<L}, blocks `input_dims` wide (log-sampled bands, include_input=True)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not (kwargs.get("include_input", True) and kwargs.get("log_sampling", True)):
            raise NotImplementedError("only include_input=True, log_sampling=True (what get_embedder builds)")
        self.input_dims = kwargs["input_dims"]
        self.num_freqs = kwargs["num_freqs"]
        self.out_dim = embed_dim(self.input_dims, self.num_freqs)

    def embed(self, inputs):
        x = inputs.detach().contiguous().float()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.input_dims)
        out = torch.empty(x2.shape[0], self.out_dim, dtype=torch.float32, device=x.device)
        L.call("cope_embed_fwd", L.ptr(x2), x2.shape[0], self.input_dims, self.num_freqs, L.ptr(out), L.stream())
        return out.reshape(*lead, self.out_dim)


def get_embedder(multires, input_dims=3):
    eo = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=eo):
        return eo.embed(x)
    return embed, eo.out_dim
