"""Llama shapes of the BASELINE configs.  Torch-free on purpose: bench.py's `--impl reference`
arm imports this and nothing else of the package, so that arm loads none of this repo's native code."""
from dataclasses import dataclass


@dataclass
class LlamaShape:
    name: str
    hidden: int
    inter: int
    layers: int
    heads: int
    head_dim: int = 128
    vocab: int = 32000
    eps: float = 1e-6

    @property
    def kv_heads(self):
        return self.heads


SHAPES = {
    "7b": LlamaShape("llama-7b", 4096, 11008, 32, 32),
    "13b": LlamaShape("llama-13b", 5120, 13824, 40, 40),
    "33b": LlamaShape("llama-33b", 6656, 17920, 60, 52),
    "65b": LlamaShape("llama-65b", 8192, 22016, 80, 64),
}


def q4_bytes(K, N, gs, M=1, act=False, accumulate=False):
    """Algorithmic bytes of one q4_matmul."""
    return K * N // 2 + 2 * (K // gs) * N + 4 * (K // gs) * (N // 8) + 2 * M * K + 2 * M * N * (2 if accumulate else 1) + (4 * K if act else 0)
