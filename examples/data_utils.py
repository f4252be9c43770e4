"""Synthetic data helpers (reference: examples/language/data_utils.py `RandomDataset`, model_utils.py)."""
import torch
from torch.utils.data import Dataset


class RandomDataset(Dataset):
    """Random token ids + full attention mask, generated once on the host (pinned) so loaders exercise real H2D."""

    def __init__(self, num_samples: int = 1000, max_length: int = 2048, vocab_size: int = 32000, seed: int = 42):
        g = torch.Generator().manual_seed(seed)
        self.input_ids = torch.randint(0, vocab_size, (num_samples, max_length), generator=g)
        self.attention_mask = torch.ones_like(self.input_ids)

    def __len__(self):
        return self.input_ids.shape[0]

    def __getitem__(self, idx):
        return {"input_ids": self.input_ids[idx], "attention_mask": self.attention_mask[idx],
                "labels": self.input_ids[idx]}


def get_model_numel(model: torch.nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def format_numel_str(numel: int) -> str:
    for unit, div in (("B", 1e9), ("M", 1e6), ("K", 1e3)):
        if numel >= div:
            return f"{numel / div:.2f} {unit}"
    return str(numel)
