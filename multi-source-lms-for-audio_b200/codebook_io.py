"""Codebook on-disk formats shared with the reference.

* writer: `VQVAE.on_train_epoch_end` dumps `codebook.weight` with `DataFrame.to_csv(index=False, header=False)`
  (src/model/vqvae.py:239-243);
* readers: `AudioBert.__init__` / plotting use `pd.read_csv(path)` with the default header handling
  (src/model/bert.py:32, src/utils/plotting.py:19) - which silently eats the first codeword of a header-less file, while the
  committed logs/best_checkpoint/codebook.csv DOES carry a header row `0,1,...,D-1`.
`load_codebook_csv` reads both variants without losing a row; `save_codebook_csv` writes either.
* checkpoints: the key is `vector_quantizer.codebook.weight` (strict loads at src/main.py:66,117,197).
"""
from __future__ import annotations

import numpy as np
import torch

STATE_DICT_KEY = "vector_quantizer.codebook.weight"


def _is_index_header(first_row: np.ndarray) -> bool:
    return np.array_equal(first_row, np.arange(first_row.size, dtype=first_row.dtype))


def load_codebook_csv(path: str, expect_rows: int | None = None) -> torch.Tensor:
    """CSV -> [K, D] fp32.  A first row equal to 0,1,...,D-1 is a pandas column header and is dropped."""
    data = np.loadtxt(path, delimiter=",", dtype=np.float64, ndmin=2)
    if data.shape[0] > 1 and _is_index_header(data[0]) and (expect_rows is None or data.shape[0] == expect_rows + 1):
        data = data[1:]
    if expect_rows is not None and data.shape[0] != expect_rows:
        raise ValueError(f"{path}: {data.shape[0]} codewords, expected {expect_rows}")
    return torch.from_numpy(data.astype(np.float32))


def save_codebook_csv(weight: torch.Tensor, path: str, header: bool = False) -> None:
    """[K, D] -> CSV; header=False reproduces vqvae.py:241-243, header=True the committed file's layout."""
    w = weight.detach().cpu().numpy()
    with open(path, "w") as f:
        if header:
            f.write(",".join(str(i) for i in range(w.shape[1])) + "\n")
        for row in w:
            f.write(",".join(str(np.float32(v)) for v in row) + "\n")


def codebook_from_state_dict(state_dict: dict) -> torch.Tensor:
    """Pull the codebook out of a reference VQVAE checkpoint (`ckpt['state_dict']`)."""
    if STATE_DICT_KEY in state_dict:
        return state_dict[STATE_DICT_KEY]
    if "codebook.weight" in state_dict:           # a bare VectorQuantizer state_dict
        return state_dict["codebook.weight"]
    raise KeyError(f"no {STATE_DICT_KEY} in state_dict")
