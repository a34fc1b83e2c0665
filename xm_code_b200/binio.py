"""The reference's ``.bin`` wire format (utils/io.py:17-58, XM_main.cu:18-33,283-305):
int32 rows, int32 cols, then rows*cols float64 in COLUMN-MAJOR order."""
from __future__ import annotations

import numpy as np


def load_matrix_from_bin(filename: str) -> np.ndarray:
    with open(filename, "rb") as f:
        rows = int.from_bytes(f.read(4), "little")
        cols = int.from_bytes(f.read(4), "little")
        data = np.fromfile(f, dtype=np.float64, count=rows * cols)
    if data.size != rows * cols:
        raise IOError(f"{filename}: expected {rows}x{cols} doubles, got {data.size}")
    return data.reshape((rows, cols), order="F")


def save_matrix_to_bin(filename: str, matrix) -> None:
    m = np.asarray(matrix, dtype=np.float64)
    if m.ndim == 1:
        m = m[:, None]
    with open(filename, "wb") as f:
        f.write(int(m.shape[0]).to_bytes(4, "little"))
        f.write(int(m.shape[1]).to_bytes(4, "little"))
        np.asfortranarray(m).ravel(order="F").tofile(f)
