"""Deterministic inputs shared by oracle/wasm_ref/gen_golden.py (which produced the fixtures by
running real pixo) and the tests that consume them.  numpy only."""
import numpy as np

from pixo_b200 import synthetic


def make_input(kind: str, w: int, h: int, ch: int, seed: int) -> np.ndarray:
    if kind == "noise":
        return synthetic.noise(w, h, ch, seed)
    if kind == "gradient":
        assert ch == 3
        return synthetic.gradient_rgb(w, h)
    rng = np.random.default_rng(seed)
    if kind == "primaries":  # saturated colours: exercises the live Cb/Cr clamp
        pal = np.array([[0, 0, 255], [255, 0, 0], [0, 255, 0], [255, 255, 255], [0, 0, 0], [255, 255, 0],
                        [0, 255, 255], [255, 0, 255]], np.uint8)
        idx = (np.arange(h)[:, None] // 3 + np.arange(w)[None, :] // 5) % 8
        img = pal[idx]
        if ch == 1:
            img = img[..., :1]
        elif ch == 4:
            img = np.concatenate([img, np.full((h, w, 1), 200, np.uint8)], -1)
        return np.ascontiguousarray(img).reshape(-1)
    if kind == "smooth":
        x = np.cumsum(rng.integers(-2, 3, (h, w * ch)), axis=1) + np.cumsum(rng.integers(-1, 2, (h, 1)), axis=0)
        return (x & 255).astype(np.uint8).reshape(-1)
    if kind == "vgrad":  # rows nearly equal to the row above: Up / Paeth territory
        base = rng.integers(0, 256, (1, w * ch))
        x = base + np.arange(h)[:, None] * 2 + rng.integers(0, 2, (h, w * ch))
        return (x & 255).astype(np.uint8).reshape(-1)
    if kind == "mixed":  # bands of different statistics so different filters win on different rows
        rows = []
        for y in range(h):
            m = (y // 7) % 4
            if m == 0:
                r = rng.integers(0, 256, w * ch)
            elif m == 1:
                r = np.cumsum(rng.integers(-1, 2, w * ch)) + 128
            elif m == 2:
                r = (rows[-1] if rows else np.zeros(w * ch)) + rng.integers(0, 2, w * ch)
            else:
                r = np.full(w * ch, (y * 37) & 255)
            rows.append(np.asarray(r) & 255)
        return np.stack(rows).astype(np.uint8).reshape(-1)
    raise ValueError(kind)
