"""Shared input generators for the parity tests (restating tests/support/synthetic.rs)."""
import numpy as np

# tests/support/synthetic.rs:276-290
EDGE_CASE_DIMENSIONS = [(1, 1), (2, 2), (7, 7), (8, 8), (9, 9), (16, 16), (15, 17), (1, 100),
                        (100, 1), (256, 256), (512, 512), (1000, 1000), (1024, 1024)]


def images(po, w, h, channels=3):
    """gradient, LCG noise, extremes (pure primaries exercise the live Cb/Cr clamp), random."""
    out = {}
    if channels == 3:
        out["gradient"] = po.gen_gradient_rgb(w, h)
    out["noise"] = po.gen_noise(w, h, channels, 42)
    prim = np.zeros((h, w, channels), np.uint8)
    pal = np.array([[0, 0, 255], [255, 0, 0], [0, 255, 0], [255, 255, 255], [0, 0, 0],
                    [255, 255, 0], [0, 255, 255], [255, 0, 255]], np.uint8)
    idx = (np.arange(h)[:, None] // 3 + np.arange(w)[None, :] // 5) % 8
    prim[...] = pal[idx][..., :channels] if channels <= 3 else np.concatenate(
        [pal[idx], np.full((h, w, 1), 255, np.uint8)], -1)
    out["primaries"] = prim.reshape(-1)
    return out
