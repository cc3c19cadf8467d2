"""Tabulated rasterization-rate maps for the foveated-target tests: the separable piecewise-linear physical -> screen mapping
a MTLRasterizationRateLayerDescriptor(horizontal:vertical:) describes (one rate in (0, 1] per cell of the screen)."""
import numpy as np


def axis_table(screen_extent: float, rates) -> np.ndarray:
    """Screen coordinate of the centre of every physical texel along one axis. The screen is cut into len(rates) equal
    cells; a cell of screen size s and rate r takes ceil(s * r) physical texels."""
    rates = np.asarray(rates, np.float64)
    cell = screen_extent / len(rates)
    out = []
    for i, r in enumerate(rates):
        n = int(np.ceil(cell * r))
        centres = (np.arange(n, dtype=np.float64) + 0.5) / n
        out.append(i * cell + centres * cell)
    return np.concatenate(out).astype(np.float32)


def layer(screen_w: float, screen_h: float, hrates, vrates):
    return axis_table(screen_w, hrates), axis_table(screen_h, vrates)


FOVEATED_H = (0.25, 0.5, 1.0, 1.0, 0.5, 0.25)
FOVEATED_V = (0.3, 0.75, 1.0, 0.6)
