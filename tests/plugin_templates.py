"""A user-defined template class, as a caller of the reference would write one against the
``WindowedTemplate`` duck type (core.py:345-346, 369-375): none of the library's on-device
generators knows it, so it exercises the generic plugin path."""
import numpy as np


class Ridge(object):
    """Gaussian ridge with a zero-mean profile across it: ``age`` is the ridge's standard
    deviation (m), ``scale`` its half length; the SNR is only trusted on one flank."""

    def __init__(self, d, s, alpha, nx, ny, de):
        self.d = d
        self.s = s
        self.alpha = -alpha
        self.nx = nx
        self.ny = ny
        self.de = de
        self.c = 3.0 * s

    def get_coordinates(self):
        x = self.de * np.linspace(1, self.nx, num=self.nx)
        y = self.de * np.linspace(1, self.ny, num=self.ny)
        x = x - np.mean(x)
        y = y - np.mean(y)
        x, y = np.meshgrid(x, y)
        xr = x * np.cos(self.alpha) + y * np.sin(self.alpha)
        yr = -x * np.sin(self.alpha) + y * np.cos(self.alpha)
        return xr, yr

    def template(self):
        xr, yr = self.get_coordinates()
        mask = (abs(xr) < self.c) & (abs(yr) < self.d)
        u2 = (xr / self.s) ** 2
        return (u2 - 1.0) * np.exp(-0.5 * u2) * mask

    def get_window_limits(self):
        m = int(np.ceil((self.c + self.d) / self.de))
        out = np.ones((self.ny, self.nx), dtype=bool)
        out[m:self.ny - m, m:self.nx - m] = False
        # a plugin's mask need not be a rectangle
        out[self.ny // 3, :] = True
        return out

    def get_err_mask(self):
        _, yr = self.get_coordinates()
        return yr > 0.75 * self.d
