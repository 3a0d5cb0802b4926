from numpy.fft import fft2, ifft2, fftshift, ifftshift, fft, ifft  # noqa: F401
