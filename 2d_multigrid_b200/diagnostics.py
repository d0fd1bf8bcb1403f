"""Error-spectrum diagnostics: what the reference notebooks compute offline from the per-iteration
text dumps (NB/2_spectral_analysis_solution.ipynb cell 5: 2D FFT of phi - phi* per iteration) done on the device with
torch.fft (cuFFT) -- post-processing next to the hot path, not part of it."""
from __future__ import annotations

import torch


def error_spectrum(phi: torch.Tensor, phi_star: torch.Tensor, L: int) -> torch.Tensor:
    """|FFT2(phi - phi*)| per dof: returns [n, L, L] magnitudes, index [dof, ky, kx] (site s = x + y*L)."""
    e = (phi - phi_star).reshape(L, L, -1).permute(2, 0, 1)
    return torch.fft.fft2(e).abs()


def mode_amplitudes(phi: torch.Tensor, phi_star: torch.Tensor, L: int):
    """Summary used in the notebooks' plots: the largest low-frequency and high-frequency error amplitudes
    (|k| <= L/4 vs the rest, per component maximum)."""
    spec = error_spectrum(phi, phi_star, L)
    k = torch.fft.fftfreq(L, d=1.0 / L, device=spec.device).abs()
    low = (k[:, None] <= L // 4) & (k[None, :] <= L // 4)
    return float(spec[:, low].max()), float(spec[:, ~low].max())
