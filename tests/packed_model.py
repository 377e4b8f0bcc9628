"""NumPy model of the DATA FLOW the CUDA kernels implement (test infrastructure).

It is not an oracle (the oracle restates the reference; this restates *our* factorisation of the same
transform) — it documents, and lets CPU tests check, the algebra behind the kernels in
watersurfacerendering_b200/csrc/wso_kernels.cuh:

1. h~(k,t) is purely real for reference-built h0 (heightAmp_conj == conj(heightAmp)), so every one of the
   7 spectra is real (H, kx*ux*H, kz*uz*H) or imaginary (i*kx*H, i*kz*H, -i*ux*H, -i*uz*H).
2. Only Re(FFT2) is kept.  Re FFT2(R) = FFT2(even part of R) and Re FFT2(i*V) = FFT2(i * odd part of V),
   "even/odd" under DFT-index reflection (m,n) -> ((N-m)%N,(N-n)%N).
3. Pair an even-type field R with an odd-type field V into ONE REAL array Z = R_even - V_odd; then
   FFT2(Z) = out_R + i*out_V.  7 outputs -> 4 real N x N arrays:
       Z0: (height, Dx)   Z1: (-, Dz)   Z2: (dxDx, slopeX)   Z3: (dzDz, slopeZ)
4. Real input => Hermitian output: pass 1 (along m) transforms column pairs (n, N-n) two-for-one and
   keeps m' in [0, N/2) with the m'=N/2 bin packed into the imaginary part of m'=0; pass 2 (along n)
   does one length-N complex FFT per (field, m') and emits output rows m' and N-m' (conjugate mirror).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def column_of_slot(n: int) -> np.ndarray:
    """Storage slot s in [0,N) of the intermediate W -> column n.  First half natural (pair index j),
    second half holds the mirror partner of pair j: slot N/2 + j <-> column (N - j) % N, j=0 -> N/2."""
    s = np.arange(n)
    j = s - n // 2
    second = np.where(j == 0, n // 2, n - j)
    return np.where(s < n // 2, s, second)


def evolve_Z(n, kv, h0_re, h0_im, omega, t, dtype=np.float64):
    """4 real fields Z[f][m][n] (centred index layout, exactly as the reference's arrays)."""
    ph = (omega * F(t)).astype(np.float32)
    c = np.cos(ph.astype(np.float64)).astype(np.float32)
    s = np.sin(ph.astype(np.float64)).astype(np.float32)
    x = (h0_re * c - h0_im * s).astype(np.float32)
    H = (x + x).astype(np.float32)
    kx, kz = kv[None, :], kv[:, None]
    d = kx * kx + kz * kz
    ln = np.sqrt(d)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = F(1) / np.sqrt(d)
        ux = np.where(ln > F(1e-5), kx * inv, F(0)).astype(np.float32)
        uz = np.where(ln > F(1e-5), kz * inv, F(0)).astype(np.float32)
    # field values at every index (fp32, reference rounding order)
    R = [H, None, kx * (ux * H), kz * (uz * H)]               # even-type (real spectra)
    V = [-ux * H, -uz * H, kx * H, kz * H]                     # odd-type  (imag spectra = i*V)

    def refl(a):  # a[(N-m)%N, (N-n)%N]
        return np.roll(a[::-1, ::-1], (1, 1), axis=(0, 1))

    Z = []
    for f in range(4):
        v = V[f].astype(dtype)
        o = 0.5 * (v - refl(v))
        if R[f] is None:
            Z.append(-o)
        else:
            r = R[f].astype(dtype)
            Z.append(0.5 * (r + refl(r)) - o)
    return np.stack(Z)


def pass1(Z):
    """Z (4,N,N) real -> W (N/2, 4, N) complex, layout [m'][f][slot]."""
    _, n, _ = Z.shape
    h = n // 2
    W = np.zeros((h, 4, n), np.complex128)
    for f in range(4):
        for j in range(h):
            nA = j
            nB = (n - j) % n if j else h
            c = Z[f][:, nA] + 1j * Z[f][:, nB]
            C = np.fft.ifft(c) * n                      # backward, unnormalised
            Cm = np.conj(np.roll(C[::-1], 1))           # conj C[(N-m')%N]
            WA = 0.5 * (C + Cm)
            WB = -0.5j * (C - Cm)
            W[1:, f, j] = WA[1:h]
            W[1:, f, h + j] = WB[1:h]
            W[0, f, j] = WA[0].real + 1j * WA[h].real   # Nyquist packed into imag of DC
            W[0, f, h + j] = WB[0].real + 1j * WB[h].real
    return W


def pass2(W, lam):
    """W (N/2,4,N) -> unnormalised disp (N,N,4), norm (N,N,4), hmin, hmax."""
    h, _, n = W.shape
    col = column_of_slot(n)
    disp = np.zeros((n, n, 4))
    norm = np.zeros((n, n, 4))
    nn = np.arange(n)
    mir = (n - nn) % n
    for mi in range(h):
        Fs = []
        for f in range(4):
            x = np.zeros(n, np.complex128)
            x[col] = W[mi, f]
            Fs.append(np.fft.ifft(x) * n)
        if mi == 0:
            rowA, rowB = 0, h
            FA = [0.5 * (g + np.conj(g[mir])) for g in Fs]
            FB = [-0.5j * (g - np.conj(g[mir])) for g in Fs]
        else:
            rowA, rowB = mi, n - mi
            FA = Fs
            FB = [np.conj(g[mir]) for g in Fs]
        for row, G in ((rowA, FA), (rowB, FB)):
            sg = np.where((row + nn) & 1, -1.0, 1.0)
            disp[row, :, 0] = sg * lam * G[0].imag
            disp[row, :, 1] = sg * G[0].real
            disp[row, :, 2] = sg * lam * G[1].imag
            disp[row, :, 3] = 1.0
            norm[row, :, 0] = sg * G[2].imag
            norm[row, :, 1] = sg * G[3].imag
            norm[row, :, 2] = sg * G[2].real
            norm[row, :, 3] = sg * G[3].real
    hh = disp[..., 1]
    return disp, norm, hh.min(), hh.max()


def compute_waves(n, tile_length, lam, h0_re, h0_im, omega, t):
    idx = np.arange(n, dtype=np.float32)
    kv = (np.pi * (F(2) * idx - F(n)).astype(np.float64) / np.float64(F(tile_length))).astype(np.float32)
    Z = evolve_Z(n, kv, h0_re, h0_im, omega, t)
    W = pass1(Z)
    disp, norm, hmin, hmax = pass2(W, lam)
    hmax = max(hmax, float(np.finfo(np.float32).tiny))
    hmin = min(hmin, float(np.finfo(np.float32).max))
    a = F(max(abs(F(hmin)), abs(F(hmax))))
    disp = disp.astype(np.float32)
    disp[..., 1] = disp[..., 1] * (F(1) / a)
    return a, disp, norm.astype(np.float32), F(hmin), F(hmax)
