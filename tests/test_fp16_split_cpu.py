"""CPU model of the operand split of the tensor-core field sum (csrc/separable.cu): fp32 factors split as
x = hi + lo with hi = fp16(x), lo = fp16(x - hi) after the pre-scaling (row factors <= 2^14 at the brightest
peak, column factors <= 2^14), products hi*hi + hi*lo + lo*hi.  The split must cost far less than the 1e-5
parity gate -- and no more than the tf32 x 3 split it replaced -- on wide, narrow and amplitude-spread beamlet
sets.  (Products and sums are exact here; the accumulation error of the kernel is measured on the GPU.)"""
import numpy as np
import pytest


def _tf32(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def _split_tf32(x):
    x = x.astype(np.float32)
    hi = _tf32(x)
    return hi.astype(np.float64), _tf32((x - hi).astype(np.float32)).astype(np.float64)


def _split_f16(x):
    x = x.astype(np.float32)
    hi = x.astype(np.float16).astype(np.float32)
    return hi.astype(np.float64), (x - hi).astype(np.float16).astype(np.float64)


def _case(rng, nb, H, W, width, amp_decades):
    r, c = np.arange(H)[None, :], np.arange(W)[None, :]
    r0, c0 = rng.uniform(0, H, (nb, 1)), rng.uniform(0, W, (nb, 1))
    amp = 10 ** rng.uniform(-amp_decades, 0, (nb, 1))
    U = amp * np.exp2(-((r - r0) / width) ** 2) * np.exp(2j * np.pi * rng.uniform(0, 1, (nb, H)))
    V = np.exp2(-((c - c0) / width) ** 2) * np.exp(2j * np.pi * rng.uniform(0, 1, (nb, W)))
    # pre-scaling of the kernels: brightest row-factor peak -> 2^14, column factors (<= 1) -> 2^14
    G = np.ceil(np.log2(np.abs(U).max()))
    return U * 2.0 ** (14 - G), V * 2.0 ** 14


def _cgemm(Ur, Ui, Vr, Vi):
    return (Ur.T @ Vr - Ui.T @ Vi) + 1j * (Ur.T @ Vi + Ui.T @ Vr)


def _x3(split, U, V):
    Urh, Url = split(U.real)
    Uih, Uil = split(U.imag)
    Vrh, Vrl = split(V.real)
    Vih, Vil = split(V.imag)
    return _cgemm(Urh, Uih, Vrh, Vih) + _cgemm(Urh, Uih, Vrl, Vil) + _cgemm(Url, Uil, Vrh, Vih)


@pytest.mark.parametrize("width,decades", [(1e4, 0), (4, 0), (1e4, 6), (4, 4), (40, 2)])
def test_fp16x3_split_is_fp32_equivalent(width, decades):
    rng = np.random.default_rng(0)
    U, V = _case(rng, 1500, 96, 96, width, decades)
    U32, V32 = U.astype(np.complex64).astype(np.complex128), V.astype(np.complex64).astype(np.complex128)
    exact = U32.T @ V32
    err16 = np.linalg.norm(_x3(_split_f16, U32, V32) - exact) / np.linalg.norm(exact)
    err32 = np.linalg.norm(_x3(_split_tf32, U32, V32) - exact) / np.linalg.norm(exact)
    assert err16 < 3e-7, err16                 # 30x below the 1e-5 gate, below the ~1e-6 MUFU error of the factors
    assert err16 < 3 * err32 + 1e-8            # no worse than the tf32 x 3 split beyond a small factor
    for part in (U.real, U.imag, V.real, V.imag):                                 # no fp16 overflow
        assert np.isfinite(part.astype(np.float32).astype(np.float16)).all()
