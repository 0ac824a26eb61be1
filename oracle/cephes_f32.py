"""Single-precision Bessel J0/J1 as TensorFlow evaluates them on float32 tensors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``tf.math.special.bessel_j0`` (reference call site signals.py:170) and its
registered gradient ``-tf.math.special.bessel_j1`` dispatch to Eigen's
``generic_j0<T, float>`` / ``generic_j1<T, float>``
(unsupported/Eigen/src/SpecialFunctions/BesselFunctionsImpl.h), which are the
Cephes single-precision routines ``j0f`` / ``j1f``.  TensorFlow is an un-vendored,
un-pinned dependency of the reference (requirements.txt:3, ``tensorflow>=2.5.0``),
so the published Cephes algorithm is restated here (coefficients as listed in
SURVEY.md Appendix C; validated there against scipy to 2e-7 abs on [0, 30]).

Every operation is carried out on ``np.float32`` arrays so each elementary op is
rounded to float32 (no FMA contraction; Eigen may contract ``pmadd`` -- a <=1 ulp
effect, far below the 1e-5 parity bar).
"""
import numpy as np

f32 = np.float32

_DR1 = f32(5.78318596294678452118)
_JP = [f32(v) for v in (-6.068350350393235e-8, 6.388945720783375e-6, -3.969646342510940e-4,
                        1.332913422519003e-2, -1.729150680240724e-1)]
_MO = [f32(v) for v in (-6.838999669318810e-2, 1.864949361379502e-1, -2.145007480346739e-1,
                        1.197549369473540e-1, -3.560281861530129e-3, -4.969382655296620e-2,
                        -3.355424622293709e-6, 7.978845717621440e-1)]
_PH = [f32(v) for v in (3.242077816988247e1, -3.630592630518434e1, 1.756221482109099e1,
                        -4.974978466280903e0, 1.001973420681837e0, -1.939906941791308e-1,
                        6.490598792654666e-2, -1.249992184872738e-1)]
_PIO4F = f32(0.7853981633974483096)

_Z1 = f32(14.6819706421238932572)
_JP1 = [f32(v) for v in (-4.878788132172128e-9, 6.009061827883699e-7, -4.541343896997497e-5,
                         1.937383947804541e-3, -3.405537384615824e-2)]
_MO1 = [f32(v) for v in (6.913942741265801e-2, -2.284801500053359e-1, 3.138238455499697e-1,
                         -2.102302420403875e-1, 5.435364690523026e-3, 1.493389585089498e-1,
                         4.976029650847191e-6, 7.978845453073848e-1)]
_PH1 = [f32(v) for v in (-4.497014141919556e1, 5.073465654089319e1, -2.485774108720340e1,
                         7.222973196770240e0, -1.544842782180211e0, 3.503787691653334e-1,
                         -1.637986776941202e-1, 3.749989509080821e-1)]
_THPIO4F = f32(2.35619449019234492885)


def _polevl(x, coef):
    """Horner, coef[0] is the leading coefficient; float32 mul then add."""
    acc = np.full_like(x, coef[0])
    for c in coef[1:]:
        acc = acc * x + c
    return acc


def _cos_f32(x):
    # float32 cosine: evaluate in double and round once (<= 0.5 ulp), which is
    # at least as accurate as Eigen's pcos<float> (<= 1-2 ulp for |x| < ~1e4).
    return np.cos(x.astype(np.float64)).astype(np.float32)


def j0f(x):
    """Cephes j0f on a float32 array."""
    x = np.asarray(x, dtype=np.float32)
    y = np.abs(x)
    z = y * y
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        tiny = f32(1.0) - f32(0.25) * z
        small = (z - _DR1) * _polevl(z, _JP)
        q = f32(1.0) / y
        w = np.sqrt(q)
        p = w * _polevl(q, _MO)
        yn = q * _polevl(q * q, _PH) - _PIO4F
        big = p * _cos_f32(yn + y)
    out = np.where(y <= f32(2.0), np.where(y < f32(1.0e-3), tiny, small), big)
    return out.astype(np.float32)


def j1f(x):
    """Cephes j1f on a float32 array (odd in x)."""
    x = np.asarray(x, dtype=np.float32)
    y = np.abs(x)
    z = y * y
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        small = (z - _Z1) * y * _polevl(z, _JP1)
        q = f32(1.0) / y
        w = np.sqrt(q)
        p = w * _polevl(q, _MO1)
        yn = q * _polevl(q * q, _PH1) - _THPIO4F
        big = p * _cos_f32(yn + y)
    out = np.where(y <= f32(2.0), small, big)
    return (np.sign(x) * out).astype(np.float32)
