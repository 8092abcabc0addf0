"""CPU model of the slack the FP16-accumulator Chamfer filter keeps (csrc/chamfer_tc.cu).

The filter compares hot values that are RN16 of the f32 hot values.  If the f32 values of a true arg-min x and of the
best candidate b satisfy the round-1 relation  x_f <= b_f (1 + r) + a  (r = 2^-14, a = 2.5e-4 rho^2 S^2 + 2^-20), the
kernel must still flag x after both were rounded to f16 (and clamped at 0):  x_h <= b_h + r' |b_h| + a'  with the
constants it uses, r' = 2^-14 + 1.125 * 2^-10, a' = 2.5025e-4 rho^2 S^2 + 1.1930e-6.  Checked here with numpy's IEEE
float16 conversion over the whole range of magnitudes, sub-normals and the worst case x_f = b_f (1 + r) + a included."""
import numpy as np

R_OLD = np.float32(6.103515625e-05)
R_NEW = np.float32(6.103515625e-05 + 1.0986328125e-03)


def thr(x, rel, abs_):
    # __fadd_ru(__fmaf_ru(|x|, rel, x), abs): evaluated in float64 and rounded UP to float32 (never below the kernel's value
    # by more than an ulp: the comparison below keeps one ulp of margin on the safe side)
    v = np.abs(x).astype(np.float64) * np.float64(rel) + x.astype(np.float64) + np.float64(abs_)
    f = v.astype(np.float32)
    return np.where(f.astype(np.float64) < v, np.nextafter(f, np.float32(np.inf)), f)


def rn16_clamped(f):
    h = f.astype(np.float16).astype(np.float32)          # IEEE round to nearest even, overflow -> inf
    return np.maximum(h, np.float32(0.0))


def test_f16_rounding_never_drops_the_argmin():
    rng = np.random.default_rng(7)
    n = 400_000
    # best f32 hot value: log-uniform over the filter's range (scaled units: |P|, |T| < 128), plus exact zeros, tiny negatives
    b = np.exp(rng.uniform(np.log(1e-9), np.log(6.5e4), n)).astype(np.float32)
    b[: n // 50] = 0.0
    b[n // 50: n // 25] = -np.abs(rng.normal(0, 1e-6, n // 25 - n // 50)).astype(np.float32)
    rho2s2 = np.exp(rng.uniform(np.log(1e-6), np.log(1.6e4), n)).astype(np.float32)      # rho^2 S^2 <= 128^2
    a_old = (np.float32(2.5e-4) * rho2s2 + np.float32(9.5367431640625e-07)).astype(np.float32)
    a_new = thr(np.zeros(n, np.float32), 0.0, np.float64(2.5025e-4) * rho2s2.astype(np.float64) + 1.1930e-06)
    # the arg-min's f32 hot value anywhere up to the round-1 threshold (the worst case is the threshold itself)
    hi = thr(b, R_OLD, a_old)
    t = rng.uniform(0, 1, n).astype(np.float32)
    t[: n // 4] = 1.0
    x = (b + (hi - b) * t).astype(np.float32)
    x = np.minimum(x, hi)
    keep = hi < 65400.0                                      # the kernel's scale keeps every value below 65 403
    xh, bh = rn16_clamped(x)[keep], rn16_clamped(b)[keep]
    limit = thr(bh, R_NEW, a_new[keep])
    bad = xh > limit
    assert not bad.any(), (x[keep][bad][:5], b[keep][bad][:5], xh[bad][:5], limit[bad][:5])
    # the test has teeth: with the f32-accumulator slack the rounded values WOULD drop arg-mins
    assert (xh > thr(bh, R_OLD, a_old[keep])).any()


def test_integer_order_of_f16_patterns():
    """Non-negative f16 numbers order like their bit patterns as signed 16-bit integers; every negative pattern is a
    negative integer (it wins an integer min and is then clamped to +0); +inf is the largest pattern."""
    pats = np.arange(0, 0x7C01, dtype=np.uint16)                                  # +0 ... +inf
    vals = pats.view(np.float16).astype(np.float32)
    assert np.all(np.diff(vals) > 0)
    assert np.all(pats.view(np.int16)[1:] > pats.view(np.int16)[:-1])
    neg = np.arange(0x8000, 0xFC01, dtype=np.uint32).astype(np.uint16)            # -0 ... -inf
    assert np.all(neg.view(np.int16) < 0)
    assert np.isposinf(np.uint16(0x7C00).view(np.float16))
