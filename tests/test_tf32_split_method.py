"""The 3xTF32 error-compensation scheme of K3 (cleora_b200/csrc/whiten_tc.cu), emulated with numpy: each f32 operand
is split a = a_hi + a_lo (a_hi = the top 19 bits), the tensor core sees tf32 operands (10-bit mantissa) and the kernel
accumulates A_lo*B_hi + A_hi*B_lo + A_hi*B_hi in f32.  Pins the METHOD's accuracy claim (fp32-class, inside the 1e-5
bar of pycleora/__init__.py:157-163's f32 matmul) on the CPU; the kernel is checked on the GPU."""
import numpy as np


def tf32(a):
    """What the tensor core keeps of an f32 operand: sign, 8 exponent bits, 10 mantissa bits (low 13 bits dropped)."""
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(a):
    hi = tf32(a)
    lo = (a.astype(np.float32) - hi).astype(np.float32)      # exact in f32
    return hi, tf32(lo)                                      # the low part is itself seen through tf32


def matmul_f32_acc(a, b):
    """Products of tf32 operands are exact in f32 (11 x 11 significant bits); accumulation is f32."""
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)


def test_three_term_split_is_fp32_class():
    rs = np.random.default_rng(0)
    n, d = 2048, 256
    x = rs.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x -= x.mean(0)
    t = (rs.standard_normal((d, d)) * np.linspace(1, 30, d)).astype(np.float32)       # whitening-like column scales
    ref = x.astype(np.float64) @ t.astype(np.float64)
    scale = np.max(np.abs(ref))
    xh, xl = split(x)
    th, tl = split(t)
    one = matmul_f32_acc(xh, th)
    three = matmul_f32_acc(xl, th) + matmul_f32_acc(xh, tl) + one                        # small terms first
    err1 = np.max(np.abs(one - ref)) / scale
    err3 = np.max(np.abs(three - ref)) / scale
    f32 = np.max(np.abs((x @ t) - ref)) / scale
    assert err1 > 1e-4            # plain TF32 would miss the 1e-5 bar
    assert err3 < 1e-5            # the compensated product is inside it
    assert err3 < 20 * f32        # ... and within an order of magnitude of a true f32 matmul
