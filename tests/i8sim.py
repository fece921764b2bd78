"""TEST INFRASTRUCTURE — NumPy mirror of the int8 digit-slice MLP kernel (cup1d_b200/csrc/c1d_mlp_i8.cu).

Same plan as the kernel: every layer on integer digit planes, inputs quantised too; layers up to 192 wide
share one row scale, wider layers give each of their two n-blocks (128, rest) its own row scale and the next
layer accumulates the two K-groups separately; ten digit pairs (levels 0..3); float64 recombination.
Integer sums are exact here as on the tensor cores, so the device coefficients agree with this mirror to
float64 rounding of the last few operations (1e-13), far inside the quantisation error itself.
"""

import numpy as np

FULL = 127.0 * 16777216.0


def _planes(q):
    """four balanced base-256 digit planes of int64 q (|q| <= FULL), most significant first"""
    u = (q.astype(np.int64) + 0x80808080) & 0xFFFFFFFF
    v = u ^ 0x80808080
    out = []
    for p in range(4):
        b = (v >> (8 * (3 - p))) & 0xFF
        out.append(np.where(b >= 128, b - 256, b).astype(np.int64))
    return out


def _quant_rows(x):
    """x[rows, k] -> (digit planes, scale[rows, 1]); scale = M / FULL with M the kernel's upper bound of max|x|:
    the next value of the high 32-bit word of the float64 maximum, low word zero"""
    m = np.ascontiguousarray(np.max(np.abs(x), axis=1, keepdims=True))
    hi = m.view(np.uint64) >> np.uint64(32)
    mf = np.where(hi > 0, ((hi + np.uint64(1)) << np.uint64(32)).view(np.float64), 0.0)
    inv = np.where(mf > 0, FULL / np.where(mf > 0, mf, 1.0), 0.0)
    q = np.rint(x * inv).astype(np.int64)
    return _planes(q), mf * (1.0 / FULL)


def _quant_rows_exact(x):
    """the input layer uses the exact float64 maximum"""
    m = np.max(np.abs(x), axis=1, keepdims=True)
    inv = np.where(m > 0, FULL / np.where(m > 0, m, 1.0), 0.0)
    q = np.rint(x * inv).astype(np.int64)
    return _planes(q), m * (1.0 / FULL)


def _level_sum(da, dw):
    """T = sum_{l<=3} acc_l 256^(3-l), acc_l = sum_{i+j=l} d_i . e_j^T (exact integers)"""
    T = np.zeros((da[0].shape[0], dw[0].shape[0]), dtype=np.int64)
    for i in range(4):
        for j in range(4 - i):
            T += (da[i] @ dw[j].T) << (8 * (3 - i - j))
    return T.astype(np.float64)


def supported(layer_dims):
    L = len(layer_dims) - 1
    g = 1
    for l in range(L):
        N = layer_dims[l + 1]
        npad = (N + 15) // 16 * 16 if l == L - 1 else (N + 31) // 32 * 32
        if g == 2 and npad > 64:
            return False
        if l == L - 1 and npad > 16:
            return False
        if npad <= 192:
            g = 1
        else:
            if l == L - 1:
                return False
            g = 2
    return True


def mirror_coefficients(spec, x):
    """spec: NN emulator spec (oracle/emulators.py layout); x[rows, n_in] raw inputs -> coefficients[rows, n_out]"""
    xmin, xmax = np.asarray(spec["xmin"], float), np.asarray(spec["xmax"], float)
    rinv = 1.0 / (xmax - xmin)
    h = (x - xmin) * rinv - 0.5
    layers = [(np.asarray(l["W"], float), np.asarray(l["b"], float)) for l in spec["layers"]]
    slope = float(spec["slope"])
    L = len(layers)
    groups = [(_quant_rows_exact(h))]  # list of (planes, scale) per K-group
    kcuts = [0, h.shape[1]]
    for il, (W, b) in enumerate(layers):
        N = W.shape[0]
        m = np.max(np.abs(W), axis=1, keepdims=True)
        s_w = m / FULL
        qw = np.rint(W / np.where(s_w > 0, s_w, 1.0)).astype(np.int64)
        sw24 = (s_w * 16777216.0).T  # [1, N]
        y = np.broadcast_to(b[None, :], (h.shape[0], N)).astype(np.float64).copy()
        # groups are accumulated first to last on the device: y = fma(T0, sa0 sw, b), then fma(T1, sa1 sw, y)
        for g, (planes, sa) in enumerate(groups):
            dw = _planes(qw[:, kcuts[g] : kcuts[g + 1]])
            T = _level_sum(planes, dw)
            y = T * (sa * sw24) + y
        if il == L - 1:
            return y
        h = np.where(y > 0, y, slope * y)
        npad = (N + 31) // 32 * 32
        if npad <= 192:
            groups = [_quant_rows(h)]
            kcuts = [0, N]
        else:
            groups = [_quant_rows(h[:, :128]), _quant_rows(h[:, 128:])]
            kcuts = [0, 128, N]
    raise AssertionError("unreachable")
