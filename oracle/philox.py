"""CPU restatement of the counter-based generator of deep_dantzig_b200/csrc/generate.cu (test infrastructure).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 1.09) restated with
numpy uint64 arithmetic; known answers are the three vectors of Random123's ``kat_vectors`` for philox4x32_10.
The reference itself draws from numpy's legacy MT19937 stream (src/data/randomlp_dataset.py:76-84), which a
counter-based generator cannot reproduce; this module pins the *integer* stream of the device generator and restates
its uniform -> normal transform so device instances can be checked to a few ulp."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_A, STREAM_X0, STREAM_EPS, STREAM_C, STREAM_MASK = 0, 1, 2, 3, 4


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32-valued, key: (k0, k1) ints -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) & MASK for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def _u53(hi, lo):
    return ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))


def normal_pairs(key, inst, stream, npairs):
    """Normals for elements 0..2*npairs-1 of `stream` of instance `inst` (Box-Muller, as generate.cu)."""
    ctr = np.zeros((npairs, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(npairs)
    ctr[:, 1] = stream
    ctr[:, 2] = inst & 0xFFFFFFFF
    ctr[:, 3] = (inst >> 32) & 0xFFFFFFFF
    o = philox4x32_10(ctr, (key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF))
    u1 = (_u53(o[:, 0], o[:, 1]) + np.uint64(1)).astype(np.float64) * 2.0 ** -53
    u2 = _u53(o[:, 2], o[:, 3]).astype(np.float64) * 2.0 ** -53
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * npairs)
    z[0::2] = rad * np.cos(2.0 * np.pi * u2)
    z[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return z


def uniforms(key, inst, stream, npairs):
    ctr = np.zeros((npairs, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(npairs); ctr[:, 1] = stream
    ctr[:, 2] = inst & 0xFFFFFFFF; ctr[:, 3] = (inst >> 32) & 0xFFFFFFFF
    o = philox4x32_10(ctr, (key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF))
    u = np.empty(2 * npairs)
    u[0::2] = _u53(o[:, 0], o[:, 1]).astype(np.float64) * 2.0 ** -53
    u[1::2] = _u53(o[:, 2], o[:, 3]).astype(np.float64) * 2.0 ** -53
    return u


def generate_instance(key, inst, m, n, density=1.0):
    """A[m,n], b[m], c[n], x0[n] of instance `inst` of stream `key` (restates generate_A_kernel / generate_bc_kernel)."""
    per = m * n
    A = normal_pairs(key, inst, STREAM_A, (per + 1) // 2)[:per]
    if density < 1.0:
        keep = uniforms(key, inst, STREAM_MASK, (per + 1) // 2)[:per] < density
        A = np.where(keep, A, 0.0)
    A = A.reshape(m, n)
    x0 = normal_pairs(key, inst, STREAM_X0, (n + 1) // 2)[:n]
    eps = normal_pairs(key, inst, STREAM_EPS, (m + 1) // 2)[:m]
    c = np.abs(normal_pairs(key, inst, STREAM_C, (n + 1) // 2)[:n])
    b = A.dot(x0) + np.abs(eps)
    return A, b, c, x0
