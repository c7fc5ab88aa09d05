"""Numerical model (numpy, CPU) of the INT8-slice route to the trailing update C -= A B on tcgen05 (the north-star's "tcgen05 / TMEM"
clause; tcgen05.mma has no FP64 kind).  Ozaki-style error-free splitting: every row of A and every column of B is scaled by a power of
two to |x| < 1 and cut into S signed slices of 7 bits (INT8 operands); slice products are EXACT in INT32 for K <= 2^17; the FP64 result
is the scaled sum of the slice-pair products with i + j < S (the rest is below the target accuracy).  The script measures the error of
that sum against the FP64 product, in the units the LU test uses (||A|| N eps), and counts the INT8 GEMMs it needs.
Not a kernel: it decides whether one is worth writing.  Run: python scripts/ozaki_model.py"""
import numpy as np

BITS = 7


def split(x, axis, S):
    """x = 2^e * sum_s q_s 2^(-BITS (s+1)),  q_s integer in [-2^BITS, 2^BITS], per row (axis=1) or column (axis=0) exponent e"""
    amax = np.abs(x).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)
    r = x / 2.0 ** e                                   # |r| <= 1/2
    q = []
    for s in range(S):
        r = r * 2.0 ** BITS
        qs = np.rint(r)                                # |qs| <= 2^(BITS-1) + rounding: fits INT8
        r = r - qs
        q.append(qs)
    return e, q


def product(a, b, S):
    ea, qa = split(a, 1, S)
    eb, qb = split(b, 0, S)
    acc = np.zeros((a.shape[0], b.shape[1]))
    ngemm = 0
    for i in range(S):
        for j in range(S - i):                         # i + j < S
            p = qa[i] @ qb[j]                          # exact: |p| <= K 2^(2 BITS) < 2^31 for K <= 2^17
            assert np.abs(p).max() < 2.0 ** 31
            acc += p * 2.0 ** (-BITS * (i + j + 2))
            ngemm += 1
    return acc * 2.0 ** ea * 2.0 ** eb, ngemm


def main():
    rng = np.random.default_rng(0)
    m, n, k = 384, 384, 512
    eps = 2.0 ** -53
    print(f"C = A B, {m} x {k} x {n}; error in units of max|A| max|B| K eps (the FP64 product itself sits at ~0.01-0.05 of it)")
    for name, a, b in (("uniform(-1/2,1/2)", rng.uniform(-.5, .5, (m, k)), rng.uniform(-.5, .5, (k, n))),
                       ("L21 / U12 of an LU (|L| <= 1, U grown)", np.tril(rng.uniform(-1, 1, (m, k))), rng.uniform(-1, 1, (k, n)) * rng.uniform(1, 30, (k, 1))),
                       ("rows spread over 2^20", rng.uniform(-1, 1, (m, k)) * 2.0 ** rng.integers(-20, 1, (1, k)), rng.uniform(-1, 1, (k, n)))):
        ref = a @ b
        scale = np.abs(a).max() * np.abs(b).max() * k * eps
        row = []
        for S in (5, 6, 7, 8, 9, 10):
            c, ng = product(a, b, S)
            row.append(f"S={S}: {np.abs(c - ref).max() / scale:9.2e} ({ng} GEMMs)")
        print(f"  {name}\n    " + "\n    ".join(row))


if __name__ == "__main__":
    main()
