"""Operand-split arithmetic of the two tcgen05 engines, emulated in numpy (CPU): what relative error does a
3-pass split GEMM make as a function of the operand magnitude?

  3xFP16 (csrc/tc16_gemm.cu): x = hi + lo / 2^11, hi = fp16(x), lo = fp16((x - hi) * 2^11); passes hi*hi + lo*hi + hi*lo
  3xTF32 (csrc/tc_gemm.cu)  : hi = x & 0xffffe000, lo = x - hi (the tensor core keeps lo's top 11 bits)

The forward operands (post-ReLU features, GRU states, weights) sit in fp16's normal range; gradients do not: below
6.1e-5 the fp16 `hi` is subnormal and the split degrades — hence the backward GEMMs run on the 3xTF32 engine
(csrc/mp_bwd.cu), and an exact power-of-two pre-scale (max |x| -> [1024, 2048)) would make the 3xFP16 engine safe
again.  Used by tests/test_split_precision.py; prints a table when run as a script."""
import numpy as np


def split16(x, lo_scale=2048.0):
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * np.float32(lo_scale)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64) / lo_scale


def split_tf32(x):
    hi = (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    lo = (x - hi)
    lo = (lo.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)      # tf32 operand truncation of the small term
    return hi.astype(np.float64), lo.astype(np.float64)


def gemm3(A, B, split):
    """C = A @ B^T with the 3-pass split; products / sums in float64 so only the operand representation shows."""
    ah, al = split(np.ascontiguousarray(A)); bh, bl = split(np.ascontiguousarray(B))
    return ah @ bh.T + al @ bh.T + ah @ bl.T


def pow2_scale(A, target=1536.0):
    amax = float(np.abs(A).max())
    return np.float32(2.0 ** np.floor(np.log2(target / amax))) if amax > 0 else np.float32(1.0)


def rel_errors(mag, K=2048, M=48, N=48, seed=0):
    rng = np.random.default_rng(seed)
    B = (rng.standard_normal((N, K)) * 0.5).astype(np.float32)            # activations / weights: O(1)
    A = (rng.standard_normal((M, K)) * mag).astype(np.float32)            # gradient-like operand of magnitude `mag`
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    den = np.abs(ref).max()
    s = pow2_scale(A)
    return {'fp16x3': np.abs(gemm3(A, B, split16) - ref).max() / den,
            'fp16x3_scaled': np.abs(gemm3(A * s, B, split16) / s - ref).max() / den,
            'tf32x3': np.abs(gemm3(A, B, split_tf32) - ref).max() / den,
            'fp32': np.abs((A @ B.T).astype(np.float64) - ref).max() / den}


if __name__ == '__main__':
    print('%10s %12s %14s %12s %12s' % ('|x|', '3xFP16', '3xFP16 scaled', '3xTF32', 'fp32 matmul'))
    for mag in (1e2, 1.0, 1e-2, 1e-4, 1e-6, 1e-8, 1e-10):
        e = rel_errors(mag)
        print('%10.0e %12.2e %14.2e %12.2e %12.2e' % (mag, e['fp16x3'], e['fp16x3_scaled'], e['tf32x3'], e['fp32']))
