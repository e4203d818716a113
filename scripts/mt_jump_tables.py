"""Jump-ahead tables for the Mersenne-Twister key stream (R_COMPAT mode of the FastMCD start).

MT19937's state transition is linear over GF(2).  With phi its characteristic polynomial (degree 19937,
found here by Berlekamp-Massey on one output bit) and g_e(x) = x^e mod phi, the raw word sequence obeys
    x[n + e] = XOR_{i : bit i of g_e} x[n + i]            for every n >= 1
(the low 31 bits of x[0] of a freshly seeded state are not part of the state, hence n >= 1).  The CUDA
generator cuts one update_beta_scale_k call's N draws into blocks of B = 624 * 110 words; block p starts
from the window  x[pB + j] = XOR_{i in g_{pB-1}} x[i + j + 1],  j = 0..623,  so every block is produced by
its own CTA.  This script computes g_{pB-1} for p = 1..P, checks them against a directly generated
stream, and writes sampleqc_b200/csrc/sqc_mt_jump.inc.  Pure Python / numpy, about a minute.
"""
import os
import sys
import time

import numpy as np

N, M = 624, 397
DEG = 19937
CHUNKS_PER_BLOCK = 110
B = N * CHUNKS_PER_BLOCK
P = 148


def r_seed_state(seed):
    s = seed & 0xFFFFFFFF
    for _ in range(50):
        s = (69069 * s + 1) & 0xFFFFFFFF
    out = []
    for _ in range(625):
        s = (69069 * s + 1) & 0xFFFFFFFF
        out.append(s)
    return np.array(out[1:], dtype=np.uint32)


def raw_stream(state, n_words):
    """x[0..n_words): x[0..623] = state, x[n+624] = x[n+397] ^ twist(x[n], x[n+1])"""
    x = np.zeros(n_words + N, dtype=np.uint32)
    x[:N] = state
    UP, LO, MAT = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)
    pos = N
    while pos < n_words:
        base = pos - N
        for lo, hi in ((0, 227), (227, 454), (454, 624)):     # dependency distance is 227
            a = x[base + lo:base + hi]; b = x[base + lo + 1:base + hi + 1]
            y = (a & UP) | (b & LO)
            x[pos + lo:pos + hi] = x[base + lo + M:base + hi + M] ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), MAT, np.uint32(0))
        pos += N
    return x[:n_words]


def berlekamp_massey(bits):
    """connection polynomial C (int, bit i = c_i, c_0 = 1) and its length L of a GF(2) sequence"""
    C, Bp, L, m = 1, 1, 0, 1
    R = 0
    mask = (1 << (DEG + 64)) - 1
    for n, a in enumerate(bits):
        R = ((R << 1) | int(a)) & mask
        d = bin(C & R).count("1") & 1 if False else (C & R).bit_count() & 1
        if d == 0:
            m += 1
        elif 2 * L <= n:
            T = C
            C ^= Bp << m
            L = n + 1 - L
            Bp = T
            m = 1
        else:
            C ^= Bp << m
            m += 1
    return C, L


def polymod(r, phi, deg):
    while True:
        d = r.bit_length() - 1
        if d < deg:
            return r
        r ^= phi << (d - deg)


def polymulmod(a, b, phi, deg):
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return polymod(r, phi, deg)


def polypowx(e, phi, deg):
    """x^e mod phi"""
    result, base = 1, 2
    while e:
        if e & 1:
            result = polymulmod(result, base, phi, deg)
        base = polymulmod(base, base, phi, deg)
        e >>= 1
    return result


def main():
    t0 = time.time()
    st = r_seed_state(12345)
    x = raw_stream(st, 2 * DEG + 2000)
    bits = (x[1:2 * DEG + 200] & 1).astype(np.uint8)
    C, L = berlekamp_massey(bits)
    assert L == DEG, L
    # characteristic polynomial in shift-operator form: phi(E) = E^L + sum_i c_i E^(L-i)  (reciprocal of C)
    phi = 0
    for i in range(L + 1):
        if (C >> i) & 1:
            phi |= 1 << (L - i)
    print(f"phi: degree {phi.bit_length() - 1}, weight {phi.bit_count()}  ({time.time() - t0:.1f}s)")
    taps = [i for i in range(DEG + 1) if (phi >> i) & 1]
    print(f"phi taps: top {taps[-1]}, next {taps[-2]} (gap {taps[-1] - taps[-2]}), lowest {taps[:3]}")
    XB = polypowx(B, phi, DEG)
    g = polypowx(B - 1, phi, DEG)
    table = []
    for p in range(1, P + 1):
        table.append(g)
        g = polymulmod(g, XB, phi, DEG)
    print(f"{P} jump polynomials  ({time.time() - t0:.1f}s)")
    # ---- verification against a directly generated stream (a different seed) -------------------------
    st = r_seed_state(0)
    need = P * B + N + DEG + 8
    x = raw_stream(st, need)
    for p in (1, 2, 3, 17, 64, P):
        gp = table[p - 1]
        idx = np.array([i for i in range(DEG) if (gp >> i) & 1], dtype=np.int64)
        win = np.zeros(N, dtype=np.uint32)
        for j0 in range(0, N, 64):
            js = np.arange(j0, min(j0 + 64, N))
            win[js] = np.bitwise_xor.reduce(x[idx[:, None] + js[None, :] + 1], axis=0)
        assert np.array_equal(win, x[p * B:p * B + N]), f"jump {p} failed"
    print(f"verified jumps against the direct stream  ({time.time() - t0:.1f}s)")
    words = np.zeros((P, N), dtype=np.uint32)
    for p in range(P):
        gp = table[p]
        for w in range(N):
            words[p, w] = (gp >> (32 * w)) & 0xFFFFFFFF
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sampleqc_b200", "csrc", "sqc_mt_jump.inc")
    with open(out, "w") as f:
        f.write("// generated by scripts/mt_jump_tables.py - do not edit.\n")
        f.write(f"// row p-1 = x^(p*B-1) mod phi_MT19937, B = {B} words ({CHUNKS_PER_BLOCK} chunks of 624), bit i of word w = coefficient of x^(32w+i)\n")
        f.write(f"#define SQC_MT_JUMP_ROWS {P}\n#define SQC_MT_CHUNKS_PER_BLOCK {CHUNKS_PER_BLOCK}\n")
        f.write(f"static const unsigned int sqc_mt_jump_table[{P}][624] = {{\n")
        for p in range(P):
            f.write("{" + ",".join("0x%08xu" % v for v in words[p]) + "},\n")
        f.write("};\n")
        # the characteristic polynomial itself: the library derives further jump polynomials (x^e mod phi for the
        # call-to-call and rank offsets of the distributed key stream) on the host at session creation
        f.write(f"// phi_MT19937 (degree {DEG}, weight {phi.bit_count()}): exponents of its terms, ascending\n")
        f.write(f"#define SQC_MT_PHI_TERMS {len(taps)}\n")
        f.write(f"static const int sqc_mt_phi_terms[{len(taps)}] = {{" + ",".join(str(t) for t in taps) + "};\n")
    print("wrote", out, f"({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
