"""Jump-ahead polynomials for MT19937, so that the GPU can generate R's Mersenne-Twister stream in parallel segments.

R's default generator is one serial stream (libR main/RNG.c, MT19937): word x[n+624] is a
GF(2)-linear function of x[n], x[n+1], x[n+397].  Seen from index 1 on (the low 31 bits of the very first state word
are never used), the word sequence y[m] = x[m+1] satisfies the linear recurrence whose characteristic polynomial
phi(t) has degree 19937, so for any distance D

    y[m + D] = XOR over the set bits i of g_D of y[m + i],     g_D(t) = t^D mod phi(t),   deg g_D < 19937

(Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008, "Efficient jump ahead for F2-linear random number
generators").  Since the y[m + i] are simply the next 19937 words of the stream, the state at the start of segment p
(words x[p S .. p S + 623]) is a binary convolution of g_{pS-1} with the first 20560 words — embarrassingly parallel
— after which every segment runs the ordinary recurrence on its own (csrc/kernels.cuh, mt_jump_kernel /
mt_segment_kernel).

This module computes phi (Berlekamp-Massey on one bit of the word sequence) and the table g_{pS-1}, p = 1..P-1, with
plain Python integers as GF(2) polynomials (bit i = coefficient of t^i).  `build.py` writes the table next to the
library (mt_jump.bin); tests/test_mt_jump.py checks phi and the table against a straightforward MT19937.
"""
import os
import struct

N, M = 624, 397
DEGREE = 19937
SEGMENT = 65536          # words per segment
TABLE_SEGMENTS = 256     # polynomials for p = 1 .. TABLE_SEGMENTS-1 (16.7 M words; longer streams end serially)
WORDS_PER_POLY = N       # 19937 bits in 624 words
MAGIC = b"MTJ1"


def mt_words(state, count):
    """x[0..count): the word sequence that starts with the 624 state words (untempered)."""
    x = list(state) + [0] * max(0, count - N)
    for n in range(N, count):
        y = (x[n - N] & 0x80000000) | (x[n - N + 1] & 0x7fffffff)
        x[n] = x[n - N + M] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
    return x[:count]


def seed_state(seed=5489):
    """Any state will do for finding phi; this is init_genrand(seed) of the MT19937 reference code."""
    s = [seed & 0xffffffff]
    for i in range(1, N):
        s.append((1812433253 * (s[-1] ^ (s[-1] >> 30)) + i) & 0xffffffff)
    return s


def berlekamp_massey(bits):
    """Connection polynomial C (C_0 = 1) and linear complexity L of a bit sequence: s_n = XOR_{i=1..L} C_i s_{n-i}."""
    C, B, L, m = 1, 1, 0, -1
    R = 0  # bit i = s_{n-i}
    for n, s in enumerate(bits):
        R = (R << 1) | s
        if bin(C & R).count("1") & 1:
            T = C
            C ^= B << (n - m)
            if 2 * L <= n:
                L, B, m = n + 1 - L, T, n
    return C, L


def characteristic_polynomial():
    """phi(t), degree 19937, as an int: sum_i phi_i y[m+i] = 0 for the word sequence y[m] = x[m+1], m >= 0."""
    x = mt_words(seed_state(), 1 + 2 * DEGREE + 64)
    bits = [(w >> 1) & 1 for w in x[1:]]   # one bit of every word, from index 1 on
    C, L = berlekamp_massey(bits)
    if L != DEGREE:
        raise RuntimeError("linear complexity %d, expected %d" % (L, DEGREE))
    phi = 0
    for i in range(L + 1):   # phi_i = C_{L-i}
        if (C >> (L - i)) & 1:
            phi |= 1 << i
    return phi


def _reduce(r, phi):
    top = phi.bit_length() - 1
    while r.bit_length() - 1 >= top:
        r ^= phi << (r.bit_length() - 1 - top)
    return r


def _mulmod(a, b, phi):
    if bin(a).count("1") > bin(b).count("1"):
        a, b = b, a
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return _reduce(r, phi)


def _powmod_t(e, phi):
    """t^e mod phi"""
    r, base = 1, 2
    while e:
        if e & 1:
            r = _mulmod(r, base, phi)
        e >>= 1
        if e:
            base = _mulmod(base, base, phi)
    return r


def jump_polynomials(segment=SEGMENT, count=TABLE_SEGMENTS, phi=None):
    """[g_{p*segment-1} for p in 1..count-1]"""
    phi = phi or characteristic_polynomial()
    step = _powmod_t(segment, phi)
    g = _powmod_t(segment - 1, phi)
    out = [g]
    for _ in range(2, count):
        g = _mulmod(g, step, phi)
        out.append(g)
    return out


def poly_words(g):
    return [(g >> (32 * k)) & 0xffffffff for k in range(WORDS_PER_POLY)]


def write_table(path, segment=SEGMENT, count=TABLE_SEGMENTS):
    polys = jump_polynomials(segment, count)
    with open(path + ".tmp", "wb") as f:
        f.write(MAGIC + struct.pack("<III", segment, count - 1, WORDS_PER_POLY))
        for g in polys:
            f.write(struct.pack("<%dI" % WORDS_PER_POLY, *poly_words(g)))
    os.replace(path + ".tmp", path)
    return path


def read_table(path):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == MAGIC
    segment, n, wpp = struct.unpack("<III", raw[4:16])
    words = struct.unpack("<%dI" % (n * wpp), raw[16:16 + 4 * n * wpp])
    return segment, [words[k * wpp:(k + 1) * wpp] for k in range(n)]


if __name__ == "__main__":
    import sys
    import time
    t0 = time.time()
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "mt_jump.bin")
    write_table(out)
    print("wrote %s in %.1f s" % (out, time.time() - t0))
