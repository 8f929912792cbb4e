"""A second, independent restatement -- pure Python big integers, no shared code with oracle/*.c -- of the pieces that have
NO reference-held golden vector: ProofOfPossession (src/proofs/possession.rs:93-163) and PublicKeySet::from_participants
(src/sharing/key_set.rs:87-144, lagrange_coefficients src/sharing/mod.rs:139-170).  It exists to cross-check the C oracle
(tests/test_pyref_crosscheck.py), which is what the GPU is compared with: two restatements written from the reference
source and the RFCs independently agreeing is the pin for these two objects.

TEST INFRASTRUCTURE ONLY.  Everything below is written from: RFC 9496 (ristretto255), RFC 8032 (edwards25519 constants),
the STROBE v1.0.2 specification as used by merlin 3.0.0 (Strobe128: AD / meta-AD / PRF / KEY), FIPS 202 (Keccak-f[1600]).
"""

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
INVSQRT_A_MINUS_D = None        # set below


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 4.2: (was_square, r) with r = sqrt(u / v) or sqrt(i u / v)."""
    u %= P
    v %= P
    r = (u * pow(v, 3, P) * pow(u * pow(v, 7, P), (P - 5) // 8, P)) % P
    check = (v * r * r) % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = (r * SQRT_M1) % P
    return (correct or flipped), _abs(r)


INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, (-1 - D) % P)[1]


class Point:
    """edwards25519 point in extended coordinates; equality / encoding are ristretto255's."""
    __slots__ = ("x", "y", "z", "t")

    def __init__(self, x, y, z, t):
        self.x, self.y, self.z, self.t = x % P, y % P, z % P, t % P

    @staticmethod
    def identity():
        return Point(0, 1, 1, 0)

    def __add__(self, o):        # add-2008-hwcd-3 (a = -1)
        a = ((self.y - self.x) * (o.y - o.x)) % P
        b = ((self.y + self.x) * (o.y + o.x)) % P
        c = (self.t * 2 * D * o.t) % P
        d = (self.z * 2 * o.z) % P
        e, f, g, h = b - a, d - c, d + c, b + a
        return Point(e * f, g * h, f * g, e * h)

    def __neg__(self):
        return Point(-self.x, self.y, self.z, -self.t)

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, k):
        k %= L
        acc, base = Point.identity(), self
        while k:
            if k & 1:
                acc = acc + base
            base = base + base
            k >>= 1
        return acc

    def __eq__(self, o):         # RFC 9496 4.3.3
        return (self.x * o.y - self.y * o.x) % P == 0 or (self.y * o.y - self.x * o.x) % P == 0

    def encode(self):            # RFC 9496 4.3.2
        x0, y0, z0, t0 = self.x, self.y, self.z, self.t
        u1 = ((z0 + y0) * (z0 - y0)) % P
        u2 = (x0 * y0) % P
        _, invsqrt = sqrt_ratio_m1(1, (u1 * u2 * u2) % P)
        den1, den2 = (invsqrt * u1) % P, (invsqrt * u2) % P
        z_inv = (den1 * den2 * t0) % P
        ix0, iy0 = (x0 * SQRT_M1) % P, (y0 * SQRT_M1) % P
        enchanted = (den1 * INVSQRT_A_MINUS_D) % P
        if _is_neg(t0 * z_inv):
            x, y, den_inv = iy0, ix0, enchanted
        else:
            x, y, den_inv = x0, y0, den2
        if _is_neg(x * z_inv):
            y = (-y) % P
        s = _abs(den_inv * (z0 - y))
        return s.to_bytes(32, "little")

    @staticmethod
    def decode(b):               # RFC 9496 4.3.1; None when the encoding is invalid
        if len(b) != 32:
            return None
        s = int.from_bytes(b, "little")
        if s >= P or s & 1:
            return None
        ss = (s * s) % P
        u1, u2 = (1 - ss) % P, (1 + ss) % P
        u2_sqr = (u2 * u2) % P
        v = (-(D * u1 * u1) - u2_sqr) % P
        was_square, invsqrt = sqrt_ratio_m1(1, (v * u2_sqr) % P)
        den_x, den_y = (invsqrt * u2) % P, (invsqrt * invsqrt * u2 * v) % P
        x = _abs(2 * s * den_x)
        y = (u1 * den_y) % P
        t = (x * y) % P
        if not was_square or _is_neg(t) or y == 0:
            return None
        return Point(x, y, 1, t)


_BY = (4 * pow(5, P - 2, P)) % P
_BX = _abs(sqrt_ratio_m1((_BY * _BY - 1) % P, (D * _BY * _BY + 1) % P)[1])      # the even root is the base point's x
G = Point(_BX, _BY, 1, _BX * _BY)


# ---------------------------------------------------------------- Keccak-f[1600] / Strobe-128 / Merlin

_RC, _ROT = [], [[0] * 5 for _ in range(5)]


def _init_keccak():
    r = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            r = ((r << 1) ^ ((r >> 7) * 0x71)) % 256
            if r & 2:
                rc ^= 1 << ((1 << j) - 1)
        _RC.append(rc)
    x, y = 1, 0
    for t in range(24):
        _ROT[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_init_keccak()
_M64 = (1 << 64) - 1


def _rol(v, n):
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(state):
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    out = bytearray(200)
    for x in range(5):
        for y in range(5):
            out[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")
    return out


class Strobe128:
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.state = keccak_f1600(st)
        self.pos = self.pos_begin = self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        self.state = keccak_f1600(self.state)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert flags == self.cur_flags
            return
        assert not flags & self.FLAG_T
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript (3.0.0) + TranscriptForGroup (src/proofs/mod.rs:29-57)."""

    def __init__(self, label):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def append_u64(self, label, x):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)

    def start_proof(self, label):
        self.append_message(b"dom-sep", label)

    def append_element(self, label, point):
        self.append_message(label, point.encode())

    def challenge_scalar(self, label):           # Ristretto::scalar_from_random_bytes: 64 bytes, reduced mod l
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % L


# ---------------------------------------------------------------- ProofOfPossession (possession.rs:93-163)

OK, MALFORMED, CHALLENGE_MISMATCH, MALFORMED_PARTICIPANT_KEYS = 0, 1, 2, 7


def pop_prove(secrets, key_bytes, label, random_blocks):
    """from_keys :93-128.  random_blocks: one 64-byte block per key (SecretKey::generate = generate_scalar: wide reduce)."""
    t = Transcript(label)
    t.start_proof(b"multi_pop")
    for kb in key_bytes:
        t.append_message(b"K", kb)
    rs = []
    for block in random_blocks:
        r = int.from_bytes(block, "little") % L
        t.append_element(b"R", G * r)
        rs.append(r)
    c = t.challenge_scalar(b"c")
    return c.to_bytes(32, "little") + b"".join(((r + x * c) % L).to_bytes(32, "little") for r, x in zip(rs, secrets))


def pop_verify(key_bytes, label, proof):
    """PublicKey::from_bytes (keys/mod.rs:161-176) per key, scalar parsing (serde), then verify :135-163."""
    keys = [Point.decode(kb) for kb in key_bytes]
    if any(k is None or k == Point.identity() for k in keys):
        return MALFORMED
    scalars = [int.from_bytes(proof[32 * i:32 * i + 32], "little") for i in range(1 + len(keys))]
    if any(s >= L for s in scalars):
        return MALFORMED
    c, responses = scalars[0], scalars[1:]
    t = Transcript(label)
    t.start_proof(b"multi_pop")
    for kb in key_bytes:
        t.append_message(b"K", kb)
    for key, s in zip(keys, responses):
        t.append_element(b"R", key * ((-c) % L) + G * s)          # vartime_double_mul_generator(-c, K, s)
    return OK if t.challenge_scalar(b"c") == c else CHALLENGE_MISMATCH


# ---------------------------------------------------------------- PublicKeySet::from_participants (key_set.rs:87-144)

def lagrange_coefficients(indexes):
    """sharing/mod.rs:139-170: (denominators^-1, scale); points are index + 1."""
    dens = []
    for i in indexes:
        sign, mag = False, 1
        for o in indexes:
            if i > o:
                sign, mag = not sign, mag * (i - o)
            elif i < o:
                mag *= o - i
            else:
                mag *= i + 1
        dens.append((-mag) % L if sign else mag % L)
    scale = 1
    for i in indexes:
        scale = (scale * (i + 1)) % L
    return [pow(d, L - 2, L) for d in dens], scale


def keyset_from_participants(shares, threshold, key_bytes):
    """-> (verdict, shared key bytes or None)."""
    keys = [Point.decode(kb) for kb in key_bytes]
    if any(k is None or k == Point.identity() for k in keys):
        return MALFORMED, None
    indexes = list(range(threshold))
    denominators, scale = lagrange_coefficients(indexes)
    start = keys[:threshold]
    shared = Point.identity()
    for d, k in zip(denominators, start):
        shared = shared + k * d
    shared = shared * scale
    inverses = [pow(v, L - 2, L) for v in range(1, shares + 1)]
    for x in range(threshold, shares):
        key_scale = 1
        for idx in indexes:
            key_scale = (key_scale * (x - idx)) % L
        key_dens = [(d * (idx + 1) * inverses[x - idx - 1]) % L for idx, d in enumerate(denominators)]
        if threshold % 2 == 0:
            key_scale = (-key_scale) % L
        interp = Point.identity()
        for d, k in zip(key_dens, start):
            interp = interp + k * d
        if not (interp * key_scale == keys[x]):
            return MALFORMED_PARTICIPANT_KEYS, None
    return OK, shared.encode()
