#!/usr/bin/env python3
"""Generates elastic_elgamal_b200/csrc/fe_ptx.cuh: the device-tuned GF(2^255-19) multiply and square.

The instruction sequence is built once as a list of (op, dst, a, b, c) tuples, then
  * emitted as one inline-PTX block per function (so every carry chain is contiguous), and
  * executed by a small PTX-subset interpreter in this script against Python big integers
    (`python tools/gen_fe_ptx.py --check`), because there is no GPU in the build container.

Scheme (8 saturated 32-bit limbs): 32x32->64 products are accumulated with mad.lo.cc / madc.hi.cc chains
into two accumulator rows, `e` (products landing on even limb positions) and `o` (odd positions, stored
shifted by one limb), so that every product is aligned with a register pair and ptxas can fuse each
lo/hi pair into one IMAD.WIDE with carry-in/out.  The rows are merged with one add chain, the 512-bit
result is folded with 2^256 = 38 (mod p).
"""
import argparse
import pathlib
import random

P = 2**255 - 19
M32 = 0xffffffff


class Prog:
    def __init__(self):
        self.ins = []
        self.regs = set()

    def r(self, name):
        self.regs.add(name)
        return name

    def emit(self, op, dst, *src):
        self.ins.append((op, dst, src))

    # --- interpreter -------------------------------------------------------------------------
    def run(self, env):
        v = dict(env)
        cc = 0

        def val(x):
            return x if isinstance(x, int) else v[x]

        for op, dst, src in self.ins:
            s = [val(x) for x in src]
            if op == "mov":
                v[dst] = s[0]
            elif op == "mul.lo":
                v[dst] = (s[0] * s[1]) & M32
            elif op == "mul.hi":
                v[dst] = (s[0] * s[1]) >> 32
            elif op in ("mad.lo.cc", "madc.lo.cc", "mad.hi.cc", "madc.hi.cc", "madc.lo", "madc.hi", "mad.lo", "mad.hi"):
                prod = s[0] * s[1]
                part = (prod & M32) if ".lo" in op else (prod >> 32)
                cin = cc if op.startswith("madc") else 0
                t = part + s[2] + cin
                v[dst] = t & M32
                if op.endswith(".cc"):
                    cc = t >> 32
                else:
                    assert t >> 32 == 0 or True
            elif op in ("add.cc", "addc.cc", "addc", "add"):
                cin = cc if op.startswith("addc") else 0
                t = s[0] + s[1] + cin
                v[dst] = t & M32
                if op.endswith(".cc"):
                    cc = t >> 32
            elif op in ("sub.cc", "subc.cc", "subc"):
                bin_ = cc if op.startswith("subc") else 0
                t = s[0] - s[1] - bin_
                v[dst] = t & M32
                if op.endswith(".cc"):
                    cc = 1 if t < 0 else 0
            elif op == "xor":
                v[dst] = s[0] ^ s[1]
            elif op == "shl":
                v[dst] = (s[0] << s[1]) & M32
            elif op == "shr":
                v[dst] = s[0] >> s[1]
            elif op == "shf.l":                       # funnel shift left: upper word of (hi:lo) << n
                v[dst] = (((s[1] << 32) | s[0]) << s[2] >> 32) & M32
            else:
                raise ValueError(op)
        return v

    # --- PTX text ----------------------------------------------------------------------------
    def ptx(self, inputs, outputs):
        """inputs/outputs: ordered lists of register names bound to asm operands."""
        opmap = {}
        for i, name in enumerate(outputs):
            opmap[name] = "%%%d" % i
        for i, name in enumerate(inputs):
            opmap[name] = "%%%d" % (len(outputs) + i)
        internal = sorted(self.regs - set(inputs) - set(outputs), key=lambda s: (s[0], int(s[1:]) if s[1:].isdigit() else 0))
        lines = ["{"]
        if internal:
            lines.append(".reg .u32 " + ", ".join(internal) + ";")
        for op, dst, src in self.ins:
            def f(x):
                return str(x) if isinstance(x, int) else opmap.get(x, x)
            if op == "mov":
                lines.append("mov.u32 %s, %s;" % (f(dst), f(src[0])))
            elif op == "shl":
                lines.append("shl.b32 %s, %s, %s;" % (f(dst), f(src[0]), f(src[1])))
            elif op == "shr":
                lines.append("shr.u32 %s, %s, %s;" % (f(dst), f(src[0]), f(src[1])))
            elif op == "shf.l":
                lines.append("shf.l.wrap.b32 %s, %s, %s, %s;" % (f(dst), f(src[0]), f(src[1]), f(src[2])))
            elif op == "xor":
                lines.append("xor.b32 %s, %s, %s;" % (f(dst), f(src[0]), f(src[1])))
            else:
                lines.append("%s.u32 %s, %s;" % (op, f(dst), ", ".join(f(x) for x in src)))
        lines.append("}")
        return lines


def fold(p, t, out):
    """out[0..7] = t[0..7] + 38 * t[8..15], fully folded into 256 bits."""
    r = [p.r("r%d" % i) for i in range(8)]
    q = [p.r("q%d" % i) for i in range(8)]
    c1, c2, c3 = p.r("c1"), p.r("c2"), p.r("c3")
    # even-position products t[8], t[10], t[12], t[14] times 38 added on top of t[0..7]
    for k in range(4):
        p.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", r[2 * k], t[8 + 2 * k], 38, t[2 * k])
        p.emit("madc.hi.cc", r[2 * k + 1], t[8 + 2 * k], 38, t[2 * k + 1])
    p.emit("addc", c1, 0, 0)
    # odd-position products t[9], t[11], t[13], t[15] times 38 (each < 2^38)
    for k in range(4):
        p.emit("mul.lo", q[2 * k], t[9 + 2 * k], 38)
        p.emit("mul.hi", q[2 * k + 1], t[9 + 2 * k], 38)
    p.emit("add.cc", r[1], r[1], q[0])
    for k in range(1, 7):
        p.emit("addc.cc", r[1 + k], r[1 + k], q[k])
    p.emit("addc", c2, q[7], c1)            # carry limb (weight 2^256), <= 38 + 37 + 2
    # r += 38 * c2
    p.emit("mad.lo.cc", r[0], c2, 38, r[0])
    for k in range(1, 8):
        p.emit("addc.cc", r[k], r[k], 0)
    p.emit("addc", c3, 0, 0)
    p.emit("mad.lo", out[0], c3, 38, r[0])  # wrapped value is tiny: stays inside limb 0
    for k in range(1, 8):
        p.emit("mov", out[k], r[k])


def fold_shift(p, t, out):
    """Same result as fold(), but 38 * t[8..15] = (h << 5) + (h << 2) + (h << 1) is built with funnel shifts and add
    chains on the ALU pipe instead of eight wide multiply-adds on the (saturated) multiplier pipe."""
    h = t[8:16]
    r = [p.r("r%d" % i) for i in range(9)]
    c3 = p.r("c3")
    first = True
    for sh in (5, 2, 1):
        sv = [p.r("s%d_%d" % (sh, i)) for i in range(9)]
        p.emit("shl", sv[0], h[0], sh)
        for k in range(1, 8):
            p.emit("shf.l", sv[k], h[k - 1], h[k], sh)
        p.emit("shr", sv[8], h[7], 32 - sh)
        base = t if first else r
        p.emit("add.cc", r[0], base[0], sv[0])
        for k in range(1, 8):
            p.emit("addc.cc", r[k], base[k], sv[k])
        if first:
            p.emit("addc", r[8], sv[8], 0)
        else:
            p.emit("addc", r[8], r[8], sv[8])
        first = False
    # r[8] <= 31 + 3 + 1 + 3 carries: fold it once more (38 * r8 < 2^12)
    p.emit("mad.lo.cc", r[0], r[8], 38, r[0])
    for k in range(1, 8):
        p.emit("addc.cc", r[k], r[k], 0)
    p.emit("addc", c3, 0, 0)
    p.emit("mad.lo", out[0], c3, 38, r[0])
    for k in range(1, 8):
        p.emit("mov", out[k], r[k])


FOLD = fold


def merge(p, e, o, t):
    """t = e + (o << 32)"""
    p.emit("mov", t[0], e[0])
    p.emit("add.cc", t[1], e[1], o[0])
    for k in range(2, 15):
        p.emit("addc.cc", t[k], e[k], o[k - 1])
    p.emit("addc", t[15], e[15], o[14])


def gen_mul():
    p = Prog()
    a = [p.r("a%d" % i) for i in range(8)]
    b = [p.r("b%d" % i) for i in range(8)]
    e = [p.r("e%d" % i) for i in range(16)]
    o = [p.r("o%d" % i) for i in range(16)]
    t = [p.r("t%d" % i) for i in range(16)]
    out = [p.r("z%d" % i) for i in range(8)]
    for i in range(8, 16):
        p.emit("mov", e[i], 0)
        p.emit("mov", o[i], 0)
    # row 0
    for j in range(0, 8, 2):
        p.emit("mul.lo", e[j], a[j], b[0]); p.emit("mul.hi", e[j + 1], a[j], b[0])
    for j in range(0, 8, 2):
        p.emit("mul.lo", o[j], a[j + 1], b[0]); p.emit("mul.hi", o[j + 1], a[j + 1], b[0])

    def chain(acc, base, limbs, bi):
        for k, j in enumerate(limbs):
            p.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[base + 2 * k], a[j], bi, acc[base + 2 * k])
            p.emit("madc.hi.cc", acc[base + 2 * k + 1], a[j], bi, acc[base + 2 * k + 1])
        if base + 8 < 16:
            p.emit("addc", acc[base + 8], acc[base + 8], 0)

    for i in range(1, 8):
        if i % 2 == 0:
            chain(e, i, [0, 2, 4, 6], b[i])
            chain(o, i, [1, 3, 5, 7], b[i])
        else:
            chain(o, i - 1, [0, 2, 4, 6], b[i])
            chain(e, i + 1, [1, 3, 5, 7], b[i])
    merge(p, e, o, t)
    FOLD(p, t, out)
    return p, a + b, out


def product_rows(p, a, b, e, o, off):
    """Schoolbook n x n limb product (n = len(a), even) into the even / odd rows at limb offset `off`:
    e[off + k] holds the products landing on even positions, o[off + k] (position off + k + 1) the odd ones.
    Rows must be zero above the first 2 limbs touched by row 0 (callers zero-initialise e/o[off + n .. off + 2n))."""
    n = len(a)
    ev, od = list(range(0, n, 2)), list(range(1, n, 2))
    for j in ev:
        p.emit("mul.lo", e[off + j], a[j], b[0]); p.emit("mul.hi", e[off + j + 1], a[j], b[0])
    for j in ev:
        p.emit("mul.lo", o[off + j], a[j + 1], b[0]); p.emit("mul.hi", o[off + j + 1], a[j + 1], b[0])

    def chain(acc, base, limbs, bi):
        for k, j in enumerate(limbs):
            p.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[off + base + 2 * k], a[j], bi, acc[off + base + 2 * k])
            p.emit("madc.hi.cc", acc[off + base + 2 * k + 1], a[j], bi, acc[off + base + 2 * k + 1])
        if base + n < 2 * n:
            p.emit("addc", acc[off + base + n], acc[off + base + n], 0)

    for i in range(1, n):
        if i % 2 == 0:
            chain(e, i, ev, b[i])
            chain(o, i, od, b[i])
        else:
            chain(o, i - 1, ev, b[i])
            chain(e, i + 1, od, b[i])


def abs_diff(p, x, y, name):
    """d = |x - y| (4 limbs), m = 0xffffffff when x < y else 0."""
    n = len(x)
    d = [p.r("%s%d" % (name, i)) for i in range(n)]
    m, dummy = p.r(name + "m"), p.r(name + "c")
    for i in range(n):
        p.emit("sub.cc" if i == 0 else "subc.cc", d[i], x[i], y[i])
    p.emit("subc", m, 0, 0)
    for i in range(n):
        p.emit("xor", d[i], d[i], m)
    p.emit("add.cc", dummy, m, m)                    # CF = (x < y)
    for i in range(n):
        p.emit("addc.cc" if i < n - 1 else "addc", d[i], d[i], 0)
    return d, m


def gen_mul_karatsuba():
    """a b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) 2^128 + z2 2^256 with three 4 x 4 limb products: 48 + 8 wide
    multiply-adds instead of 64 + 8, paid for with ~75 more add / logic instructions on the ALU pipe."""
    p = Prog()
    a = [p.r("a%d" % i) for i in range(8)]
    b = [p.r("b%d" % i) for i in range(8)]
    e = [p.r("e%d" % i) for i in range(16)]
    o = [p.r("o%d" % i) for i in range(16)]
    f = [p.r("f%d" % i) for i in range(8)]
    g = [p.r("g%d" % i) for i in range(8)]
    t = [p.r("t%d" % i) for i in range(16)]
    zm = [p.r("m%d" % i) for i in range(8)]
    mid = [p.r("n%d" % i) for i in range(9)]
    u = [p.r("u%d" % i) for i in range(16)]
    out = [p.r("z%d" % i) for i in range(8)]
    for i in list(range(4, 8)) + list(range(12, 16)):
        p.emit("mov", e[i], 0)
        p.emit("mov", o[i], 0)
    for i in range(4, 8):
        p.emit("mov", f[i], 0)
        p.emit("mov", g[i], 0)
    da, ma = abs_diff(p, a[0:4], a[4:8], "x")
    db, mb = abs_diff(p, b[4:8], b[0:4], "y")
    ms, dummy = p.r("ms"), p.r("mc")
    p.emit("xor", ms, ma, mb)
    product_rows(p, a[0:4], b[0:4], e, o, 0)
    product_rows(p, a[4:8], b[4:8], e, o, 8)
    product_rows(p, da, db, f, g, 0)
    merge(p, e, o, t)                                 # t = z0 + z2 2^256
    # zm = f + (g << 32)
    p.emit("mov", zm[0], f[0])
    p.emit("add.cc", zm[1], f[1], g[0])
    for k in range(2, 7):
        p.emit("addc.cc", zm[k], f[k], g[k - 1])
    p.emit("addc", zm[7], f[7], g[6])
    # mid = z0 + z2 +- zm   (9 limbs, non-negative)
    p.emit("add.cc", mid[0], t[0], t[8])
    for k in range(1, 8):
        p.emit("addc.cc", mid[k], t[k], t[8 + k])
    p.emit("addc", mid[8], 0, 0)
    for k in range(8):
        p.emit("xor", zm[k], zm[k], ms)
    p.emit("add.cc", dummy, ms, ms)                   # CF = 1 when the cross term is negative (two's complement + 1)
    for k in range(8):
        p.emit("addc.cc", mid[k], mid[k], zm[k])
    p.emit("addc", mid[8], mid[8], ms)
    # u = t + mid 2^128
    for k in range(4):
        p.emit("mov", u[k], t[k])
    p.emit("add.cc", u[4], t[4], mid[0])
    for k in range(5, 13):
        p.emit("addc.cc", u[k], t[k], mid[k - 4])
    p.emit("addc.cc", u[13], t[13], 0)
    p.emit("addc.cc", u[14], t[14], 0)
    p.emit("addc", u[15], t[15], 0)
    FOLD(p, u, out)
    return p, a + b, out


def gen_sq():
    p = Prog()
    a = [p.r("a%d" % i) for i in range(8)]
    e = [p.r("e%d" % i) for i in range(16)]
    o = [p.r("o%d" % i) for i in range(16)]
    t = [p.r("t%d" % i) for i in range(16)]
    u = [p.r("u%d" % i) for i in range(16)]
    out = [p.r("z%d" % i) for i in range(8)]
    for i in range(16):
        p.emit("mov", e[i], 0)
        p.emit("mov", o[i], 0)
    # off-diagonal products a_i a_j (i < j), position i + j
    for i in range(7):
        ev = [j for j in range(i + 1, 8) if (i + j) % 2 == 0]
        od = [j for j in range(i + 1, 8) if (i + j) % 2 == 1]
        for acc, js, shift in ((e, ev, 0), (o, od, 1)):
            if not js:
                continue
            # consecutive j differ by 2 -> consecutive register pairs
            base = i + js[0] - shift
            for k, j in enumerate(js):
                p.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", acc[base + 2 * k], a[i], a[j], acc[base + 2 * k])
                p.emit("madc.hi.cc", acc[base + 2 * k + 1], a[i], a[j], acc[base + 2 * k + 1])
            top = base + 2 * len(js)
            if top < 16:
                p.emit("addc", acc[top], acc[top], 0)
    merge(p, e, o, t)
    # double: u = 2 t  (t < 2^511)
    p.emit("add.cc", u[0], t[0], t[0])
    for k in range(1, 15):
        p.emit("addc.cc", u[k], t[k], t[k])
    p.emit("addc", u[15], t[15], t[15])
    # diagonal
    for i in range(8):
        p.emit("mad.lo.cc" if i == 0 else "madc.lo.cc", u[2 * i], a[i], a[i], u[2 * i])
        p.emit("madc.hi.cc" if i < 7 else "madc.hi", u[2 * i + 1], a[i], a[i], u[2 * i + 1])
    FOLD(p, u, out)
    return p, a, out


def check(n=20000):
    rnd = random.Random(7)
    pm, in_m, out_m = gen_mul()
    ps, in_s, out_s = gen_sq()
    pk, in_k, out_k = gen_mul_karatsuba()
    edge = [0, 1, 2**256 - 1, 2**255 - 19, 2**255 - 20, 2**256 - 38, 2**256 - 39, 38, 2**32 - 1, 2**224, (2**256 - 1) ^ (2**128),
            2**128 - 1, 2**128, (2**128 - 1) << 128, (1 << 128) | 1, ((2**128 - 1) << 128) | 1, (1 << 255) | (2**128 - 1), 2**127, (2**127) << 128 | 2**127]

    def limbs(x):
        return [(x >> (32 * i)) & M32 for i in range(8)]

    def value(v, names):
        return sum(v[nm] << (32 * i) for i, nm in enumerate(names))

    cases = [(x, y) for x in edge for y in edge] + [(rnd.getrandbits(256), rnd.getrandbits(256)) for _ in range(n)]
    for x, y in cases:
        env = {("a%d" % i): l for i, l in enumerate(limbs(x))}
        env.update({("b%d" % i): l for i, l in enumerate(limbs(y))})
        got = value(pm.run(env), out_m)
        assert got < 2**256 and got % P == (x * y) % P, (hex(x), hex(y))
        got = value(pk.run(env), out_k)
        assert got < 2**256 and got % P == (x * y) % P, ("karatsuba", hex(x), hex(y))
        env = {("a%d" % i): l for i, l in enumerate(limbs(x))}
        got = value(ps.run(env), out_s)
        assert got < 2**256 and got % P == (x * x) % P, hex(x)
    print("fe_ptx check ok:", len(cases), "cases; mul", len(pm.ins), "ins, sq", len(ps.ins), "ins")


def emit_function(name, prog, inputs, outputs, args):
    lines = prog.ptx(inputs, outputs)
    body = "\n".join('        "%s\\n\\t"' % l for l in lines)
    outs = ", ".join('"=r"(r.v[%d])' % i for i in range(8))
    if args == 2:
        ins = ", ".join('"r"(a.v[%d])' % i for i in range(8)) + ",\n          " + ", ".join('"r"(b.v[%d])' % i for i in range(8))
        sig = "fe &r, const fe &a, const fe &b"
    else:
        ins = ", ".join('"r"(a.v[%d])' % i for i in range(8))
        sig = "fe &r, const fe &a"
    # outputs are written only at the very end of the block, after all inputs are dead -> no early-clobber needed,
    # but '=&r' keeps ptxas from aliasing an output with an input that is still live in the fold.
    outs = outs.replace('"=r"', '"=&r"')
    return ("__device__ __forceinline__ void %s(%s) {\n    asm(\n%s\n        : %s\n        : %s);\n}\n" % (name, sig, body, outs, ins))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default=str(pathlib.Path(__file__).resolve().parent.parent / "elastic_elgamal_b200" / "csrc" / "fe_ptx.cuh"))
    args = ap.parse_args()
    check(2000 if not args.check else 20000)
    pm, in_m, out_m = gen_mul()
    ps, in_s, out_s = gen_sq()
    text = ("// fe_ptx.cuh -- GENERATED by tools/gen_fe_ptx.py (do not edit; re-run the generator).\n"
            "// Device-tuned GF(2^255-19) multiply / square: mad.lo.cc / madc.hi.cc carry chains on aligned\n"
            "// even/odd accumulator rows (fused by ptxas into IMAD.WIDE with carry), 2^256 = 38 fold.\n"
            "// Included by fe.cuh for device compilation only; the sequence is verified against Python big\n"
            "// integers by the generator's PTX-subset interpreter and on the GPU against the oracle.\n"
            "#pragma once\nnamespace eg {\n\n")
    text += emit_function("fe_mul_ptx", pm, in_m, out_m, 2) + "\n"
    text += emit_function("fe_sq_ptx", ps, in_s, out_s, 1) + "\n"
    pk, in_k, out_k = gen_mul_karatsuba()
    text += emit_function("fe_mul_ptx_k", pk, in_k, out_k, 2) + "\n"
    # variant with the 2^256 = 38 fold on the ALU pipe (funnel shifts + add chains); selected by EG_FE_SHIFT_FOLD
    global FOLD
    FOLD = fold_shift
    check(2000 if not args.check else 20000)
    pm, in_m, out_m = gen_mul()
    ps, in_s, out_s = gen_sq()
    FOLD = fold
    text += emit_function("fe_mul_ptx_sf", pm, in_m, out_m, 2) + "\n"
    text += emit_function("fe_sq_ptx_sf", ps, in_s, out_s, 1) + "\n}  // namespace eg\n"
    pathlib.Path(args.out).write_text(text)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
