"""Host emulation of the K1v micro-op plan (tplx_gpu_stage_vec_plan): one row at a time, the accumulator and the dense register
file modelled exactly as vecvm.cuh uses them. Test infrastructure: it lets the planner (accumulator chains, fused compare/filter,
dropped stores, slot renumbering) be checked against the oracle without a GPU."""
import math
import os
import re
import struct

M64 = (1 << 64) - 1
NOSLOT = 0xFFFF
A_CONST, B_CONST = 64, 128
X_A_ACC, X_B_ACC, X_NOSTORE, X_FILTER, X_A_MASK = 1, 2, 4, 8, 16
EC_ZERODIV = 136


def _vops():
    src = open(os.path.join(os.path.dirname(__file__), "..", "tuplex_b200", "csrc", "vecvm.cuh")).read()
    body = re.search(r"enum VOp : uint32_t \{(.*?)\};", src, re.S).group(1)
    names = [t.split("=")[0].strip() for t in body.replace("\n", " ").split(",") if t.strip()]
    return {n: i for i, n in enumerate(names)}


V = _vops()
NAME = {i: n for n, i in V.items()}


def s64(x):
    x &= M64
    return x - (1 << 64) if x >> 63 else x


def f(x):
    return struct.unpack("<d", struct.pack("<Q", x & M64))[0]


def u(d):
    return struct.unpack("<Q", struct.pack("<d", d))[0]


def _floordiv(a, b):
    return (a // b) & M64


def _f2i(d):
    return int(d) & M64 if math.isfinite(d) else 0


READS = {"V_NOP": 0, "V_LDCOL": 0, "V_LDI": 0, "V_LDROW": 0, "V_RAISE": 0, "V_MOV": 1, "V_INEG": 1, "V_IABS": 1, "V_FNEG": 1, "V_FABS": 1,
         "V_I2F": 1, "V_F2I": 1, "V_BNOT": 1, "V_ISHRK": 1, "V_IANDK": 1, "V_FILTER": 1, "V_SEL": 7}  # default: a and b
RAISING = {"V_IFLOORDIV", "V_IMOD", "V_FDIV", "V_FMOD", "V_FFLOORDIV"}


def run_row(uops, cols, row, row_index):
    """-> ('out', slots) | ('drop',) | ('exc', code, opidx). cols: list of raw 64-bit column values per row."""
    slots = {}
    acc = 0
    for uo in uops:
        name = NAME[uo["vop"]]
        xf, fl = uo["xflags"], uo["flags"]
        guarded = uo["guard"] != NOSLOT
        if guarded and slots.get(uo["guard"], 0) == 0:
            continue
        # the accumulator is operand a's home: an a that is not already there is loaded into it first (vecvm.cuh, VX_LOAD_A), so a
        # b taken from the accumulator is only meaningful together with a from the accumulator (x op x)
        if READS.get(name, 3) & 1 and not (xf & X_A_ACC):
            acc = (uo["imm2"] & M64) if fl & A_CONST else slots.get(uo["a"], 0)
        A = acc
        B = acc if xf & X_B_ACC else ((uo["imm"] & M64) if fl & B_CONST else slots.get(uo["b"], 0))
        if xf & X_B_ACC:
            assert xf & X_A_ACC, "b from the accumulator needs a from the accumulator"
        if name in RAISING:
            assert not (xf & (X_A_ACC | X_B_ACC | X_NOSTORE)), "raising micro-ops keep operands and result in slots"
        if name == "V_LDCOL":
            r = cols[uo["imm"]][row] & M64
        elif name == "V_LDI":
            r = uo["imm"] & M64
        elif name == "V_LDROW":
            r = row_index & M64
        elif name == "V_MOV":
            r = A
        elif name == "V_SEL":
            r = A if slots.get(uo["c"], 0) else B
        elif name == "V_IADD":
            r = (A + B) & M64
        elif name == "V_ISUB":
            r = (A - B) & M64
        elif name == "V_IMUL":
            r = (A * B) & M64
        elif name == "V_INEG":
            r = (-A) & M64
        elif name == "V_IAND":
            r = A & B
        elif name == "V_IOR":
            r = A | B
        elif name == "V_IXOR":
            r = A ^ B
        elif name == "V_ISHL":
            r = (A << (B & 63)) & M64
        elif name == "V_ISHR":
            r = (s64(A) >> (B & 63)) & M64
        elif name == "V_ISHRK":
            r = (s64(A) >> (uo["imm"] & 63)) & M64
        elif name == "V_IANDK":
            r = A & (uo["imm"] & M64)
        elif name == "V_IABS":
            r = (-A) & M64 if s64(A) < 0 else A
        elif name == "V_FADD":
            r = u(f(A) + f(B))
        elif name == "V_FSUB":
            r = u(f(A) - f(B))
        elif name == "V_FMUL":
            r = u(f(A) * f(B))
        elif name == "V_FNEG":
            r = A ^ (1 << 63)
        elif name == "V_FABS":
            r = A & ~(1 << 63) & M64
        elif name == "V_I2F":
            r = u(float(s64(A)))
        elif name == "V_F2I":
            r = _f2i(f(A))
        elif name == "V_BAND":
            r = int(A != 0 and B != 0)
        elif name == "V_BOR":
            r = int(A != 0 or B != 0)
        elif name == "V_BNOT":
            r = int(A == 0)
        elif name.startswith("V_ICMP_"):
            x = A & (uo["imm2"] & M64) if xf & X_A_MASK else A
            x, y = s64(x), s64(B)
            r = int({"EQ": x == y, "NE": x != y, "LT": x < y, "LE": x <= y, "GT": x > y, "GE": x >= y}[name[7:]])
        elif name.startswith("V_FCMP_"):
            x, y = f(A), f(B)
            r = int({"EQ": x == y, "NE": x < y or x > y, "LT": x < y, "LE": x <= y, "GT": x > y, "GE": x >= y}[name[7:]])
        elif name == "V_IFLOORDIV":
            if s64(B) == 0:
                return ("exc", EC_ZERODIV, uo["opidx"])
            r = _floordiv(s64(A), s64(B))
        elif name == "V_IMOD":
            if s64(B) == 0:
                return ("exc", EC_ZERODIV, uo["opidx"])
            r = (s64(A) % s64(B)) & M64
        elif name == "V_FDIV":
            if f(B) == 0.0:
                return ("exc", EC_ZERODIV, uo["opidx"])
            r = u(f(A) / f(B))
        elif name == "V_FMOD":
            if f(B) == 0.0:
                return ("exc", EC_ZERODIV, uo["opidx"])
            m = math.fmod(f(A), f(B))
            if m != 0.0 and ((m < 0.0) != (f(B) < 0.0)):
                m = m + f(B)
            r = u(m)
        elif name == "V_FFLOORDIV":
            xi, yi = int(f(A)), int(f(B))
            if f(B) == 0.0 or yi == 0:
                return ("exc", EC_ZERODIV, uo["opidx"])
            r = u(float(xi // yi))
        elif name == "V_FILTER":
            if A == 0:
                return ("drop",)
            continue
        elif name == "V_RAISE":
            return ("exc", uo["imm"] & 0xFFFF, uo["opidx"])
        else:
            raise AssertionError(f"unknown micro-op {name}")
        acc = r
        if not (xf & X_NOSTORE):
            assert uo["dst"] != NOSLOT
            slots[uo["dst"]] = r
        if xf & X_FILTER:
            assert not guarded
            if r == 0:
                return ("drop",)
    return ("out", slots)


def run(uops, out_slots, cols, n):
    """-> (output rows as tuples of raw 64-bit values, [(row, code, opidx)])"""
    out, exc = [], []
    for row in range(n):
        res = run_row(uops, cols, row, row)
        if res[0] == "out":
            out.append(tuple(res[1].get(s, 0) for s in out_slots))
        elif res[0] == "exc":
            exc.append((row, res[1], res[2]))
    return out, exc
