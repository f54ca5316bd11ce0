"""CPU fuzz of the VM's string primitives (tuplex_b200/csrc/strops.cuh compiled for the host) against
CPython and against the oracle's libc-based restatement."""
import ctypes as ct
import os
import random
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def H(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("strops") / "strops_host.so")
    subprocess.check_call(["g++", "-O1", "-x", "c++", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "strops_host.cpp")])
    L = ct.CDLL(so)
    for f in ("h_find", "h_rfind"):
        getattr(L, f).restype = ct.c_longlong
    L.h_slice_index.restype = L.h_floordiv.restype = L.h_floormod.restype = ct.c_longlong
    L.h_slice_index.argtypes = L.h_floordiv.argtypes = L.h_floormod.argtypes = [ct.c_longlong, ct.c_longlong]
    L.h_lower4.restype = L.h_upper4.restype = ct.c_uint
    return L


def _buf(s: bytes, pad_front: int):
    """place s at an arbitrary alignment inside a 4-byte-multiple buffer (the library's memory contract)"""
    raw = b"\xAA" * pad_front + s + b"\xBB" * 3
    raw += b"\xCC" * ((-len(raw)) % 4)
    arr = np.frombuffer(raw, dtype=np.uint8).copy()
    # np allocations are at least 8-byte aligned
    assert arr.ctypes.data % 4 == 0
    return arr


def _case(s: str, flag: int) -> str:
    if flag == 1:
        return "".join(c.lower() if "A" <= c <= "Z" else c for c in s)
    if flag == 2:
        return "".join(c.upper() if "a" <= c <= "z" else c for c in s)
    return s


def test_find_rfind_eq_fuzz(H):
    rnd = random.Random(1234)
    alphabet = "abAB ,bd/$"
    for it in range(30000):
        hl = rnd.randint(0, 40)
        nl = rnd.choice([0, 1, 1, 2, 2, 3, 3, 4, 5, 9])
        h = "".join(rnd.choice(alphabet) for _ in range(hl))
        if hl and rnd.random() < 0.6:
            a = rnd.randint(0, hl - 1)
            n = h[a:a + nl]
            if rnd.random() < 0.3 and n:
                n = n[:-1] + rnd.choice(alphabet)
        else:
            n = "".join(rnd.choice(alphabet) for _ in range(nl))
        hf, nf = rnd.choice([0, 0, 1, 2]), rnd.choice([0, 0, 0, 1])
        hb, nb = _buf(h.encode(), rnd.randint(0, 7)), _buf(n.encode(), rnd.randint(0, 7))
        ho, no = int(np.where(hb == 0xAA)[0].size), int(np.where(nb == 0xAA)[0].size)
        args = (hb.ctypes.data_as(ct.c_void_p), ho, len(h), hf, nb.ctypes.data_as(ct.c_void_p), no, len(n), nf)
        hh, nn = _case(h, hf), _case(n, nf)
        assert H.h_find(*args) == hh.find(nn), (h, n, hf, nf)
        assert H.h_rfind(*args) == hh.rfind(nn), (h, n, hf, nf)
        assert bool(H.h_eq(*args)) == (hh == nn), (h, n, hf, nf)


def test_str_copy_fuzz(H):
    """str_copy (word-wise, case-mapped) == byte-wise copy for every source / destination alignment and never touches a byte
    outside [dst, dst + len)."""
    rnd = random.Random(99)
    alphabet = "abYZ09 ,./"
    for it in range(20000):
        n = rnd.randint(0, 45)
        h = "".join(rnd.choice(alphabet) for _ in range(n))
        flag = rnd.choice([0, 0, 1, 2])
        sb = _buf(h.encode(), rnd.randint(0, 7))
        so = int(np.where(sb == 0xAA)[0].size)
        dofs = rnd.randint(0, 7)
        dst = np.full(dofs + n + 9, 0xEE, dtype=np.uint8)
        H.h_copy(dst.ctypes.data_as(ct.c_void_p), dofs, sb.ctypes.data_as(ct.c_void_p), so, n, flag)
        assert dst[dofs:dofs + n].tobytes() == _case(h, flag).encode(), (h, flag, so, dofs)
        assert (dst[:dofs] == 0xEE).all() and (dst[dofs + n:] == 0xEE).all(), (h, so, dofs)


def test_case_words(H):
    rnd = random.Random(7)
    for _ in range(20000):
        w = rnd.getrandbits(32)
        b = w.to_bytes(4, "little")
        lo = bytes((c + 32) if 65 <= c <= 90 else c for c in b)
        up = bytes((c - 32) if 97 <= c <= 122 else c for c in b)
        assert H.h_lower4(w) == int.from_bytes(lo, "little")
        assert H.h_upper4(w) == int.from_bytes(up, "little")


def test_atoi_matches_oracle_and_reference_quirks(H):
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import pyoracle
    cases = ["12", " 7 ", "-3", "x9", "", "  ", "-", "0042", "1e3", "9 9", "77\t", "+5", "123456789012", "\n5\r", "--1", "5-",
             "99999999999999999999", "-9223372036854775808", " - 1", "1_0", "٣"]
    rnd = random.Random(3)
    for _ in range(3000):
        cases.append("".join(rnd.choice(" -0123456789x\t") for _ in range(rnd.randint(0, 6))))
    for s in cases:
        b = s.encode()
        buf = _buf(b, 1)
        out = ct.c_longlong()
        ok = H.h_atoi(buf.ctypes.data_as(ct.c_void_p), 1, len(b), ct.byref(out))
        ook, ov = pyoracle.atoi64(s) if b"\0" not in b else (None, None)
        assert bool(ok) == ook, s
        if ok:
            assert out.value == ov, s
    # the reference quirks are real: "-" parses to 0, "+5" is a ValueError (StringUtils.cc:22-63)
    assert pyoracle.atoi64("-") == (True, 0) and pyoracle.atoi64("+5")[0] is False


def test_slice_floor(H):
    for n in range(0, 7):
        s = "abcdef"[:n]
        for i in range(-9, 10):
            # Python clamps slice indices exactly like processSliceIndex (BlockGeneratorVisitor.cc:4618-4690)
            assert s[H.h_slice_index(i, n):] == s[i:]
    for a in range(-20, 21):
        for b in list(range(-7, 0)) + list(range(1, 8)):
            assert H.h_floordiv(a, b) == a // b and H.h_floormod(a, b) == a % b
