#!/usr/bin/env python3
"""Design check for an ADDITIVE (padding-based) shared-memory layout of the NTT tile, the next lever named in
DESIGN.md §8.  No GPU needed.

Today the tile uses an XOR swizzle; every shared-memory access of a round costs one LOP3 on the saturated ALU
pipe (the Gray-code walk in ntt_round).  With an additive layout  addr(e) = e + Σ_i ((e >> s_i) * m_i)  the 16
addresses of a thread's group are base + j·stride (an IMAD on the idle FMA pipe, or an LDS immediate when the
window position is a template parameter).  This script re-implements the kernel's lane → element maps
(ntt_round, ntt_load_batches, store_perm of ronkathon_b200/csrc/ntt_kernel.cuh) and checks, for a candidate
pad list:
  1. addr is injective and the padded tile fits the shared-memory budget,
  2. every phase is bank-conflict-free: the 16 lanes of each half-warp hit 16 distinct 8-byte bank pairs,
  3. within a round, addr(e0 + j·2^wb) - addr(e0) is the same for every thread (so it is a stride/immediate).
Usage: python tools/smem_layout_audit.py            (checks the shapes of the 2^24 and 4096×2^16 workloads)"""
import itertools
import sys


def bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def make_addr(pads):
    def addr(e):
        return e + sum(((e >> s) * m) for s, m in pads)
    return addr


def conflict_free(addrs16):
    return len({a & 15 for a in addrs16}) == 16


def check_shape(name, mode, log_m, log_c, log_c2, nthr, pads, verbose=True):
    """mode: 'pass1' | 'pass2' | 'single'.  tile index = [batch | field (log_m) | column (log_c)]."""
    assert log_m % 4 == 0, "this candidate only covers log_m ≡ 0 (mod 4)"
    tile_log = log_m + log_c
    T = 1 << tile_log
    addr = make_addr(pads)
    problems = []
    # 1. injective
    seen = {}
    for e in range(T):
        a = addr(e)
        if a in seen:
            problems.append(f"addr collision {seen[a]} / {e}")
            break
        seen[a] = e
    words = max(seen) + 1
    # 2a/3. rounds
    groups = T // 16
    for r in range(log_m // 4):
        wb = log_c + log_m - 4 * (r + 1)
        low = (1 << wb) - 1
        deltas = None
        for t0 in range(0, groups, 16):   # half-warps of consecutive group indices (tid, tid+nthr, … keep alignment)
            e0s = [(((t >> wb) << (wb + 4)) | (t & low)) for t in range(t0, t0 + 16)]
            for j in range(16):
                if not conflict_free([addr(e0 + (j << wb)) for e0 in e0s]):
                    problems.append(f"round wb={wb}: bank conflict at j={j}, groups {t0}..")
                    break
            d = [tuple(addr(e0 + (j << wb)) - addr(e0) for j in range(16)) for e0 in e0s]
            if deltas is None:
                deltas = d[0]
            if any(x != deltas for x in d):
                problems.append(f"round wb={wb}: offsets differ between threads (not a stride)")
                break
        if deltas and any(deltas[j] != j * deltas[1] for j in range(16)):
            problems.append(f"round wb={wb}: offsets are thread-independent but not linear in j: {deltas}")
    # 2b. load phase: lanes hold consecutive tile indices e = tid + j·nthr
    for base in range(0, T, 16):
        if not conflict_free([addr(base + l) for l in range(16)]):
            problems.append(f"load phase: bank conflict at e={base}..")
            break
    # 2c. store phase: g = tid + j·nthr enumerates the HBM order, e = store_perm(g)
    def store_perm(g):
        if mode == "single":
            M1 = (1 << log_m) - 1
            return ((g >> log_m) << log_m) | bitrev(g & M1, log_m)
        if mode == "pass1":
            cl = log_c + log_c2
            rem = g & ((1 << cl) - 1)
            k1 = ((g >> cl) << log_c2) | (rem & ((1 << log_c2) - 1))
            return (bitrev(k1, log_m) << log_c) | (rem >> log_c2)
        return (bitrev(g >> log_c, log_m) << log_c) | (g & ((1 << log_c) - 1))
    for base in range(0, T, 16):
        if not conflict_free([addr(store_perm(base + l)) for l in range(16)]):
            problems.append(f"store phase: bank conflict at g={base}..")
            break
    if verbose:
        kib = words * 8 / 1024
        print(f"{name:34s} pads={pads}  tile {T * 8 // 1024} KiB -> {kib:.1f} KiB  " + ("OK" if not problems else "FAIL"))
        for p in problems[:4]:
            print("     ", p)
    return not problems, words


def search(name, mode, log_m, log_c, log_c2, nthr):
    """Smallest pad list (shifts on window boundaries log_c + 4k, small multipliers) that passes every check."""
    shifts = [log_c + 4 * k for k in range(1, log_m // 4 + 1)]
    best = None
    for n_terms in (0, 1, 2, 3):
        for ss in itertools.combinations(shifts, n_terms):
            for ms in itertools.product((1, 2, 4, 8), repeat=n_terms):
                pads = list(zip(ss, ms))
                ok, words = check_shape(name, mode, log_m, log_c, log_c2, nthr, pads, verbose=False)
                if ok and (best is None or words < best[1]):
                    best = (pads, words)
        if best:
            break
    if best:
        check_shape(name, mode, log_m, log_c, log_c2, nthr, best[0])
    else:
        print(f"{name:34s} no pad list with <= 3 terms found")
    return best


if __name__ == "__main__":
    # 2^24: N1 = N2 = 2^12; pass 1 tile 2^14 (4 columns, C2 = 2), pass 2 tile 2^13 (2 columns)
    search("2^24 pass 1 (m=12, c=2, c2=1)", "pass1", 12, 2, 1, 512)
    search("2^24 pass 2 (m=12, c=1)", "pass2", 12, 1, 0, 256)
    # 4096 x 2^16: N1 = N2 = 2^8; pass 1 tile 2^14 (64 columns, C2 = 32), pass 2 tile 2^13 (32 rows)
    search("2^16 pass 1 (m=8, c=6, c2=5)", "pass1", 8, 6, 5, 512)
    search("2^16 pass 2 (m=8, c=5)", "pass2", 8, 5, 0, 256)
    # single-pass 2^12 and 2^8 (16 transforms per tile)
    search("single 2^12 (m=12, c=0)", "single", 12, 0, 0, 128)
    sys.exit(0)
