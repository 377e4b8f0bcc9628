"""Dev tool: count shared-memory wavefronts of the Stockham FFT stages (wso_device.cuh) for the
padding rule pad(i) = i + i/16.  64-bit (float2) accesses: a warp request is served 16 lanes at a
time; bank = (element index) mod 16.  Prints measured/ideal wavefronts per stage (1.00 = conflict free)."""
import sys
import numpy as np

PLANS = {4: [16], 5: [2, 16], 6: [4, 16], 7: [8, 16], 8: [16, 16], 9: [2, 16, 16], 10: [4, 16, 16],
         11: [8, 16, 16], 12: [16, 16, 16], 13: [2, 16, 16, 16], 14: [4, 16, 16, 16]}


def pad(i):
    return i + (i >> 4)


def wavefronts(addr):
    """addr: physical element indices of up to 16 lanes."""
    return np.bincount(addr % 16, minlength=16).max()


def analyse(logn, B=4, verbose=True):
    N = 1 << logn
    LS = N + N // 16 + 4
    T = B * N // 16
    NS = 1
    res = []
    for R in PLANS[logn]:
        JN = N // R
        NB = 16 // R
        ld = st = ideal = 0
        for t0 in range(0, T, 16):
            tids = np.arange(t0, min(t0 + 16, T))
            for i in range(NB):
                u = tids + T * i
                line = u // JN
                j = u % JN
                k = j % NS
                base = (j // NS) * NS * R + k
                for r in range(R):
                    ld += wavefronts(line * LS + pad(j + r * JN))
                    st += wavefronts(line * LS + pad(base + r * NS))
                    ideal += 1
        res.append((R, ld / ideal, st / ideal))
        NS *= R
    if verbose:
        print(f"N={N:6d} B={B}: " + "  ".join(f"R{R}: ld x{a:.2f} st x{b:.2f}" for R, a, b in res))
    return res


if __name__ == "__main__":
    for logn in range(4, 15):
        N = 1 << logn
        B = max(1, min(8, 16384 // N))
        analyse(logn, B)
