"""Per-butterfly SASS histogram of the register-pass loops of a transform kernel (developer tool).
usage: sass_bfly.py <binary> [kernel-regex]"""
import collections, re, subprocess, sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
import sass_loop as S

def main(path, pat="ntt_kernel"):
    for name, ins in S.kernels(path).items():
        if not re.search(pat, name):
            continue
        print(name[:100], len(ins))
        for lo, hi in S.loops(ins):
            body = [t for a, t in ins if lo <= a <= hi]
            if not 400 <= len(body) <= 1500:
                continue
            c = collections.Counter(S.opcode(t) for t in body)
            wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
            imad = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
            print("   loop %4d instr: %.2f/bfly; wide %.2f, other IMAD* %.2f, rest %.2f" % (len(body), len(body) / 32, wide / 32, imad / 32, (len(body) - wide - imad) / 32),
                  {k: round(v / 32, 2) for k, v in c.most_common(10)})

if __name__ == "__main__":
    main(*sys.argv[1:])
