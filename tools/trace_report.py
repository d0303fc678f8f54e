import sys, numpy as np
lines = open(sys.argv[1]).read().splitlines()
t = np.array([int(x) for x in lines if x and not x.startswith('#')], dtype=np.int64)
for l in lines:
    if l.startswith('##'):
        print(' ', l)
    elif l.startswith('#'):
        e = [int(x) for x in l.split() if x.lstrip('-').isdigit()]
        print(f'  kernel entry -> prologue done {e[1]-e[0]} cyc; prologue done -> first step {t[0]-e[1]}; last step end -> exit {e[2]-t[-1]}; entry -> exit {e[2]-e[0]}')
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
per = 3 * L * S
nr = len(t) // per
if nr < 3:
    print("too few rounds", len(t)); sys.exit(0)
a = t[: nr * per].reshape(nr, L, S, 3)
print(f"{nr} rounds of warp 0 (warpgroup 0): median cycles, rounds 1..")
for l in range(L):
    for q in range(S):
        w = np.median(a[1:, l, q, 1] - a[1:, l, q, 0]); k = np.median(a[1:, l, q, 2] - a[1:, l, q, 1])
        print(f"  layer {l+1} slot {q}: wait for MMA {w:6.0f}   epilogue {k:6.0f}")
print("  round period:", np.median(np.diff(a[:, 0, 0, 0])), "cycles for", S, "tiles per warpgroup")
