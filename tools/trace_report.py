import sys, numpy as np
lines = open(sys.argv[1]).read().splitlines()
t = np.array([int(x) for x in lines if x and not x.startswith('#')], dtype=np.int64)
for l in lines:
    if l.startswith('###'):
        continue
    elif l.startswith('##'):
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

iss = np.array([[int(x) for x in l.split()[2:]] for l in lines if l.startswith('#### issuer')], dtype=np.int64)
if len(iss) >= 3 * L * S:
    ni = len(iss) // (L * S)
    b_ = iss[: ni * L * S].reshape(ni, L, S, 6)
    print(f"issuer warp of warpgroup 0, {ni} rounds: median cycles per step (rounds 1..)")
    print("    step                wait-ready  wait-tile  issue-MMA  commit  refill   step-total")
    for l in range(L):
        for q in range(S):
            d = np.median(np.diff(b_[1:, l, q, :], axis=-1), axis=0)
            print(f"    layer {l+1} slot {q}:   {d[0]:9.0f} {d[1]:10.0f} {d[2]:10.0f} {d[3]:7.0f} {d[4]:7.0f} {d.sum():10.0f}")
    print("    issuer round period:", np.median(np.diff(b_[:, 0, 0, 0])))
cta = [(int(l.split()[2]), int(l.split()[4]), int(l.split()[6])) for l in lines if l.startswith('### cta')]
if cta:
    import numpy as np
    life = np.array([c[1] for c in cta]); sm = np.array([c[2] for c in cta]); b = np.array([c[0] for c in cta])
    print(f"  per-CTA lifetime us: min {life.min()/1e3:.1f} median {np.median(life)/1e3:.1f} max {life.max()/1e3:.1f}")
    n56 = b < (8192 % len(cta))
    print(f"  56-tile CTAs: median {np.median(life[n56])/1e3:.1f}  55-tile CTAs: median {np.median(life[~n56])/1e3:.1f}")
    order = np.argsort(life)
    print("  slowest 12 (cta, smid, us):", [(int(b[i]), int(sm[i]), round(life[i]/1e3,1)) for i in order[-12:]])
    print("  fastest 12 (cta, smid, us):", [(int(b[i]), int(sm[i]), round(life[i]/1e3,1)) for i in order[:12]])
    print("  corr(lifetime, smid) =", round(float(np.corrcoef(life, sm)[0,1]),3), " corr(lifetime, smid%2) =", round(float(np.corrcoef(life, sm%2)[0,1]),3))
